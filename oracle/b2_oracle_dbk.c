/* b2_oracle_dbk.c -- TEST INFRASTRUCTURE: CPU restatement of JM 18.5's in-loop deblocking filter for frame pictures
 * (SURVEY 8f-3).  Follows JM/lencod/src/loopFilter.c:196-377 (DeblockMb: edge order, left / top picture-border rule,
 * 8x8-transform edges), JM/lencod/src/loop_filter_normal.c:52-165 / :167-283 (GetStrengthVer / GetStrengthHor: bS 4 / 3 at
 * intra edges, 2 where either 4x4 block holds coefficients, else reference-picture identity and |mv| differences >= 4 with the
 * list-swapped comparison of B pictures), :285-437 / :445-585 (EdgeLoopLumaVer / Hor) and :590-758 (EdgeLoopChromaVer / Hor),
 * with the alpha / beta / clip tables of JM/lencod/inc/loop_filter.h:34-46 (H.264 tables 8-16, 8-17).
 * Macroblocks in raster order, as DeblockFrame (loopFilter.c:103-111) walks them.  8-bit 4:2:0, DFDisableIdc 0 or 1.
 * Records = include/b2me.h b2dbk_mb / b2dbk_blk. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t intra, qp, qpc_u, qpc_v, transform8x8, disable; int8_t alpha_off, beta_off; uint16_t cbp_blk, pad_; } OrcDbkMb;
typedef struct { int16_t mv[2][2]; int16_t ref[2]; } OrcDbkBlk;

static const uint8_t DBK_ALPHA[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
static const uint8_t DBK_BETA[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
static const uint8_t DBK_CLIP[52][5] = {
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,1,1,1},{0,0,1,1,1},{0,1,1,1,1},
  {0,1,1,1,1},{0,1,1,1,1},{0,1,1,1,1},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,2,3,3},{0,1,2,3,3},{0,2,2,3,3},{0,2,2,4,4},{0,2,3,4,4},
  {0,2,3,4,4},{0,3,3,5,5},{0,3,4,6,6},{0,3,4,6,6},{0,4,5,7,7},{0,4,5,8,8},{0,4,6,9,9},{0,5,7,10,10},{0,6,8,11,11},{0,6,8,13,13},{0,7,10,14,14},{0,8,11,16,16},
  {0,9,12,18,18},{0,10,13,20,20},{0,11,15,23,23},{0,13,17,25,25}};

static int dbk_clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static int dbk_mvne(const OrcDbkBlk *a, int la, const OrcDbkBlk *b, int lb)
{ return (abs(a->mv[la][0] - b->mv[lb][0]) >= 4) | (abs(a->mv[la][1] - b->mv[lb][1]) >= 4); }

/* bS of the 4-sample segment between 4x4 blocks p (left / above) and q; mbedge: the segment lies on a macroblock edge */
static int dbk_strength(const OrcDbkMb *mp, const OrcDbkMb *mq, int cp, int cq, const OrcDbkBlk *bp, const OrcDbkBlk *bq, int mbedge)
{
  if (mp->intra || mq->intra) return mbedge ? 4 : 3;
  if (cp || cq) return 2;
  {
    const int p0 = bp->ref[0], p1 = bp->ref[1], q0 = bq->ref[0], q1 = bq->ref[1];
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
    if (p0 != p1) {
      if (p0 == q0) return dbk_mvne(bp, 0, bq, 0) | dbk_mvne(bp, 1, bq, 1);
      return dbk_mvne(bp, 0, bq, 1) | dbk_mvne(bp, 1, bq, 0);
    }
    return (dbk_mvne(bp, 0, bq, 0) | dbk_mvne(bp, 1, bq, 1)) && (dbk_mvne(bp, 0, bq, 1) | dbk_mvne(bp, 1, bq, 0));
  }
}

/* one line of samples across an edge: s[-k*st] = p(k-1), s[k*st] = q(k) */
static void dbk_luma_line(uint8_t *q, int st, int bs, int alpha, int beta, int c0)
{
  const int p0 = q[-st], p1 = q[-2 * st], p2 = q[-3 * st], q0 = q[0], q1 = q[st], q2 = q[2 * st];
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) {
    const int small = abs(q0 - p0) < ((alpha >> 2) + 2);
    const int ap = (abs(p0 - p2) < beta) & small, aq = (abs(q0 - q2) < beta) & small, s = p0 + q0;
    if (ap) { const int p3 = q[-4 * st]; q[-st] = (uint8_t)((q1 + ((p1 + s) << 1) + p2 + 4) >> 3); q[-2 * st] = (uint8_t)((p2 + p1 + s + 2) >> 2); q[-3 * st] = (uint8_t)((((p3 + p2) << 1) + p2 + p1 + s + 4) >> 3); }
    else q[-st] = (uint8_t)(((p1 << 1) + p0 + q1 + 2) >> 2);
    if (aq) { const int q3 = q[3 * st]; q[0] = (uint8_t)((p1 + ((q1 + s) << 1) + q2 + 4) >> 3); q[st] = (uint8_t)((q2 + q0 + p0 + q1 + 2) >> 2); q[2 * st] = (uint8_t)((((q3 + q2) << 1) + q2 + q1 + s + 4) >> 3); }
    else q[0] = (uint8_t)(((q1 << 1) + q0 + p1 + 2) >> 2);
  } else {
    const int avg = (p0 + q0 + 1) >> 1, ap = abs(p0 - p2) < beta, aq = abs(q0 - q2) < beta, tc = c0 + ap + aq;
    const int dif = dbk_clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    if (ap) q[-2 * st] = (uint8_t)(p1 + dbk_clip3(-c0, c0, (p2 + avg - (p1 << 1)) >> 1));
    if (dif) { q[-st] = (uint8_t)dbk_clip3(0, 255, p0 + dif); q[0] = (uint8_t)dbk_clip3(0, 255, q0 - dif); }
    if (aq) q[st] = (uint8_t)(q1 + dbk_clip3(-c0, c0, (q2 + avg - (q1 << 1)) >> 1));
  }
}
static void dbk_chroma_line(uint8_t *q, int st, int bs, int alpha, int beta, int c0)
{
  const int p0 = q[-st], p1 = q[-2 * st], q0 = q[0], q1 = q[st];
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) { q[-st] = (uint8_t)(((p1 << 1) + p0 + q1 + 2) >> 2); q[0] = (uint8_t)(((q1 << 1) + q0 + p1 + 2) >> 2); }
  else {
    const int tc = c0 + 1, dif = dbk_clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    if (dif) { q[-st] = (uint8_t)dbk_clip3(0, 255, p0 + dif); q[0] = (uint8_t)dbk_clip3(0, 255, q0 - dif); }
  }
}

/* y [H][W], u / v [H/2][W/2] filtered in place; mbs [mbh*mbw]; blks [H/4][W/4] */
void orc_deblock_frame(int W, int H, uint8_t *y, uint8_t *u, uint8_t *v, const OrcDbkMb *mbs, const OrcDbkBlk *blks)
{
  const int mbw = W / 16, mbh = H / 16, bw = W / 4, cw = W / 2;
  int mb, dir, e, k, i;
  for (mb = 0; mb < mbw * mbh; mb++) {
    const int mbx = mb % mbw, mby = mb / mbw;
    const OrcDbkMb *mq = &mbs[mb];
    if (mq->disable) continue;
    for (dir = 0; dir < 2; dir++)                                   /* vertical edges first, then horizontal */
      for (e = 0; e < 4; e++) {
        const OrcDbkMb *mp = mq;
        int bs[4], any = 0;
        if (e == 0) { if ((dir ? mby : mbx) == 0) continue; mp = &mbs[dir ? mb - mbw : mb - 1]; }
        for (k = 0; k < 4; k++) {                                   /* the four 4-sample segments of the edge */
          const int qx = dir ? k : e, qy = dir ? e : k;             /* 4x4 block q inside the macroblock */
          const int px = dir ? qx : (qx + 3) & 3, py = dir ? (qy + 3) & 3 : qy;
          const OrcDbkBlk *bq = &blks[(size_t)(mby * 4 + qy) * bw + mbx * 4 + qx];
          const OrcDbkBlk *bp = dir ? bq - bw : bq - 1;
          bs[k] = dbk_strength(mp, mq, (mp->cbp_blk >> (py * 4 + px)) & 1, (mq->cbp_blk >> (qy * 4 + qx)) & 1, bp, bq, e == 0);
          any |= bs[k];
        }
        if (!any) continue;
        if (!((e & 1) && mq->transform8x8)) {                       /* luma (edges 1, 3 of an 8x8-transform macroblock are not edges) */
          const int qp = (mp->qp + mq->qp + 1) >> 1, ia = dbk_clip3(0, 51, qp + mq->alpha_off), ib = dbk_clip3(0, 51, qp + mq->beta_off);
          const int alpha = DBK_ALPHA[ia], beta = DBK_BETA[ib];
          if (alpha | beta)
            for (i = 0; i < 16; i++) {
              uint8_t *s = dir ? y + (size_t)(mby * 16 + e * 4) * W + mbx * 16 + i : y + (size_t)(mby * 16 + i) * W + mbx * 16 + e * 4;
              if (bs[i >> 2]) dbk_luma_line(s, dir ? W : 1, bs[i >> 2], alpha, beta, DBK_CLIP[ia][bs[i >> 2]]);
            }
        }
        if (!(e & 1)) {                                             /* chroma edges 0 and 4 <-> luma edges 0 and 2 */
          int pl;
          for (pl = 0; pl < 2; pl++) {
            uint8_t *c = pl ? v : u;
            const int qp = ((pl ? mp->qpc_v : mp->qpc_u) + (pl ? mq->qpc_v : mq->qpc_u) + 1) >> 1;
            const int ia = dbk_clip3(0, 51, qp + mq->alpha_off), ib = dbk_clip3(0, 51, qp + mq->beta_off);
            const int alpha = DBK_ALPHA[ia], beta = DBK_BETA[ib];
            if (!(alpha | beta)) continue;
            for (i = 0; i < 8; i++) {
              uint8_t *s = dir ? c + (size_t)(mby * 8 + e * 2) * cw + mbx * 8 + i : c + (size_t)(mby * 8 + i) * cw + mbx * 8 + e * 2;
              if (bs[i >> 1]) dbk_chroma_line(s, dir ? cw : 1, bs[i >> 1], alpha, beta, DBK_CLIP[ia][bs[i >> 1]]);
            }
          }
        }
      }
  }
}
