/* b2_oracle_tq.c -- TEST INFRASTRUCTURE ONLY: CPU restatement of the residual transform + quant path.
 *   orc_forward4x4 / orc_inverse4x4   <- JM/lcommon/src/transform.c:20-67 / 70-118
 *   orc_forward8x8 / orc_inverse8x8   <- JM/lcommon/src/transform.c:353-448 / 450-548
 *   orc_tq4x4 (mode 0)  <- residual_transform_quant_luma_4x4 JM/lencod/src/block.c:660-724 with
 *                          quant_4x4_normal JM/lencod/src/quant4x4_normal.c:39-115, check_zero block.c:626,
 *                          sample_reconstruct JM/lcommon/src/blk_prediction.c:48
 *   orc_tq4x4 (mode 1)  <- dct_luma V1/src/block.c:836-1045 (restated only: block.c of version1 does not
 *                          compile under gcc as shipped; its differences to JM are the
 *                          rounding offset, the missing level clip and the missing zero shortcuts, Q-F7)
 *   orc_tq8x8           <- residual_transform_quant_luma_8x8 JM/lencod/src/transform8x8.c:522-602 with
 *                          quant_8x8_normal JM/lencod/src/quant8x8_normal.c:43
 * Pinned against the unmodified JM objects (oracle/jm_harness_tq.c) by tests/test_oracle_tq.py and
 * tests/golden/jm_tq.npz (oracle/gen_golden_tq.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {          /* same layout as b2tq_params (include/b2me.h) */
  int32_t qp, mode, cavlc, field_scan, disthres, reserved[3];
  int32_t scale[64], offset[64], invscale[64];
} orc_tq_params;

static const uint8_t SNGL4[16][2] = {{0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{1,3},{2,2},{3,1},{3,2},{2,3},{3,3}};
static const uint8_t FIELD4[16][2] = {{0,0},{0,1},{1,0},{0,2},{0,3},{1,1},{1,2},{1,3},{2,0},{2,1},{2,2},{2,3},{3,0},{3,1},{3,2},{3,3}};
static const uint8_t COST4[2][16] = {{3,2,2,1,1,1,0,0,0,0,0,0,0,0,0,0},{9,9,9,9,9,9,9,9,9,9,9,9,9,9,9,9}};
static uint8_t SNGL8[64][2], FIELD8[64][2];
static int scans_ready;
static void build_scans(void)
{
  /* 8x8 zig-zag generated from its definition (anti-diagonals, alternating direction); the 8x8
   * field scan is the listed table of H.264 Table 8-13 as raster index 8*j + i. */
  static const uint8_t f8[64] = {0,8,16,1,9,24,32,17,2,25,40,48,56,33,10,3,18,41,49,57,26,11,4,19,34,42,50,58,27,12,5,20,
                                 35,43,51,59,28,13,6,21,36,44,52,60,29,14,22,37,45,53,61,30,7,15,38,46,54,62,23,31,39,47,55,63};
  int d, k = 0, i;
  for (d = 0; d < 15; d++) {
    int lo = d < 8 ? 0 : d - 7, hi = d < 8 ? d : 7, t;
    for (t = lo; t <= hi; t++) {
      int a = (d & 1) ? hi - (t - lo) : t;     /* a = i (horizontal); odd diagonals start at the right */
      SNGL8[k][0] = (uint8_t)a; SNGL8[k][1] = (uint8_t)(d - a); k++;
    }
  }
  for (i = 0; i < 64; i++) { FIELD8[i][0] = (uint8_t)(f8[i] & 7); FIELD8[i][1] = (uint8_t)(f8[i] >> 3); }
  scans_ready = 1;
}
static int cost8(int run, int dis) { return dis ? 9 : (run < 4 ? 3 : run < 12 ? 2 : run < 24 ? 1 : 0); }

void orc_forward4x4(const int *in, int *out)
{
  int tmp[16], i;
  for (i = 0; i < 4; i++) {
    int p0 = in[4*i], p1 = in[4*i+1], p2 = in[4*i+2], p3 = in[4*i+3];
    int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3;
    tmp[4*i] = t0 + t1; tmp[4*i+1] = (t3 << 1) + t2; tmp[4*i+2] = t0 - t1; tmp[4*i+3] = t3 - (t2 << 1);
  }
  for (i = 0; i < 4; i++) {
    int p0 = tmp[i], p1 = tmp[4+i], p2 = tmp[8+i], p3 = tmp[12+i];
    int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3;
    out[i] = t0 + t1; out[4+i] = t2 + (t3 << 1); out[8+i] = t0 - t1; out[12+i] = t3 - (t2 << 1);
  }
}
void orc_inverse4x4(const int *in, int *out)
{
  int tmp[16], i;
  for (i = 0; i < 4; i++) {
    int t0 = in[4*i], t1 = in[4*i+1], t2 = in[4*i+2], t3 = in[4*i+3];
    int p0 = t0 + t2, p1 = t0 - t2, p2 = (t1 >> 1) - t3, p3 = t1 + (t3 >> 1);
    tmp[4*i] = p0 + p3; tmp[4*i+1] = p1 + p2; tmp[4*i+2] = p1 - p2; tmp[4*i+3] = p0 - p3;
  }
  for (i = 0; i < 4; i++) {
    int t0 = tmp[i], t1 = tmp[4+i], t2 = tmp[8+i], t3 = tmp[12+i];
    int p0 = t0 + t2, p1 = t0 - t2, p2 = (t1 >> 1) - t3, p3 = t1 + (t3 >> 1);
    out[i] = p0 + p3; out[4+i] = p1 + p2; out[8+i] = p1 - p2; out[12+i] = p0 - p3;
  }
}
static void f8_1d(const int *p, int s, int *o, int so)
{
  int a0 = p[0] + p[7*s], a1 = p[s] + p[6*s], a2 = p[2*s] + p[5*s], a3 = p[3*s] + p[4*s];
  int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2, b4, b5, b6, b7;
  a0 = p[0] - p[7*s]; a1 = p[s] - p[6*s]; a2 = p[2*s] - p[5*s]; a3 = p[3*s] - p[4*s];
  b4 = a1 + a2 + ((a0 >> 1) + a0); b5 = a0 - a3 - ((a2 >> 1) + a2);
  b6 = a0 + a3 - ((a1 >> 1) + a1); b7 = a1 - a2 + ((a3 >> 1) + a3);
  o[0] = b0 + b1; o[so] = b4 + (b7 >> 2); o[2*so] = b2 + (b3 >> 1); o[3*so] = b5 + (b6 >> 2);
  o[4*so] = b0 - b1; o[5*so] = b6 - (b5 >> 2); o[6*so] = (b2 >> 1) - b3; o[7*so] = (b4 >> 2) - b7;
}
static void i8_1d(const int *p, int s, int *o, int so)
{
  int a0 = p[0] + p[4*s], a1 = p[0] - p[4*s], a2 = p[6*s] - (p[2*s] >> 1), a3 = p[2*s] + (p[6*s] >> 1);
  int b0 = a0 + a3, b2 = a1 - a2, b4 = a1 + a2, b6 = a0 - a3, b1, b3, b5, b7;
  a0 = -p[3*s] + p[5*s] - p[7*s] - (p[7*s] >> 1); a1 = p[s] + p[7*s] - p[3*s] - (p[3*s] >> 1);
  a2 = -p[s] + p[7*s] + p[5*s] + (p[5*s] >> 1);   a3 = p[3*s] + p[5*s] + p[s] + (p[s] >> 1);
  b1 = a0 + (a3 >> 2); b3 = a1 + (a2 >> 2); b5 = a2 - (a1 >> 2); b7 = a3 - (a0 >> 2);
  o[0] = b0 + b7; o[so] = b2 - b5; o[2*so] = b4 + b3; o[3*so] = b6 + b1;
  o[4*so] = b6 - b1; o[5*so] = b4 - b3; o[6*so] = b2 + b5; o[7*so] = b0 - b7;
}
void orc_forward8x8(const int *in, int *out)
{
  int tmp[64], i;
  for (i = 0; i < 8; i++) f8_1d(in + 8 * i, 1, tmp + 8 * i, 1);
  for (i = 0; i < 8; i++) f8_1d(tmp + i, 8, out + i, 8);
}
void orc_inverse8x8(const int *in, int *out)
{
  int tmp[64], i;
  for (i = 0; i < 8; i++) i8_1d(in + 8 * i, 1, tmp + 8 * i, 1);
  for (i = 0; i < 8; i++) i8_1d(tmp + i, 8, out + i, 8);
}
static int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

void orc_tq4x4(const orc_tq_params *P, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level, uint8_t *run,
               uint8_t *recon, int32_t *cost, uint8_t *nonzero)
{
  int k, i;
  const int qp_per = P->qp / 6, q_bits = 15 + qp_per;
  const uint8_t (*scan)[2] = P->field_scan ? FIELD4 : SNGL4;
  for (k = 0; k < nblk; k++) {
    int res[16], t[16], r[16], any = 0, nz = 0, c = 0, n = 0, runc = 0, s;
    memset(level + 16 * k, 0, 32); memset(run + 16 * k, 0, 16);
    for (i = 0; i < 16; i++) { res[i] = (int)orig[16*k+i] - (int)pred[16*k+i]; any |= res[i]; }
    if (any != 0 || P->mode == 1) {
      orc_forward4x4(res, t);
      for (s = 0; s < 16; s++) {
        int idx = scan[s][1] * 4 + scan[s][0], m7 = t[idx], lv = 0;
        if (m7 != 0 || P->mode == 1) lv = (abs(m7) * P->scale[idx] + P->offset[idx]) >> q_bits;
        if (lv != 0) {
          int sl;
          if (P->cavlc && P->mode == 0 && lv > 2063) lv = 2063;
          c += (lv > 1) ? 999999 : COST4[P->disthres][runc];
          sl = m7 < 0 ? -lv : lv;
          if (P->mode == 0) t[idx] = (((sl * P->invscale[idx]) << qp_per) + 8) >> 4;
          else { int il = (lv * P->invscale[idx]) << qp_per; t[idx] = m7 < 0 ? -il : il; }
          level[16*k+n] = (int16_t)sl; run[16*k+n] = (uint8_t)runc; n++; runc = 0; nz = 1;
        } else { t[idx] = 0; runc++; }
      }
    }
    if (nz || P->mode == 1) {
      orc_inverse4x4(t, r);
      for (i = 0; i < 16; i++) recon[16*k+i] = (uint8_t)clip255(((r[i] + 32) >> 6) + pred[16*k+i]);
    } else memcpy(recon + 16 * k, pred + 16 * k, 16);
    cost[k] = c; nonzero[k] = (uint8_t)nz;
  }
}

void orc_tq8x8(const orc_tq_params *P, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level, uint8_t *run,
               uint8_t *recon, int32_t *cost, uint8_t *nonzero)
{
  int k, i;
  const int qp_per = P->qp / 6, q_bits = 16 + qp_per;
  const uint8_t (*scan)[2];
  if (!scans_ready) build_scans();
  scan = P->field_scan ? FIELD8 : SNGL8;
  for (k = 0; k < nblk; k++) {
    int res[64], t[64], r[64], any = 0, nz = 0, c = 0, n = 0, runc = 0, s;
    memset(level + 64 * k, 0, 128); memset(run + 64 * k, 0, 64);
    for (i = 0; i < 64; i++) { res[i] = (int)orig[64*k+i] - (int)pred[64*k+i]; any |= res[i]; }
    if (any != 0) {
      orc_forward8x8(res, t);
      for (s = 0; s < 64; s++) {
        int idx = scan[s][1] * 8 + scan[s][0], m7 = t[idx], lv = 0;
        if (m7 != 0) lv = (abs(m7) * P->scale[idx] + P->offset[idx]) >> q_bits;
        if (lv != 0) {
          int sl = m7 < 0 ? -lv : lv;
          c += (lv > 1) ? 999999 : cost8(runc, P->disthres);
          t[idx] = (((sl * P->invscale[idx]) << qp_per) + 32) >> 6;
          level[64*k+n] = (int16_t)sl; run[64*k+n] = (uint8_t)runc; n++; runc = 0; nz = 1;
        } else { t[idx] = 0; runc++; }
      }
    }
    if (nz) {
      orc_inverse8x8(t, r);
      for (i = 0; i < 64; i++) recon[64*k+i] = (uint8_t)clip255(((r[i] + 32) >> 6) + pred[64*k+i]);
    } else memcpy(recon + 64 * k, pred + 64 * k, 64);
    cost[k] = c; nonzero[k] = (uint8_t)nz;
  }
}
void orc_scan8(int field, uint8_t *out) { if (!scans_ready) build_scans(); memcpy(out, field ? FIELD8 : SNGL8, 128); }

/* ---- Intra16x16 luma (restated): residual_transform_quant_luma_16x16 JM/lencod/src/block.c:207-345 with hadamard4x4 /
 * ihadamard4x4 (JM/lcommon/src/transform.c:121-214), quant_dc4x4_normal (quant4x4_normal.c:200-270: q_bits + 1, offset << 1,
 * the level itself goes back into the block) and quant_ac4x4_normal (:117-190: scan positions 1..15).  P = the intra 4x4
 * table of the plane.  orig / pred / recon [nmb][256] raster; dc lists [nmb][16]; ac lists [nmb][16 blocks raster][16]. */
void orc_hadamard4x4_dc(const int *in, int *out)
{
  int tmp[16], i;
  for (i = 0; i < 4; i++) {
    int p0 = in[4*i], p1 = in[4*i+1], p2 = in[4*i+2], p3 = in[4*i+3];
    int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3;
    tmp[4*i] = t0 + t1; tmp[4*i+1] = t3 + t2; tmp[4*i+2] = t0 - t1; tmp[4*i+3] = t3 - t2;
  }
  for (i = 0; i < 4; i++) {
    int p0 = tmp[i], p1 = tmp[4+i], p2 = tmp[8+i], p3 = tmp[12+i];
    int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3;
    out[i] = (t0 + t1) >> 1; out[4+i] = (t2 + t3) >> 1; out[8+i] = (t0 - t1) >> 1; out[12+i] = (t3 - t2) >> 1;
  }
}
void orc_ihadamard4x4_dc(const int *in, int *out)
{
  int tmp[16], i;
  for (i = 0; i < 4; i++) {
    int t0 = in[4*i], t1 = in[4*i+1], t2 = in[4*i+2], t3 = in[4*i+3];
    int p0 = t0 + t2, p1 = t0 - t2, p2 = t1 - t3, p3 = t1 + t3;
    tmp[4*i] = p0 + p3; tmp[4*i+1] = p1 + p2; tmp[4*i+2] = p1 - p2; tmp[4*i+3] = p0 - p3;
  }
  for (i = 0; i < 4; i++) {
    int t0 = tmp[i], t1 = tmp[4+i], t2 = tmp[8+i], t3 = tmp[12+i];
    int p0 = t0 + t2, p1 = t0 - t2, p2 = t1 - t3, p3 = t1 + t3;
    out[i] = p0 + p3; out[4+i] = p1 + p2; out[8+i] = p1 - p2; out[12+i] = p0 - p3;
  }
}
void orc_tq16x16(const orc_tq_params *P, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                 int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *ac_coef)
{
  int k, b, i, j, s;
  const int qp_per = P->qp / 6, q_bits = 15 + qp_per;
  const uint8_t (*scan)[2] = P->field_scan ? FIELD4 : SNGL4;
  for (k = 0; k < nmb; k++) {
    int t[16][16], dc[16], hd[16], n = 0, runc = 0, nzdc = 0, ac = 0;
    memset(dc_level + 16 * k, 0, 32); memset(dc_run + 16 * k, 0, 16);
    memset(ac_level + 256 * k, 0, 512); memset(ac_run + 256 * k, 0, 256);
    for (b = 0; b < 16; b++) {
      int res[16];
      for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) { int o = 256 * k + (4 * (b >> 2) + j) * 16 + 4 * (b & 3) + i; res[4 * j + i] = (int)orig[o] - (int)pred[o]; }
      orc_forward4x4(res, t[b]);
      dc[b] = t[b][0];
    }
    orc_hadamard4x4_dc(dc, hd);
    for (s = 0; s < 16; s++) {                       /* quant_dc4x4_normal */
      int idx = scan[s][1] * 4 + scan[s][0], m7 = hd[idx];
      if (m7 != 0) {
        int lv = (abs(m7) * P->scale[0] + (P->offset[0] << 1)) >> (q_bits + 1);
        if (lv != 0) {
          if (P->cavlc && lv > 2063) lv = 2063;
          lv = m7 < 0 ? -lv : lv;
          hd[idx] = lv; dc_level[16 * k + n] = (int16_t)lv; dc_run[16 * k + n] = (uint8_t)runc; n++; runc = 0; nzdc = 1;
        } else { runc++; hd[idx] = 0; }
      } else runc++;
    }
    if (nzdc) {
      orc_ihadamard4x4_dc(hd, dc);
      for (b = 0; b < 16; b++) t[b][0] = (((dc[b] * P->invscale[0]) << qp_per) + 32) >> 6;
    } else for (b = 0; b < 16; b++) t[b][0] = 0;
    for (b = 0; b < 16; b++) {                       /* quant_ac4x4_normal + inverse4x4 */
      int r[16], nz = 0;
      n = 0; runc = 0;
      for (s = 1; s < 16; s++) {
        int idx = scan[s][1] * 4 + scan[s][0], m7 = t[b][idx];
        if (m7 != 0) {
          int lv = (abs(m7) * P->scale[idx] + P->offset[idx]) >> q_bits;
          if (lv != 0) {
            if (P->cavlc && lv > 2063) lv = 2063;
            lv = m7 < 0 ? -lv : lv;
            t[b][idx] = (((lv * P->invscale[idx]) << qp_per) + 8) >> 4;
            ac_level[(16 * k + b) * 16 + n] = (int16_t)lv; ac_run[(16 * k + b) * 16 + n] = (uint8_t)runc; n++; runc = 0; nz = 1;
          } else { t[b][idx] = 0; runc++; }
        } else runc++;
      }
      if (nz) ac = 15;
      if (t[b][0] != 0 || nz) orc_inverse4x4(t[b], r); else memcpy(r, t[b], sizeof(r));
      for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) {
        int o = 256 * k + (4 * (b >> 2) + j) * 16 + 4 * (b & 3) + i;
        recon[o] = (uint8_t)clip255(((r[4 * j + i] + 32) >> 6) + pred[o]);
      }
    }
    ac_coef[k] = (uint8_t)ac;
  }
}

/* ---- chroma of a 4:2:0 macroblock, one plane: residual_transform_quant_chroma_4x4 (JM/lencod/src/block.c:953-1200) -----------
 * forward4x4 of the four 4x4 blocks, hadamard2x2 of their DC coefficients (lcommon/src/transform.c:302-315), quant_dc2x2_normal
 * (quantChroma_normal.c:37-96: natural order, offset << 1, q_bits + 1, levels de-quantised in place), ihadamard2x2 (:317-331)
 * and >> 5, quant_ac4x4_normal per block (quant4x4_normal.c:117-190) with its coefficient cost, the rule that throws ALL AC
 * levels away when that cost stays below _CHROMA_COEFF_COST_ = 4 (block.c:1137-1168), inverse4x4 of the blocks that hold
 * anything, sample_reconstruct.  P->qp = the chroma qp (qpc[uv]), P's tables = the chroma plane's.
 * dc_level / dc_run [nmb][4]; ac_level / ac_run [nmb][4][16] (blocks in raster order); cr_cbp: 0 nothing, 1 DC only, 2 AC. */
void orc_tq_chroma(const orc_tq_params *P, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                   int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *cr_cbp)
{
  int k, b, i, j, s;
  const int qp_per = P->qp / 6, q_bits = 15 + qp_per;
  const uint8_t (*scan)[2] = P->field_scan ? FIELD4 : SNGL4;
  for (k = 0; k < nmb; k++) {
    int t[4][16], m[4], d[4], n = 0, runc = 0, dcnz = 0, cost = 0, nzb[4] = {0, 0, 0, 0}, anynz = 0, cbp = 0;
    memset(dc_level + 4 * k, 0, 8); memset(dc_run + 4 * k, 0, 4);
    memset(ac_level + 64 * k, 0, 128); memset(ac_run + 64 * k, 0, 64);
    for (b = 0; b < 4; b++) {
      int res[16];
      for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) { int o = 64 * k + (4 * (b >> 1) + j) * 8 + 4 * (b & 1) + i; res[4 * j + i] = (int)orig[o] - (int)pred[o]; }
      orc_forward4x4(res, t[b]);
    }
    m[0] = t[0][0] + t[1][0] + t[2][0] + t[3][0]; m[1] = t[0][0] - t[1][0] + t[2][0] - t[3][0];
    m[2] = t[0][0] + t[1][0] - t[2][0] - t[3][0]; m[3] = t[0][0] - t[1][0] - t[2][0] + t[3][0];
    for (s = 0; s < 4; s++) {
      if (m[s] != 0) {
        int lv = (abs(m[s]) * P->scale[0] + (P->offset[0] << 1)) >> (q_bits + 1);
        if (lv != 0) {
          if (P->cavlc && lv > 2063) lv = 2063;
          lv = m[s] < 0 ? -lv : lv;
          m[s] = (lv * P->invscale[0]) << qp_per;
          dc_level[4 * k + n] = (int16_t)lv; dc_run[4 * k + n] = (uint8_t)runc; n++; runc = 0; dcnz = 1;
        } else { runc++; m[s] = 0; }
      } else runc++;
    }
    if (dcnz) cbp = 1;
    d[0] = m[0] + m[1] + m[2] + m[3]; d[1] = m[0] - m[1] + m[2] - m[3]; d[2] = m[0] + m[1] - m[2] - m[3]; d[3] = m[0] - m[1] - m[2] + m[3];
    for (b = 0; b < 4; b++) t[b][0] = d[b] >> 5;
    for (b = 0; b < 4; b++) {
      n = 0; runc = 0;
      for (s = 1; s < 16; s++) {
        int idx = scan[s][1] * 4 + scan[s][0], m7 = t[b][idx];
        if (m7 != 0) {
          int lv = (abs(m7) * P->scale[idx] + P->offset[idx]) >> q_bits;
          if (lv != 0) {
            if (P->cavlc && lv > 2063) lv = 2063;
            cost += (lv > 1) ? 999999 : COST4[P->disthres][runc];
            lv = m7 < 0 ? -lv : lv;
            t[b][idx] = (((lv * P->invscale[idx]) << qp_per) + 8) >> 4;
            ac_level[(4 * k + b) * 16 + n] = (int16_t)lv; ac_run[(4 * k + b) * 16 + n] = (uint8_t)runc; n++; runc = 0; nzb[b] = 1;
          } else { t[b][idx] = 0; runc++; }
        } else runc++;
      }
      if (nzb[b]) anynz = 1;
    }
    if (anynz && cost < 4) {                         /* _CHROMA_COEFF_COST_: too few, too small AC levels -- drop them all */
      for (b = 0; b < 4; b++)
        if (nzb[b]) {
          nzb[b] = 0;
          for (s = 1; s < 16; s++) t[b][s] = 0;
          memset(ac_level + (4 * k + b) * 16, 0, 32);       /* ACRun keeps its values, as in the reference (only the levels are cleared) */
        }
    } else if (anynz) cbp = 2;
    for (b = 0; b < 4; b++) {
      int r[16];
      if (t[b][0] != 0 || nzb[b]) orc_inverse4x4(t[b], r); else memcpy(r, t[b], sizeof(r));
      for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) {
        int o = 64 * k + (4 * (b >> 1) + j) * 8 + 4 * (b & 1) + i;
        recon[o] = (uint8_t)clip255(((r[4 * j + i] + 32) >> 6) + pred[o]);
      }
    }
    cr_cbp[k] = (uint8_t)cbp;
  }
}
