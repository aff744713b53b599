// deblock.cu -- in-loop deblocking filter of a frame picture on the device.
//
// Replaces DeblockFrame (JM/lencod/src/loopFilter.c:63-111; DeblockMb :196-377) with the non-MBAFF functions of
// JM/lencod/src/loop_filter_normal.c: GetStrengthVer / GetStrengthHor (:52-283), EdgeLoopLumaVer / Hor (:285-585),
// EdgeLoopChromaVer / Hor (:590-758); alpha / beta / clip tables: JM/lencod/inc/loop_filter.h:34-46 (H.264 tables 8-16, 8-17).
//
// The filter is sequential in macroblock order: macroblock (x, y) modifies three sample columns of (x - 1, y) and three rows of
// (x, y - 1), and (x + 1, y - 1) modifies three columns of (x, y - 1) -- including the 3 x 3 corner that (x, y)'s top edge touches
// afterwards.  So the top edge of (x, y) may run once (x - 1, y) is done and the LEFT edge of (x + 1, y - 1) has been filtered: the
// 2:1 wavefront the reference itself describes under JM_PARALLEL_DEBLOCK (loopFilter.c:92-109), one left edge tighter.
//   k_dbk_prep   everything that does not depend on samples, for the whole picture at once (one warp per macroblock): the 32
//                boundary strengths, alpha / beta / clip values per (neighbour kind, plane) -> one 96-byte DbkRec per macroblock
//   k_deblock    the wavefront: one CTA per macroblock ROW, warp 0 the luma plane, warp 1 both chroma planes (independent chains,
//                each with its own half of the row's progress counter), warps 2 / 3 the luma chain's fetcher / publisher; all rows are
//                resident at once, so the waits cannot deadlock.
// Per macroblock the serial chain holds sample work only: vertical edges on the staged tile (a lane keeps one row through all four
// edges; the left neighbour's four columns stay in the tile from the previous step; own samples and record were fetched one step
// ahead) -> after the left edge the neighbour's last columns are written and the row's counter released -> wait for the row above ->
// its four rows over this macroblock (one L2 round trip) -> horizontal edges (a lane keeps one column) -> write-back.
// Sample loads bypass L1 (ld.global.cg): a line fetched for one macroblock also holds samples that the row above rewrites later.
// The luma chain's two global round trips run on helper warps of the row's CTA: warp 2 waits for the row above and stages its last
// four rows (s_top), warp 3 writes the left neighbour's final columns and releases the counter (s_left); the luma warp's own chain per
// macroblock holds neither a fence nor a global wait.  The luma edges run on registers (a row through the four vertical edges, a
// column through the four horizontal ones).
// Measured (1080p, B200): 2.55 ms for the one-warp-per-row kernel of the first version; 1.31 ms with the pre-pass / two warps / early
// release (per macroblock of a middle row, -DDBK_PROF: vertical edges 2 700 cycles + 1 020 for the write + release, wait 4 700, rows
// above 770, horizontal edges 2 480); 1.01 ms with the edges on registers; 0.72 ms with the helper warps.
#include <cstdio>
#include "b2_common.cuh"
#include "../../include/b2me.h"
#include "b2_ctx.h"

namespace b2 {

__constant__ uint8_t c_dbk_alpha[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
__constant__ uint8_t c_dbk_beta[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
__constant__ uint8_t c_dbk_clip[52][5] = {
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,1,1,1},{0,0,1,1,1},{0,1,1,1,1},
  {0,1,1,1,1},{0,1,1,1,1},{0,1,1,1,1},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,2,3,3},{0,1,2,3,3},{0,2,2,3,3},{0,2,2,4,4},{0,2,3,4,4},
  {0,2,3,4,4},{0,3,3,5,5},{0,3,4,6,6},{0,3,4,6,6},{0,4,5,7,7},{0,4,5,8,8},{0,4,6,9,9},{0,5,7,10,10},{0,6,8,11,11},{0,6,8,13,13},{0,7,10,14,14},{0,8,11,16,16},
  {0,9,12,18,18},{0,10,13,20,20},{0,11,15,23,23},{0,13,17,25,25}};

struct DbkArgs {
  int W, H, mbw, mbh;
  uint8_t *y; int yp; uint8_t *u, *v; int cp;
  const b2dbk_mb *mbs; const b2dbk_blk *blks; int *progress;     // progress: one counter per macroblock row, DBK_PSTRIDE ints apart
};

__device__ __forceinline__ int dbk_mvne(const b2dbk_blk &a, int la, const b2dbk_blk &b, int lb)
{ return (abs(a.mv[la][0] - b.mv[lb][0]) >= 4) | (abs(a.mv[la][1] - b.mv[lb][1]) >= 4); }

// one line of samples across an edge in the staged tile: q[-k * st] = p(k - 1), q[k * st] = q(k)
__device__ __forceinline__ void dbk_luma_line(uint8_t *q, int st, int bs, int alpha, int beta, int c0)
{
  const int p0 = q[-st], p1 = q[-2 * st], p2 = q[-3 * st], q0 = q[0], q1 = q[st], q2 = q[2 * st];
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) {
    const int small_gap = abs(q0 - p0) < ((alpha >> 2) + 2);
    const int ap = (abs(p0 - p2) < beta) & small_gap, aq = (abs(q0 - q2) < beta) & small_gap, s = p0 + q0;
    if (ap) {
      const int p3 = q[-4 * st];
      q[-st] = (uint8_t)((q1 + ((p1 + s) << 1) + p2 + 4) >> 3); q[-2 * st] = (uint8_t)((p2 + p1 + s + 2) >> 2); q[-3 * st] = (uint8_t)((((p3 + p2) << 1) + p2 + p1 + s + 4) >> 3);
    } else q[-st] = (uint8_t)(((p1 << 1) + p0 + q1 + 2) >> 2);
    if (aq) {
      const int q3 = q[3 * st];
      q[0] = (uint8_t)((p1 + ((q1 + s) << 1) + q2 + 4) >> 3); q[st] = (uint8_t)((q2 + q0 + p0 + q1 + 2) >> 2); q[2 * st] = (uint8_t)((((q3 + q2) << 1) + q2 + q1 + s + 4) >> 3);
    } else q[0] = (uint8_t)(((q1 << 1) + q0 + p1 + 2) >> 2);
  } else {
    const int avg = (p0 + q0 + 1) >> 1, ap = abs(p0 - p2) < beta, aq = abs(q0 - q2) < beta, tc = c0 + ap + aq;
    const int dif = iclamp((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    if (ap) q[-2 * st] = (uint8_t)(p1 + iclamp((p2 + avg - (p1 << 1)) >> 1, -c0, c0));
    if (dif) { q[-st] = (uint8_t)iclamp(p0 + dif, 0, 255); q[0] = (uint8_t)iclamp(q0 - dif, 0, 255); }
    if (aq) q[st] = (uint8_t)(q1 + iclamp((q2 + avg - (q1 << 1)) >> 1, -c0, c0));
  }
}
// the same filter on a line held in registers (p3 .. q3 by reference; p3 / q3 are only read)
__device__ __forceinline__ void dbk_luma_regs(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3, int bs, int alpha, int beta, int c0)
{
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) {
    const int small_gap = abs(q0 - p0) < ((alpha >> 2) + 2);
    const int ap = (abs(p0 - p2) < beta) & small_gap, aq = (abs(q0 - q2) < beta) & small_gap, s = p0 + q0;
    const int P0 = p0, P1 = p1, P2 = p2, Q0 = q0, Q1 = q1, Q2 = q2;
    if (ap) { p0 = (Q1 + ((P1 + s) << 1) + P2 + 4) >> 3; p1 = (P2 + P1 + s + 2) >> 2; p2 = (((p3 + P2) << 1) + P2 + P1 + s + 4) >> 3; }
    else p0 = ((P1 << 1) + P0 + Q1 + 2) >> 2;
    if (aq) { q0 = (P1 + ((Q1 + s) << 1) + Q2 + 4) >> 3; q1 = (Q2 + Q0 + P0 + Q1 + 2) >> 2; q2 = (((q3 + Q2) << 1) + Q2 + Q1 + s + 4) >> 3; }
    else q0 = ((Q1 << 1) + Q0 + P1 + 2) >> 2;
  } else {
    const int avg = (p0 + q0 + 1) >> 1, ap = abs(p0 - p2) < beta, aq = abs(q0 - q2) < beta, tc = c0 + ap + aq;
    const int dif = iclamp((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    const int P1 = p1, Q1 = q1;
    if (ap) p1 = P1 + iclamp((p2 + avg - (P1 << 1)) >> 1, -c0, c0);
    if (aq) q1 = Q1 + iclamp((q2 + avg - (Q1 << 1)) >> 1, -c0, c0);
    if (dif) { p0 = iclamp(p0 + dif, 0, 255); q0 = iclamp(q0 - dif, 0, 255); }
  }
}
__device__ __forceinline__ void dbk_chroma_line(uint8_t *q, int st, int bs, int alpha, int beta, int c0)
{
  const int p0 = q[-st], p1 = q[-2 * st], q0 = q[0], q1 = q[st];
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) { q[-st] = (uint8_t)(((p1 << 1) + p0 + q1 + 2) >> 2); q[0] = (uint8_t)(((q1 << 1) + q0 + p1 + 2) >> 2); }
  else {
    const int tc = c0 + 1, dif = iclamp((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    if (dif) { q[-st] = (uint8_t)iclamp(p0 + dif, 0, 255); q[0] = (uint8_t)iclamp(q0 - dif, 0, 255); }
  }
}

__device__ __forceinline__ uint32_t ld_cg32(const uint8_t *p) { return __ldcg(reinterpret_cast<const uint32_t *>(p)); }
__device__ __forceinline__ int ld_acquire(const int *p)
{ int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_relaxed(const int *p)
{ int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_add(int *p, int v)
{ asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Everything of a macroblock's filter that does not depend on samples, computed for the whole picture at once (k_dbk_prep, one
// warp per macroblock) so that the wavefront kernel's serial chain holds sample work only: the 32 boundary strengths
// (GetStrengthVer / GetStrengthHor) and alpha / beta / the clip-table row per (neighbour kind, plane).
struct __align__(16) DbkRec {
  uint8_t bs[2][4][4];                         // [direction][edge][segment]
  uint8_t alpha[3][3], beta[3][3];             // [0 left edge, 1 top edge, 2 internal edges][Y, U, V]
  uint8_t c0[3][3][4];                         // clip-table entry per strength (index bs: 1..3 used; bs == 4 takes the strong filter)
  uint8_t t8, pad_[9];
};
constexpr int DBK_RECW = 24;                   // words per record
static_assert(sizeof(DbkRec) == 4 * DBK_RECW, "DbkRec is one 96-byte record per macroblock");

__global__ void __launch_bounds__(128) k_dbk_prep(const DbkArgs a, DbkRec *rec)
{
  const int mb = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (mb >= a.mbw * a.mbh) return;
  const int row = mb / a.mbw, mbx = mb - row * a.mbw, bw = a.W >> 2;
  const b2dbk_mb mq = a.mbs[mb];
  const b2dbk_mb ml = mbx ? a.mbs[mb - 1] : mq, mt = row ? a.mbs[mb - a.mbw] : mq;
  {
    const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
    const int qx = dir ? k : e, qy = dir ? e : k, px = dir ? qx : (qx + 3) & 3, py = dir ? (qy + 3) & 3 : qy;
    int s = 0;
    if (!mq.disable && !(e == 0 && (dir ? row : mbx) == 0)) {
      const b2dbk_mb &mp = e ? mq : (dir ? mt : ml);
      if (mp.intra || mq.intra) s = e ? 3 : 4;
      else if (((mp.cbp_blk >> (py * 4 + px)) & 1) || ((mq.cbp_blk >> (qy * 4 + qx)) & 1)) s = 2;
      else {
        const b2dbk_blk *pq = a.blks + (size_t)(row * 4 + qy) * bw + mbx * 4 + qx;
        const b2dbk_blk bq = *pq, bp = *(dir ? pq - bw : pq - 1);
        const int p0 = bp.ref[0], p1 = bp.ref[1], q0 = bq.ref[0], q1 = bq.ref[1];
        if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) s = 1;
        else if (p0 != p1) s = p0 == q0 ? (dbk_mvne(bp, 0, bq, 0) | dbk_mvne(bp, 1, bq, 1)) : (dbk_mvne(bp, 0, bq, 1) | dbk_mvne(bp, 1, bq, 0));
        else s = (dbk_mvne(bp, 0, bq, 0) | dbk_mvne(bp, 1, bq, 1)) && (dbk_mvne(bp, 0, bq, 1) | dbk_mvne(bp, 1, bq, 0));
      }
    }
    rec[mb].bs[dir][e][k] = (uint8_t)s;
  }
  if (lane < 9) {
    const int t = lane / 3, pl = lane - 3 * t;
    const b2dbk_mb &mp = t == 0 ? ml : (t == 1 ? mt : mq);
    const int qp = pl == 0 ? (mp.qp + mq.qp + 1) >> 1 : (pl == 1 ? (mp.qpc_u + mq.qpc_u + 1) >> 1 : (mp.qpc_v + mq.qpc_v + 1) >> 1);
    const int ia = iclamp(qp + mq.alpha_off, 0, 51), ib = iclamp(qp + mq.beta_off, 0, 51);
    rec[mb].alpha[t][pl] = c_dbk_alpha[ia]; rec[mb].beta[t][pl] = c_dbk_beta[ib];
#pragma unroll
    for (int b = 0; b < 4; b++) rec[mb].c0[t][pl][b] = c_dbk_clip[ia][b];
  }
  if (lane == 9) rec[mb].t8 = mq.transform8x8;
}

constexpr int DBK_PSTRIDE = 32;                // a row's progress counter has a 128-byte line to itself (67 rows poll and add concurrently)
constexpr int DBK_YP = 36, DBK_CP = 20;      // tile pitches (bytes): rows of a tile start 9 / 5 banks apart

// The wavefront.  One CTA per macroblock row, two warps: warp 0 filters the luma plane, warp 1 both chroma planes (independent
// of each other, each with its own half of the row's progress counter).  Per macroblock the serial chain is: vertical edges (need
// only the row's own samples: the left neighbour's four columns stay in the tile from the previous step, the macroblock's own
// samples and its DbkRec were fetched one step ahead) -> wait for (x + 1, row - 1) -> the four rows above (one L2 round trip) ->
// horizontal edges -> write-back.  Columns 12..15 (chroma 4..7) of a macroblock are written by the NEXT step, right after its left
// edge -- the last thing that modifies them -- and that write is followed by the release the row below waits for: release k of a
// row says that its macroblocks 0 .. k-1 are final, so the rows run one left edge (not one macroblock) behind each other.
__global__ void __launch_bounds__(128) k_deblock(const DbkArgs a, const DbkRec *__restrict__ recs)
{
  __shared__ __align__(16) uint8_t ty[20 * DBK_YP];          // luma rows -4..15, columns -4..15 (+ pad)
  __shared__ __align__(16) uint8_t tc[2][12 * DBK_CP];       // chroma rows -4..7, columns -4..7, per plane
  __shared__ __align__(16) uint32_t srec[2][DBK_RECW];             // the macroblock's DbkRec, one copy per warp
  // The luma chain's two global round trips run on helper warps (warp 2 fetches, warp 3 publishes), so that the luma warp's own
  // per-macroblock chain holds no fence and no global wait: mailboxes of two slots each, hand-over through shared counters.
  __shared__ uint32_t s_left[2][16];                         // left-neighbour columns after the left edge of step x (slot x & 1): luma warp -> publisher
  __shared__ uint32_t s_top[2][16];                          // the four rows above macroblock x (slot x & 1): fetcher -> luma warp
  __shared__ volatile int s_left_ready, s_published, s_top_ready, s_top_taken;
  const int row = blockIdx.x, lane = threadIdx.x & 31, role = threadIdx.x >> 5, chroma = role & 1;
  if (threadIdx.x == 0) { s_left_ready = 0; s_published = 0; s_top_ready = 0; s_top_taken = 0; }
  __syncthreads();
  if (role == 2) {                                           // ---- fetcher: waits for the row above, stages its last four rows ----
    if (row == 0) return;
    for (int x = 0; x < a.mbw; x++) {
      if (lane == 0) {
        const int *pr = a.progress + (row - 1) * DBK_PSTRIDE;
        while ((ld_relaxed(pr) & 0xffff) < x + 1) { }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        while (s_top_taken < x - 1) { }                      // the slot's previous content (macroblock x - 2) has been taken
      }
      __syncwarp();
      if (lane < 16) s_top[x & 1][lane] = ld_cg32(a.y + (size_t)(row * 16 - 4 + (lane >> 2)) * a.yp + x * 16 + 4 * (lane & 3));
      __syncwarp();
      if (lane == 0) { __threadfence_block(); s_top_ready = x + 1; }
    }
    return;
  }
  if (role == 3) {                                           // ---- publisher: writes the left columns, releases the row's counter ----
    for (int x = 1; x < a.mbw; x++) {
      if (lane == 0) { while (s_left_ready < x) { } __threadfence_block(); }
      __syncwarp();
      if (lane < 16) reinterpret_cast<uint32_t *>(a.y + (size_t)(row * 16 + lane) * a.yp + x * 16)[-1] = s_left[x & 1][lane];
      __syncwarp();
      if (lane == 0) { red_release_add(a.progress + row * DBK_PSTRIDE, 1); __threadfence_block(); s_published = x; }
    }
    return;
  }
  const DbkRec &R = *reinterpret_cast<const DbkRec *>(srec[chroma]);
  const DbkRec *rrow = recs + (size_t)row * a.mbw;
  uint32_t pf[4] = {0, 0, 0, 0}, prec = 0;
#ifdef DBK_PROF
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define DBK_T(i) { const long long t_ = clock64(); pc[i] += t_ - pt; pt = t_; }
#else
#define DBK_T(i)
#endif
  // ---- first macroblock: samples and record straight into the tile ----
  if (!chroma) {
    if (lane < 16) {
      const uint8_t *src = a.y + (size_t)(row * 16 + lane) * a.yp;
#pragma unroll
      for (int k = 0; k < 4; k++) reinterpret_cast<uint32_t *>(ty + (4 + lane) * DBK_YP + 4)[k] = ld_cg32(src + 4 * k);
    }
  } else if (lane < 16) {
    const int pl = lane >> 3, r = lane & 7;
    const uint8_t *src = (pl ? a.v : a.u) + (size_t)(row * 8 + r) * a.cp;
#pragma unroll
    for (int k = 0; k < 2; k++) reinterpret_cast<uint32_t *>(tc[pl] + (4 + r) * DBK_CP + 4)[k] = ld_cg32(src + 4 * k);
  }
  if (lane < DBK_RECW) srec[chroma][lane] = reinterpret_cast<const uint32_t *>(rrow)[lane];
  __syncwarp();
  for (int mbx = 0; mbx < a.mbw; mbx++) {
    const bool more = mbx + 1 < a.mbw;
    const bool t8 = R.t8 != 0;
    DBK_T(0)
    // ---- vertical edges ----
    if (!chroma) {
      // a lane filters the same row at all four edges: no exchange between the lanes until the horizontal edges.  After the
      // LEFT edge the left neighbour's last four columns are final -- and with them everything of that macroblock the row
      // below reads: written and released at once, so the row below follows one left edge behind instead of one macroblock
      // the row lives in registers through all four edges (20 samples unpacked once, packed once)
      uint32_t *trow = reinterpret_cast<uint32_t *>(ty + (4 + (lane & 15)) * DBK_YP);
      int px[20];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const uint32_t wv = trow[k];
#pragma unroll
        for (int b = 0; b < 4; b++) px[4 * k + b] = (int)((wv >> (8 * b)) & 255u);
      }
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const uint32_t b4 = *reinterpret_cast<const uint32_t *>(R.bs[0][e]);
        if (b4 != 0u && !((e & 1) && t8) && lane < 16) {
          const int bs = (b4 >> (8 * (lane >> 2))) & 0xff, t = e ? 2 : 0;
          const int alpha = R.alpha[t][0], beta = R.beta[t][0];
          if (bs && (alpha | beta))
            dbk_luma_regs(px[4 * e], px[4 * e + 1], px[4 * e + 2], px[4 * e + 3], px[4 * e + 4], px[4 * e + 5], px[4 * e + 6], px[4 * e + 7], bs, alpha, beta, R.c0[t][0][bs & 3]);
        }
        if (e == 0 && mbx) {
          DBK_T(1)
          if (lane == 0) while (s_published < mbx - 2) { }   // the slot's previous content (step mbx - 2) has been published
          __syncwarp();
          if (lane < 16) s_left[mbx & 1][lane] = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16) | ((uint32_t)px[3] << 24);
          __syncwarp();
          if (lane == 0) { __threadfence_block(); s_left_ready = mbx; }
          DBK_T(7)
        }
      }
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 5; k++) trow[k] = (uint32_t)px[4 * k] | ((uint32_t)px[4 * k + 1] << 8) | ((uint32_t)px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
      }
      __syncwarp();
    } else {
      const int pl = lane >> 4, ed = (lane >> 3) & 1, i = lane & 7, t = ed ? 2 : 0;
      const int bs = R.bs[0][2 * ed][i >> 1];
      const int alpha = R.alpha[t][1 + pl], beta = R.beta[t][1 + pl];
      if (bs && (alpha | beta)) dbk_chroma_line(tc[pl] + (4 + i) * DBK_CP + 4 + 4 * ed, 1, bs, alpha, beta, R.c0[t][1 + pl][bs & 3]);
      __syncwarp();
      if (mbx) {
        if (lane < 16) reinterpret_cast<uint32_t *>((lane >> 3 ? a.v : a.u) + (size_t)(row * 8 + (lane & 7)) * a.cp + mbx * 8)[-1] =
                         reinterpret_cast<const uint32_t *>(tc[lane >> 3] + (4 + (lane & 7)) * DBK_CP)[0];
        __syncwarp();
        if (lane == 0) red_release_add(a.progress + row * DBK_PSTRIDE, 0x10000);
      }
    }
    // ---- one step ahead: the next macroblock's own samples (nobody modifies them before this row does) and its record; issued
    //      after the release above (a fence waits for the loads in flight) and consumed at the end of the step ----
    if (more) {
      if (!chroma) {
        if (lane < 16) {
          const uint8_t *src = a.y + (size_t)(row * 16 + lane) * a.yp + (mbx + 1) * 16;
#pragma unroll
          for (int k = 0; k < 4; k++) pf[k] = ld_cg32(src + 4 * k);
        }
      } else if (lane < 16) {
        const int pl = lane >> 3, r = lane & 7;
        const uint8_t *src = (pl ? a.v : a.u) + (size_t)(row * 8 + r) * a.cp + (mbx + 1) * 8;
#pragma unroll
        for (int k = 0; k < 2; k++) pf[k] = ld_cg32(src + 4 * k);
      }
      if (lane < DBK_RECW) prec = reinterpret_cast<const uint32_t *>(rrow + mbx + 1)[lane];
    }
    DBK_T(1)
    // ---- the row above: (mbx + 1, row - 1) must be done; then its last four (chroma: two) rows over this macroblock ----
    if (row > 0 && !chroma) {
      if (lane == 0) { while (s_top_ready < mbx + 1) { } __threadfence_block(); }
      __syncwarp();
      DBK_T(2)
      if (lane < 16) reinterpret_cast<uint32_t *>(ty + (lane >> 2) * DBK_YP + 4)[lane & 3] = s_top[mbx & 1][lane];
      __syncwarp();
      if (lane == 0) s_top_taken = mbx + 1;
    }
    if (row > 0 && chroma) {
      const int need = mbx + 1;                       // release k of the row above: its macroblocks 0 .. k-1 are final
      if (lane == 0) {
        const int *pr = a.progress + (row - 1) * DBK_PSTRIDE;
        while (((ld_relaxed(pr) >> 16) & 0xffff) < need) { }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");       // acquire: one fence after the spin instead of one per poll
      }
      __syncwarp();
      DBK_T(2)
      if (lane < 8) {
        const int pl = lane >> 2, r = 2 + ((lane >> 1) & 1), k = lane & 1;
        reinterpret_cast<uint32_t *>(tc[pl] + r * DBK_CP + 4)[k] = ld_cg32((pl ? a.v : a.u) + (size_t)(row * 8 - 4 + r) * a.cp + mbx * 8 + 4 * k);
      }
      __syncwarp();
    }
    DBK_T(3)
    // ---- horizontal edges ----
    if (!chroma) {
      // a lane keeps one COLUMN (rows -4 .. 15) in registers through the four horizontal edges
      uint8_t *tcol = ty + 4 + (lane & 15);
      int px[20];
#pragma unroll
      for (int k = 0; k < 20; k++) px[k] = tcol[k * DBK_YP];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const uint32_t b4 = *reinterpret_cast<const uint32_t *>(R.bs[1][e]);
        if (b4 != 0u && !((e & 1) && t8) && lane < 16) {
          const int bs = (b4 >> (8 * (lane >> 2))) & 0xff, t = e ? 2 : 1;
          const int alpha = R.alpha[t][0], beta = R.beta[t][0];
          if (bs && (alpha | beta))
            dbk_luma_regs(px[4 * e], px[4 * e + 1], px[4 * e + 2], px[4 * e + 3], px[4 * e + 4], px[4 * e + 5], px[4 * e + 6], px[4 * e + 7], bs, alpha, beta, R.c0[t][0][bs & 3]);
        }
      }
      if (lane < 16) {
#pragma unroll
        for (int k = 1; k < 19; k++) tcol[k * DBK_YP] = (uint8_t)px[k];      // rows -3 .. 14: what an edge can change
      }
      __syncwarp();
    } else {
      const int pl = lane >> 4, ed = (lane >> 3) & 1, i = lane & 7, t = ed ? 2 : 1;
      const int bs = R.bs[1][2 * ed][i >> 1];
      const int alpha = R.alpha[t][1 + pl], beta = R.beta[t][1 + pl];
      if (bs && (alpha | beta)) dbk_chroma_line(tc[pl] + (4 + 4 * ed) * DBK_CP + 4 + i, DBK_CP, bs, alpha, beta, R.c0[t][1 + pl][bs & 3]);
      __syncwarp();
    }
    DBK_T(4)
    // ---- write-back: the three (one) rows above that the top edge may have changed; own rows from the left neighbour's
    //      four columns up to column 11 (3), the last four columns only at the end of the row ----
    if (!chroma) {
      if (lane < 16) {
        uint8_t *dst = a.y + (size_t)(row * 16 + lane) * a.yp + mbx * 16;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(ty + (4 + lane) * DBK_YP + 4);
#pragma unroll
        for (int k = 0; k < 3; k++) reinterpret_cast<uint32_t *>(dst)[k] = src[k];
        if (!more) reinterpret_cast<uint32_t *>(dst)[3] = src[3];
      } else if (row > 0 && lane < 28) {
        const int r = 1 + (lane - 16) / 4, k = lane & 3;
        reinterpret_cast<uint32_t *>(a.y + (size_t)(row * 16 - 4 + r) * a.yp + mbx * 16)[k] = reinterpret_cast<const uint32_t *>(ty + r * DBK_YP + 4)[k];
      }
    } else {
      if (lane < 16) {
        const int pl = lane >> 3, r = lane & 7;
        uint8_t *dst = (pl ? a.v : a.u) + (size_t)(row * 8 + r) * a.cp + mbx * 8;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tc[pl] + (4 + r) * DBK_CP + 4);
        reinterpret_cast<uint32_t *>(dst)[0] = src[0];
        if (!more) reinterpret_cast<uint32_t *>(dst)[1] = src[1];
      } else if (row > 0 && lane < 20) {
        const int pl = (lane >> 1) & 1, k = lane & 1;
        reinterpret_cast<uint32_t *>((pl ? a.v : a.u) + (size_t)(row * 8 - 1) * a.cp + mbx * 8)[k] = reinterpret_cast<const uint32_t *>(tc[pl] + 3 * DBK_CP + 4)[k];
      }
    }
    __syncwarp();
    if (!more && lane == 0) {                              // the row's last release: everything is final
      if (!chroma) { while (s_published < a.mbw - 1) { } __threadfence_block(); }   // after the publisher's last one (the counter only counts)
      red_release_add(a.progress + row * DBK_PSTRIDE, chroma ? 0x10000 : 1);
    }
    DBK_T(5)
    // ---- next macroblock: its left neighbour's columns are this tile's last four; own samples and record from the registers ----
    if (more) {
      if (!chroma) {
        if (lane < 16) {
          uint32_t *t = reinterpret_cast<uint32_t *>(ty + (4 + lane) * DBK_YP);
          t[0] = t[4];
#pragma unroll
          for (int k = 0; k < 4; k++) t[1 + k] = pf[k];
        }
      } else if (lane < 16) {
        uint32_t *t = reinterpret_cast<uint32_t *>(tc[lane >> 3] + (4 + (lane & 7)) * DBK_CP);
        t[0] = t[2]; t[1] = pf[0]; t[2] = pf[1];
      }
      if (lane < DBK_RECW) srec[chroma][lane] = prec;
      __syncwarp();
    }
    DBK_T(6)
  }
#ifdef DBK_PROF
  if (lane == 0 && (row == 0 || row == a.mbh / 2))
    printf("[dbk] row %d %s: cycles per macroblock: prefetch issue %lld, vertical edges (+ release) %lld, wait %lld, rows above %lld, horizontal edges %lld, write-back %lld, carry %lld; write + release of the left columns %lld\n",
           row, chroma ? "chroma" : "luma", pc[0] / a.mbw, pc[1] / a.mbw, pc[2] / a.mbw, pc[3] / a.mbw, pc[4] / a.mbw, pc[5] / a.mbw, pc[6] / a.mbw, pc[7] / a.mbw);
#endif
}

}  // namespace b2

using namespace b2;

static thread_local char g_dbkerr[256];
extern "C" const char *b2dbk_last_error(void) { return g_dbkerr; }

extern "C" int b2dbk_frame_dev(int W, int H, uint8_t *y, int y_pitch, uint8_t *u, uint8_t *v, int c_pitch,
                               const b2dbk_mb *mbs, const b2dbk_blk *blks, int *progress, void *stream)
{
  if (W <= 0 || H <= 0 || (W & 15) || (H & 15) || !y || !u || !v || !mbs || !blks || y_pitch < W || c_pitch < W / 2 || (y_pitch & 3) || (c_pitch & 3) ||
      ((uintptr_t)y & 3) || ((uintptr_t)u & 3) || ((uintptr_t)v & 3)) {
    snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: picture size must be a multiple of 16, planes and pitches 4-byte aligned");
    return B2ME_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0, occ = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_deblock, 128, 0);
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  if (H / 16 > occ * sms) {                         // every macroblock row must be resident: a row waits on the row above
    snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %d macroblock rows exceed the %d resident row CTAs of this device", H / 16, occ * sms);
    return B2ME_EUNSUPPORTED;
  }
  DbkArgs a;
  a.W = W; a.H = H; a.mbw = W / 16; a.mbh = H / 16; a.y = y; a.yp = y_pitch; a.u = u; a.v = v; a.cp = c_pitch; a.mbs = mbs; a.blks = blks; a.progress = progress;
  b2_pool_retain(dev);
  DbkRec *rec = nullptr;
  const int nmb = a.mbw * a.mbh;
  const size_t cnt_bytes = sizeof(int) * DBK_PSTRIDE * (size_t)a.mbh;
  // stream-ordered scratch: the rows' progress counters (one 128-byte line each; the caller's `progress` array is too dense
  // for 67 rows polling and adding at once and is left untouched) + one DbkRec per macroblock
  e = cudaMallocAsync(reinterpret_cast<void **>(&rec), cnt_bytes + sizeof(DbkRec) * (size_t)nmb, s);
  if (e == cudaSuccess) { a.progress = reinterpret_cast<int *>(rec); rec = reinterpret_cast<DbkRec *>(reinterpret_cast<uint8_t *>(rec) + cnt_bytes); }
  if (e == cudaSuccess) e = cudaMemsetAsync(a.progress, 0, cnt_bytes, s);
  if (e == cudaSuccess) { k_dbk_prep<<<(nmb + 3) / 4, 128, 0, s>>>(a, rec); e = cudaGetLastError(); }
  if (e == cudaSuccess) { k_deblock<<<a.mbh, 128, 0, s>>>(a, rec); e = cudaGetLastError(); }
  if (rec) { const cudaError_t e2 = cudaFreeAsync(a.progress, s); if (e == cudaSuccess) e = e2; }
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  return B2ME_OK;
}

extern "C" int b2dbk_frame(int device, int W, int H, uint8_t *y, uint8_t *u, uint8_t *v, const b2dbk_mb *mbs, const b2dbk_blk *blks)
{
  if (W <= 0 || H <= 0 || (W & 15) || (H & 15) || !y || !u || !v || !mbs || !blks) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: bad arguments"); return B2ME_EINVAL; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  const size_t ny = (size_t)W * H, nc = ny / 4, nm = sizeof(b2dbk_mb) * (W / 16) * (H / 16), nb = sizeof(b2dbk_blk) * (W / 4) * (H / 4), np = sizeof(int) * (H / 16);
  const size_t o_u = ny, o_v = o_u + nc, o_m = (o_v + nc + 15) & ~(size_t)15, o_b = (o_m + nm + 15) & ~(size_t)15, o_p = (o_b + nb + 15) & ~(size_t)15;
  uint8_t *d = nullptr;
  e = cudaMalloc(&d, o_p + np);
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  cudaMemcpy(d, y, ny, cudaMemcpyHostToDevice); cudaMemcpy(d + o_u, u, nc, cudaMemcpyHostToDevice); cudaMemcpy(d + o_v, v, nc, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_m, mbs, nm, cudaMemcpyHostToDevice); cudaMemcpy(d + o_b, blks, nb, cudaMemcpyHostToDevice);
  int r = b2dbk_frame_dev(W, H, d, W, d + o_u, d + o_v, W / 2, reinterpret_cast<const b2dbk_mb *>(d + o_m), reinterpret_cast<const b2dbk_blk *>(d + o_b),
                          reinterpret_cast<int *>(d + o_p), 0);
  if (!r) {
    e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(y, d, ny, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(u, d + o_u, nc, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(v, d + o_v, nc, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  }
  cudaFree(d);
  return r;
}
