// bipred.cu -- bi-predictive block search (B slices): one CTA per call of
//   full_search_bipred_motion_estimation   JM/lencod/src/me_fullsearch.c:112-176
//   sub_pel_bipred_motion_estimation       JM/lencod/src/me_fullsearch.c:300-399
// with the distortions behind them,
//   computeBiPredSAD1/2   JM/lencod/src/me_distortion.c:525-620 / 628-737
//   computeBiPredSATD1/2  :943-1040 / 1048-1182      computeBiPredSSE1/2  :1353-1442 / 1450-1549
// sequenced as BiPredBlockMotionSearch does (mv_search.c:1100-1126).
//
// The block of the searched list moves over the spiral, the block of the other list stays where the caller's vector
// puts it; the distortion is taken against their (weighted) average.  Every candidate is independent: candidates
// are spread over the CTA's threads and the reference's strict-'<' scan order is the lexicographic minimum of
// (cost, spiral position), kept with a 64-bit shared atomicMin that starts from the caller's bound.  The early
// exits of the reference only ever skip candidates that cannot win (they return a value >= the bound).
// Both blocks come straight from the 16 quarter-pel planes (UMVLine4X clamp, refbuf.h:25); this path serves B
// slices, which none of the BASELINE configs code, so it is written for exactness first.
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

struct BiArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, nrefs, R;
  int lambda[3], metric_h, metric_q, do_subpel, full81, test8x8, wp, denom;
  const b2me_bipred_job *jobs; b2me_bipred_result *out; int njobs;
  int *errflag;
};

__device__ __forceinline__ const uint8_t *bi_umv(const BiArgs &a, int ref, int qx, int qy)
{
  const uint8_t *pl = a.planes + ((size_t)ref * 16 + (qy & 3) * 4 + (qx & 3)) * a.plane_size;
  const int yy = iclamp(qy >> 2, -PADY, a.H + 3), xx = iclamp(qx >> 2, -PADX, a.W + 15);
  return pl + (size_t)(yy + PADY) * a.Wp + (xx + PADX);
}

struct BiWp { int wp, w1, w2, off, shift, lround; };
__device__ __forceinline__ int bi_pel(int r1, int r2, const BiWp &w)
{
  if (!w.wp) return (r1 + r2 + 1) >> 1;
  return iclamp(((w.w1 * r1 + w.w2 * r2 + w.lround) >> w.shift) + w.off, 0, 255);
}

__device__ __forceinline__ int bi_had4(int *d)          // sum |H4 D H4|, (s + 1) >> 1   (HadamardSAD4x4, me_distortion.c:175-258)
{
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int p = d[4 * i], q = d[4 * i + 1], r = d[4 * i + 2], s = d[4 * i + 3];
    d[4 * i] = p + q + r + s; d[4 * i + 1] = p - q + r - s; d[4 * i + 2] = p + q - r - s; d[4 * i + 3] = p - q - r + s;
  }
  int t = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int p = d[i], q = d[4 + i], r = d[8 + i], s = d[12 + i];
    t += abs(p + q + r + s) + abs(p - q + r - s) + abs(p + q - r - s) + abs(p - q - r + s);
  }
  return (t + 1) >> 1;
}
__device__ __noinline__ int bi_had8(short *d)           // sum |H8 D H8|, (s + 2) >> 2   (HadamardSAD8x8, :266-341)
{
  for (int j = 0; j < 8; j++)
    for (int k = 1; k < 8; k <<= 1)
      for (int i = 0; i < 8; i++)
        if (!(i & k)) { const short p = d[8 * j + i], q = d[8 * j + (i | k)]; d[8 * j + i] = p + q; d[8 * j + (i | k)] = p - q; }
  int t = 0;
  for (int i = 0; i < 8; i++) {
    int c[8];
    for (int j = 0; j < 8; j++) c[j] = d[8 * j + i];
    for (int k = 1; k < 8; k <<= 1)
      for (int j = 0; j < 8; j++)
        if (!(j & k)) { const int p = c[j], q = c[j | k]; c[j] = p + q; c[j | k] = p - q; }
    for (int j = 0; j < 8; j++) t += abs(c[j]);
  }
  return (t + 2) >> 2;
}

// distortion of the block against the average of (ref1 at c1) and (ref2 at c2); metric 0 SAD, 1 SSE, 2 SATD
__device__ int bi_dist(const BiArgs &a, const uint8_t *cur /* smem, pitch 16 */, int bsx, int bsy, int ref1, int ref2,
                       int c1x, int c1y, int c2x, int c2y, int metric, const BiWp &w)
{
  int s = 0;
  if (metric != 2) {                                  // one origin clamp per reference (me_distortion.c:546-547)
    const uint8_t *r1 = bi_umv(a, ref1, c1x, c1y), *r2 = bi_umv(a, ref2, c2x, c2y);
    for (int y = 0; y < bsy; y++)
      for (int x = 0; x < bsx; x++) {
        const int d = (int)cur[y * 16 + x] - bi_pel(r1[(size_t)y * a.Wp + x], r2[(size_t)y * a.Wp + x], w);
        s += metric == 0 ? abs(d) : d * d;
      }
    return s;
  }
  if (!a.test8x8) {                                   // every 4x4 tile origin is clamped on its own (:971-972)
    for (int by = 0; by < bsy; by += 4)
      for (int bx = 0; bx < bsx; bx += 4) {
        const uint8_t *r1 = bi_umv(a, ref1, c1x + (bx << 2), c1y + (by << 2)), *r2 = bi_umv(a, ref2, c2x + (bx << 2), c2y + (by << 2));
        int d[16];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int i = 0; i < 4; i++)
            d[j * 4 + i] = (int)cur[(by + j) * 16 + bx + i] - bi_pel(r1[(size_t)j * a.Wp + i], r2[(size_t)j * a.Wp + i], w);
        s += bi_had4(d);
      }
    return s;
  }
  for (int by = 0; by < bsy; by += 8)
    for (int bx = 0; bx < bsx; bx += 8) {
      const uint8_t *r1 = bi_umv(a, ref1, c1x + (bx << 2), c1y + (by << 2)), *r2 = bi_umv(a, ref2, c2x + (bx << 2), c2y + (by << 2));
      short d[64];
      for (int j = 0; j < 8; j++)
        for (int i = 0; i < 8; i++)
          d[j * 8 + i] = (short)((int)cur[(by + j) * 16 + bx + i] - bi_pel(r1[(size_t)j * a.Wp + i], r2[(size_t)j * a.Wp + i], w));
      s += bi_had8(d);
    }
  return s;
}

constexpr long long BI_DISTBLK_MAX = ((long long)0x7fffffff) << 5;

__global__ void __launch_bounds__(128, 8) k_bipred(const BiArgs a)   // 8 CTAs per SM: 64 registers (127 left four CTAs = 16 warps per SM for a latency-bound kernel)
{
  __shared__ uint8_t cur[256];
  __shared__ unsigned long long key;
  __shared__ int mv1[2];
  __shared__ long long bound;
  const int tid = threadIdx.x;
  const b2me_bipred_job J = a.jobs[blockIdx.x];
  b2me_bipred_result *O = &a.out[blockIdx.x];
  const int bt = J.blocktype;
  const bool bad = bt < 1 || bt > 7 || J.ref1 < 0 || J.ref1 >= a.nrefs || J.ref2 < 0 || J.ref2 >= a.nrefs || J.search_range < -1 ||
                   J.search_range > a.R || J.pos_x < 0 || J.pos_y < 0 || J.pos_x + part_geom(part_first(bt < 1 || bt > 7 ? 7 : bt)).w > a.W ||
                   J.pos_y + part_geom(part_first(bt < 1 || bt > 7 ? 7 : bt)).h > a.H ||
                   (J.search_range >= 0 && ((J.mv1[0] | J.mv1[1]) & 3)) || (J.search_range < 0 && !a.do_subpel);
  if (bad) {                                          // uniform over the CTA
    if (tid == 0) { *a.errflag = 1; O->cost_int = O->cost_sub = -1; O->mv_int[0] = O->mv_int[1] = O->mv_sub[0] = O->mv_sub[1] = 0; }
    return;
  }
  const PartGeom gm = part_geom(part_first(bt));
  const int bsx = gm.w, bsy = gm.h;
  for (int i = tid; i < 256; i += 128) {
    const int x = i & 15, y = i >> 4;
    cur[i] = (x < bsx && y < bsy && J.pos_x + x < a.W && J.pos_y + y < a.H) ? a.cur[(size_t)(J.pos_y + y) * a.cur_pitch + J.pos_x + x] : 0;
  }
  BiWp w; w.wp = a.wp; w.w1 = J.weight1; w.w2 = J.weight2; w.off = J.offset_bi; w.shift = a.denom + 1;
  w.lround = 2 * (a.denom ? 1 << (a.denom - 1) : 0);
  const int ox = J.pos_x << 2, oy = J.pos_y << 2;
  const int c2x = ox + J.mv2[0], c2y = oy + J.mv2[1];
  long long min_mcost = J.min_mcost < 0 ? 0 : (J.min_mcost > BI_DISTBLK_MAX ? BI_DISTBLK_MAX : J.min_mcost);
  if (tid == 0) { key = (unsigned long long)min_mcost << 20; mv1[0] = J.mv1[0]; mv1[1] = J.mv1[1]; }
  __syncthreads();
  // ---- integer pel: (2 sr + 1)^2 spiral positions around mv1, SAD ----
  if (J.search_range >= 0) {                          // search_range == -1: sub_pel_bipred_motion_estimation alone
    const int sr = J.search_range, max_pos = (2 * sr + 1) * (2 * sr + 1);
    const long long c2 = (long long)a.lambda[0] * (mvbits(J.mv2[0] - J.pred2[0]) + mvbits(J.mv2[1] - J.pred2[1]));
    for (int pos = tid; pos < max_pos; pos += 128) {
      int sx, sy; spiral_xy(pos, &sx, &sy);
      const int cx = J.mv1[0] + 4 * sx, cy = J.mv1[1] + 4 * sy;
      long long mcost = (long long)a.lambda[0] * (mvbits(cx - J.pred1[0]) + mvbits(cy - J.pred1[1])) + c2;
      if ((unsigned long long)mcost >= (*reinterpret_cast<volatile unsigned long long *>(&key) >> 20)) continue;
      mcost += (long long)bi_dist(a, cur, bsx, bsy, J.ref1, J.ref2, ox + cx, oy + cy, c2x, c2y, 0, w) << 5;
      atomicMin(&key, ((unsigned long long)mcost << 20) | (unsigned)pos);
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int pos = (int)(key & 0xfffffull);
    int sx, sy; spiral_xy(pos, &sx, &sy);
    mv1[0] += 4 * sx; mv1[1] += 4 * sy;
    long long c = (long long)(key >> 20);
    if (pos == 0 && c == BI_DISTBLK_MAX && J.min_mcost > BI_DISTBLK_MAX) c = J.min_mcost;
    O->mv_int[0] = (int16_t)mv1[0]; O->mv_int[1] = (int16_t)mv1[1]; O->cost_int = c;
    O->mv_sub[0] = (int16_t)mv1[0]; O->mv_sub[1] = (int16_t)mv1[1]; O->cost_sub = c;
    bound = c;
  }
  if (!a.do_subpel) return;
  // ---- half pel, then quarter pel: positions start..8 of the spiral, step 2 / 1 quarter-pel ----
  const int start_hp_cfg = (0 != a.metric_h) ? 0 : 1, start_qp = (a.metric_h != a.metric_q) ? 0 : 1;
  // full81: full_sub_pel_bipred_motion_estimation (me_fullsearch.c:478-538) -- ONE stage, the 81 quarter-pel positions of
  // spiral_search, computeBiPredQPel, lambda[Q_PEL], every position from 0 on
  const int nstage = a.full81 ? 1 : 2;
#pragma unroll 1
  for (int stage = 0; stage < nstage; stage++) {
    __syncthreads();
    long long b = bound;
    if (stage == 0 && !start_hp_cfg && J.search_range >= 0) b = BI_DISTBLK_MAX;   // the caller's reset, mv_search.c:1119-1120
    const int start = a.full81 ? 0 : (stage ? start_qp : (b == BI_DISTBLK_MAX ? 0 : start_hp_cfg));
    if (stage == 1 && !start_qp) b = BI_DISTBLK_MAX;                  // me_fullsearch.c:364-365
    const int m0 = mv1[0], m1 = mv1[1];
    __syncthreads();
    if (tid == 0) key = (unsigned long long)b << 20;
    __syncthreads();
    const int lam = a.full81 ? a.lambda[2] : a.lambda[1 + stage], metric = (stage || a.full81) ? a.metric_q : a.metric_h, step = (stage || a.full81) ? 1 : 2;
    if (tid >= start && tid < (a.full81 ? 81 : 9)) {
      int sx, sy; spiral_xy(tid, &sx, &sy);
      const int cx = m0 + step * sx, cy = m1 + step * sy;
      long long mcost = (long long)lam * (mvbits(cx - J.pred1[0]) + mvbits(cy - J.pred1[1]) + mvbits(J.mv2[0] - J.pred2[0]) + mvbits(J.mv2[1] - J.pred2[1]));
      if (mcost < b) {
        mcost += (long long)bi_dist(a, cur, bsx, bsy, J.ref1, J.ref2, ox + cx, oy + cy, c2x, c2y, metric, w) << 5;
        atomicMin(&key, ((unsigned long long)mcost << 20) | (unsigned)tid);
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int pos = (int)(key & 0xfffffull);
      int sx, sy; spiral_xy(pos, &sx, &sy);
      mv1[0] += step * sx; mv1[1] += step * sy;
      bound = (long long)(key >> 20);
    }
  }
  __syncthreads();
  if (tid == 0) { O->mv_sub[0] = (int16_t)mv1[0]; O->mv_sub[1] = (int16_t)mv1[1]; O->cost_sub = bound; }
}

// ---- bi-predictive distortion at explicit candidate pairs: computeBiPredSAD1 / SAD2, SSE1 / SSE2, SATD1 / SATD2
//      (me_distortion.c:525-737, 943-1182, 1353-1549) at their own boundary (mv_block->computeBiPred1[] / computeBiPred2[]):
//      the record is b2me_bipred_job (position, block type, two slots, mv1 = candidate of list A, mv2 = candidate of list B,
//      weights); out = distortion << 5.  One warp per record, the serial distortion routine of k_bipred by lane 0. ----
__global__ void __launch_bounds__(32) k_bicand(const BiArgs a, int metric, long long *out)
{
  __shared__ uint8_t cur[256];
  const int tid = threadIdx.x;
  const b2me_bipred_job J = a.jobs[blockIdx.x];
  const int bt = J.blocktype;
  const bool bad = bt < 1 || bt > 7 || J.ref1 < 0 || J.ref1 >= a.nrefs || J.ref2 < 0 || J.ref2 >= a.nrefs || J.pos_x < 0 || J.pos_y < 0 ||
                   J.pos_x + part_geom(part_first(bt < 1 || bt > 7 ? 7 : bt)).w > a.W || J.pos_y + part_geom(part_first(bt < 1 || bt > 7 ? 7 : bt)).h > a.H;
  if (bad) { if (tid == 0) { *a.errflag = 1; out[blockIdx.x] = -1; } return; }
  const PartGeom gm = part_geom(part_first(bt));
  for (int i = tid; i < 256; i += 32) {
    const int x = i & 15, y = i >> 4;
    cur[i] = (x < gm.w && y < gm.h) ? a.cur[(size_t)(J.pos_y + y) * a.cur_pitch + J.pos_x + x] : 0;
  }
  __syncwarp();
  if (tid == 0) {
    BiWp w; w.wp = a.wp; w.w1 = J.weight1; w.w2 = J.weight2; w.off = J.offset_bi; w.shift = a.denom + 1;
    w.lround = 2 * (a.denom ? 1 << (a.denom - 1) : 0);
    const int ox = J.pos_x << 2, oy = J.pos_y << 2;
    out[blockIdx.x] = (long long)bi_dist(a, cur, gm.w, gm.h, J.ref1, J.ref2, ox + J.mv1[0], oy + J.mv1[1], ox + J.mv2[0], oy + J.mv2[1], metric, w) << 5;
  }
}

// ---- BIDPartitionCost (mv_search.c:1159-1250): the bi-predictive direction's motion cost in the mode decision.  One warp per
//      partition.  The region (<= 16x16) is predicted sub-block by sub-block -- each with its own vector pair and ONE origin clamp
//      per list (OneComponentLumaPrediction -> UMVLine4X, mc_prediction.c:117-136), bi_prediction / weighted_bi_prediction -- into a
//      shared residual tile; then a lane per 4x4 block (per 8x8 block with the 8x8 transform on and blocktype <= 4) takes the mode
//      decision's distortion (distortion4x4 / 8x8 of select_distortion: SAD, SSE or the Hadamard sums) and the warp adds them up:
//      cost = lambda_factor * mvd_bits + (sum << 5). ----
struct BidArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, nrefs, metric, t8, wp, denom, n;
  const b2me_bid_job *jobs; long long *out; int *errflag;
};
__constant__ uint8_t c_bid_bx0[5][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 2, 0, 0}, {0, 2, 0, 2}};     // mv_search.c:57-58
__constant__ uint8_t c_bid_by0[5][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 2, 0, 0}, {0, 0, 0, 0}, {0, 0, 2, 2}};
__constant__ uint8_t c_bid_bs[8][2] = {{16, 16}, {16, 16}, {16, 8}, {8, 16}, {8, 8}, {8, 4}, {4, 8}, {4, 4}};      // block_size[][]

__global__ void __launch_bounds__(128) k_bid_cost(const BidArgs a)
{
  __shared__ short sdiff[4][256];                    // residual of the region, pitch 16, per warp
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, job = blockIdx.x * 4 + w;
  if (job >= a.n) return;
  const b2me_bid_job J = a.jobs[job];
  const int bt = J.blocktype;
  const bool bad = bt < 1 || bt > 7 || J.block8x8 < 0 || J.block8x8 > 3 || J.ref_l0 < 0 || J.ref_l0 >= a.nrefs || J.ref_l1 < 0 || J.ref_l1 >= a.nrefs ||
                   J.mb_x < 0 || J.mb_y < 0 || J.mb_x + 16 > a.W || J.mb_y + 16 > a.H;
  if (bad) { if (lane == 0) { *a.errflag = 1; a.out[job] = -1; } return; }
  const int pt = bt < 4 ? bt : 4;
  const int bx = c_bid_bx0[pt][J.block8x8] << 2, by = c_bid_by0[pt][J.block8x8] << 2;
  const int w0 = c_bid_bs[pt][0], h0 = c_bid_bs[pt][1], sw = c_bid_bs[bt][0], sh = c_bid_bs[bt][1];
  const int nsx = w0 / sw;
  BiArgs ba; ba.planes = a.planes; ba.plane_size = a.plane_size; ba.W = a.W; ba.H = a.H; ba.Wp = a.Wp;     // what bi_umv reads
  BiWp wp; wp.wp = a.wp; wp.w1 = J.weight_l0; wp.w2 = J.weight_l1; wp.off = J.offset_bi; wp.shift = a.denom + 1;
  wp.lround = 2 * (a.denom ? 1 << (a.denom - 1) : 0);
  short *d = sdiff[w];
  for (int i = lane; i < w0 * h0; i += 32) {
    const int x = i % w0, y = i / w0;
    const int sb = (y / sh) * nsx + x / sw, ox = (x / sw) * sw, oy = (y / sh) * sh;     // sub-block and its origin inside the region
    const int qx = (J.mb_x + bx + ox) << 2, qy = (J.mb_y + by + oy) << 2;
    const uint8_t *r0 = bi_umv(ba, J.ref_l0, qx + J.mv_l0[sb][0], qy + J.mv_l0[sb][1]);
    const uint8_t *r1 = bi_umv(ba, J.ref_l1, qx + J.mv_l1[sb][0], qy + J.mv_l1[sb][1]);
    const size_t o = (size_t)(y - oy) * a.Wp + (x - ox);
    d[y * 16 + x] = (short)((int)a.cur[(size_t)(J.mb_y + by + y) * a.cur_pitch + J.mb_x + bx + x] - bi_pel(r0[o], r1[o], wp));
  }
  __syncwarp();
  const bool use8 = a.t8 && bt <= 4;
  const int bsz = use8 ? 8 : 4, nbx = w0 / bsz, nb = nbx * (h0 / bsz);
  int sum = 0;
  if (lane < nb) {
    const short *p = d + (lane / nbx) * bsz * 16 + (lane % nbx) * bsz;
    if (use8) {
      short t[64];
      for (int j = 0; j < 8; j++) for (int i = 0; i < 8; i++) t[j * 8 + i] = p[j * 16 + i];
      if (a.metric == 2) sum = bi_had8(t);
      else for (int i = 0; i < 64; i++) sum += a.metric == 0 ? abs((int)t[i]) : (int)t[i] * t[i];
    } else {
      int t[16];
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) t[j * 4 + i] = p[j * 16 + i];
      if (a.metric == 2) sum = bi_had4(t);
      else {
#pragma unroll
        for (int i = 0; i < 16; i++) sum += a.metric == 0 ? abs(t[i]) : t[i] * t[i];
      }
    }
  }
  sum = __reduce_add_sync(0xffffffffu, sum);
  if (lane == 0) a.out[job] = (long long)J.lambda_factor * J.mvd_bits + ((long long)sum << 5);
}

extern "C" int b2me_bid_partition_cost(b2me_ctx *c, int metric, int transform8x8, int apply_weights, int luma_log_weight_denom, int n,
                                       const b2me_bid_job *jobs, int64_t *out)
{
  if (!c || n < 0 || (n && (!jobs || !out)) || metric < 0 || metric > 2 || luma_log_weight_denom < 0 || luma_log_weight_denom > 7) return B2ME_EINVAL;
  if (!n) return B2ME_OK;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  uint8_t *d = nullptr;
  const size_t bj = (sizeof(b2me_bid_job) * (size_t)n + 15) & ~(size_t)15;
  B2_CUDA_CHECK(c, cudaMallocAsync(&d, bj + sizeof(long long) * (size_t)n, c->stream));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(d, jobs, sizeof(b2me_bid_job) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  BidArgs a;
  a.cur = c->d_cur; a.cur_pitch = c->W; a.planes = c->d_planes; a.plane_size = c->plane_size;
  a.W = c->W; a.H = c->H; a.Wp = c->Wp; a.nrefs = c->nrefs; a.metric = metric; a.t8 = transform8x8 ? 1 : 0; a.wp = apply_weights ? 1 : 0;
  a.denom = luma_log_weight_denom; a.n = n; a.jobs = reinterpret_cast<const b2me_bid_job *>(d); a.out = reinterpret_cast<long long *>(d + bj);
  a.errflag = c->d_errflag;
  k_bid_cost<<<(n + 3) / 4, 128, 0, c->stream>>>(a);
  int r = B2ME_OK, flag = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + bj, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "b2me_bid_partition_cost: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  else if (flag) {
    cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream);
    snprintf(c->err, sizeof(c->err), "b2me_bid_partition_cost: a record is out of range (blocktype, partition index, reference slot or position)");
    r = B2ME_EINVAL;
  }
  c->launches++;
  cudaFreeAsync(d, c->stream);
  return r;
}

// ---- distortion of a list of (block, reference, candidate) triples ---------------------------------------------
// computeSAD / computeSSE / computeSATD (me_distortion.c:349-426, 1190-1255, 745-825) at their own boundary
// (mv_block->computePredFPel / HPel / QPel): what a search whose control flow stays on the host (EPZS, UMHex) asks
// for per predictor set.  One warp per candidate; the value is the reference's return without early exit,
// dist_scale(distortion) = distortion << 5, no motion-vector cost.
struct CandArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, nrefs, metric, test8x8, n;
  const b2me_candidate *cands; long long *out; int *errflag;
};

__global__ void __launch_bounds__(128) k_cand_dist(const CandArgs a)
{
  const int lane = threadIdx.x & 31, i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= a.n) return;
  const b2me_candidate c = a.cands[i];
  if (c.blocktype < 1 || c.blocktype > 7 || c.ref < 0 || c.ref >= a.nrefs || c.pos_x < 0 || c.pos_y < 0 ||
      c.pos_x + part_geom(part_first(c.blocktype)).w > a.W || c.pos_y + part_geom(part_first(c.blocktype)).h > a.H) {
    if (lane == 0) { *a.errflag = 1; a.out[i] = -1; }
    return;
  }
  const PartGeom gm = part_geom(part_first(c.blocktype));
  const int bsx = gm.w, bsy = gm.h;
  const uint8_t *cur = a.cur + (size_t)c.pos_y * a.cur_pitch + c.pos_x;
  const int qx = (c.pos_x << 2) + c.mv[0], qy = (c.pos_y << 2) + c.mv[1];
  BiArgs u; u.planes = a.planes; u.plane_size = a.plane_size; u.W = a.W; u.H = a.H; u.Wp = a.Wp;
  int s = 0;
  if (a.metric != 2) {
    const uint8_t *r = bi_umv(u, c.ref, qx, qy);
    for (int k = lane; k < bsx * bsy; k += 32) {
      const int x = k % bsx, y = k / bsx;
      const int d = (int)cur[(size_t)y * a.cur_pitch + x] - (int)r[(size_t)y * a.Wp + x];
      s += a.metric == 0 ? abs(d) : d * d;
    }
  } else if (!a.test8x8) {
    const int tw = bsx >> 2, nt = tw * (bsy >> 2);
    if (lane < nt) {
      const int bx = (lane % tw) * 4, by = (lane / tw) * 4;
      const uint8_t *r = bi_umv(u, c.ref, qx + (bx << 2), qy + (by << 2));
      int d[16];
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) d[j * 4 + k] = (int)cur[(size_t)(by + j) * a.cur_pitch + bx + k] - (int)r[(size_t)j * a.Wp + k];
      s = bi_had4(d);
    }
  } else {
    const int tw = bsx >> 3, nt = tw * (bsy >> 3);
    if (lane < nt) {
      const int bx = (lane % tw) * 8, by = (lane / tw) * 8;
      const uint8_t *r = bi_umv(u, c.ref, qx + (bx << 2), qy + (by << 2));
      short d[64];
      for (int j = 0; j < 8; j++)
        for (int k = 0; k < 8; k++) d[j * 8 + k] = (short)((int)cur[(size_t)(by + j) * a.cur_pitch + bx + k] - (int)r[(size_t)j * a.Wp + k]);
      s = bi_had8(d);
    }
  }
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) a.out[i] = (long long)s << 5;
}

// ---- EPZS integer-pel search (JM/lencod/src/me_epzs.c:54-407 macroblock variant, :417-750 sub-macroblock variant) -------
// One warp per job.  The control flow is scalar (every lane runs the same state machine); a search point's SAD is computed
// by the 32 lanes together.  The reference's EPZSMap (one uint16 per window position, "visited in this call" = BlkCount) is a
// per-warp list of the visited vectors here: a call visits tens of points, and membership is one strided compare.
struct EpzsArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, nrefs, njobs, npats;
  const uint8_t *spl; int Wq, Hq, spad;          // the references' byte-shifted search planes (k_search_plane): aligned words at any column
  const b2me_epzs_job *jobs; const int16_t *preds; const b2me_epzs_pattern *pats; b2me_epzs_result *out; int *errflag;
};
constexpr int EPZS_VCAP = 768;        // visited vectors per job (overflow: the job fails loudly)
#ifndef EPZS_MINBV
#define EPZS_MINBV 5
#endif
constexpr int EPZS_MINB = EPZS_MINBV;   // resident CTAs per SM asked of ptxas.  Measured: 8 / 12 / 16 (64 / 40 / 32 registers, spills) give 0.89 / 1.49 / 2.01 ms against 0.89 at 5: the kernel is bound by its own instruction count (1 900 per job, 66 % issue), not by latency

// The visited set: vectors 0..63 live in two registers per lane (lane l holds vectors l and l + 32: a membership test is one
// compare per register and a vote, an insertion a predicated move), later ones in the shared-memory list.
struct EpzsVis { uint32_t v0, v1; uint32_t *list; int n; bool ovf; };
__device__ __forceinline__ bool epzs_visit(EpzsVis &V, int x, int y)
{
  const int lane = threadIdx.x & 31;
  const uint32_t key = ((uint32_t)(x & 0xffff) << 16) | (uint32_t)(y & 0xffff);
  bool hit = (lane < V.n && V.v0 == key) || (lane + 32 < V.n && V.v1 == key);
  for (int i = 64 + lane; i < V.n; i += 32) hit |= V.list[i] == key;
  if (__any_sync(0xffffffffu, hit)) return false;
  if (V.n >= EPZS_VCAP) { V.ovf = true; return true; }
  if (V.n < 32) { if (lane == V.n) V.v0 = key; }
  else if (V.n < 64) { if (lane == V.n - 32) V.v1 = key; }
  else { if (lane == 0) V.list[V.n] = key; __syncwarp(); }
  V.n++;
  return true;
}

__global__ void __launch_bounds__(128, EPZS_MINB) k_epzs(const EpzsArgs a)
{
  __shared__ uint32_t s_vis[4][EPZS_VCAP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, ji = blockIdx.x * 4 + w;
  if (ji >= a.njobs) return;
  const b2me_epzs_job J = a.jobs[ji];
  EpzsVis V; V.v0 = V.v1 = 0u; V.list = s_vis[w]; V.n = 0; V.ovf = false;
  bool &ovf = V.ovf;
  bool bad = J.blocktype < 1 || J.blocktype > 7 || J.ref < 0 || J.ref >= a.nrefs || J.pos_x < 0 || J.pos_y < 0;
  const PartGeom gm = part_geom(part_first(bad ? 1 : J.blocktype));
  const int bsx = gm.w, bsy = gm.h;
  bad = bad || J.pos_x + bsx > a.W || J.pos_y + bsy > a.H;
  const int pidx[5] = {J.pat_init, J.pat_sd, J.pat_sq, J.pat_else, J.pat_dual};
  for (int i = 0; i < 5; i++) bad = bad || pidx[i] < 0 || pidx[i] >= a.npats;
  if (bad) { if (lane == 0) { *a.errflag = 1; a.out[ji].cost = -1; a.out[ji].mv[0] = a.out[ji].mv[1] = 0; a.out[ji].early = 1; a.out[ji].npoints = 0; } return; }
  const uint8_t *cur = a.cur + (size_t)J.pos_y * a.cur_pitch + J.pos_x;
  BiArgs u; u.planes = a.planes; u.plane_size = a.plane_size; u.W = a.W; u.H = a.H; u.Wp = a.Wp;
  int npts = 0;
  // The block's current samples stay in registers as packed words (a lane owns word `lane` and `lane + 32` of the block's
  // bsx / 4 x bsy words); an integer-pel candidate inside the search plane is read as ALIGNED words of the byte-shifted plane
  // (column & 3) -- one VABSDIFF4 per four samples instead of two byte loads per sample.  The search plane is the
  // edge-replicated integer picture, i.e. exactly what UMVLine4X's origin clamp delivers (sad_fs.cu, exactness note).
  const int wpr = bsx >> 2, nw = wpr * bsy, lw = bsx == 16 ? 2 : (bsx == 8 ? 1 : 0);      // words per row (1, 2 or 4: a shift), words per block
  uint32_t cw[2] = {0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int k = lane + 32 * h;
    if (k < nw) cw[h] = *reinterpret_cast<const uint32_t *>(cur + (size_t)(k >> lw) * a.cur_pitch + 4 * (k & (wpr - 1)));
  }
  const bool cur_aligned = ((J.pos_x | a.cur_pitch) & 3) == 0;
  auto sad = [&](int tx, int ty) -> long long {                  // computeSAD << 5 (me_distortion.c:349-426), no early exit
    if (cur_aligned && !((tx | ty) & 3)) {
      const int X = J.pos_x + (tx >> 2) + a.spad, Y = J.pos_y + (ty >> 2) + a.spad;
      if (X >= 0 && Y >= 0 && X + bsx <= a.Wq && Y + bsy <= a.Hq) {
        const int c = X & 3;
        const uint8_t *pl = a.spl + ((size_t)(J.ref * 16 + c) * a.Hq + Y) * a.Wq + (X - c);
        uint32_t s = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k = lane + 32 * h;
          if (k < nw) s = sad4(cw[h], __ldg(reinterpret_cast<const uint32_t *>(pl + (size_t)(k >> lw) * a.Wq + 4 * (k & (wpr - 1)))), s);
        }
        npts++;
        return (long long)__reduce_add_sync(0xffffffffu, s) << 5;
      }
    }
    const uint8_t *r = bi_umv(u, J.ref, (J.pos_x << 2) + tx, (J.pos_y << 2) + ty);
    int s = 0;
    for (int k = lane; k < bsx * bsy; k += 32) {
      const int x = k % bsx, y = k / bsx;
      s += abs((int)cur[(size_t)y * a.cur_pitch + x] - (int)r[(size_t)y * a.Wp + x]);
    }
    npts++;
    return (long long)__reduce_add_sync(0xffffffffu, s) << 5;
  };
  auto mvc = [&](int tx, int ty) -> long long { return (long long)J.lambda_factor * (mvbits(tx - J.pred[0]) + mvbits(ty - J.pred[1])); };
  auto inrange = [&](int tx, int ty) -> bool { return abs(tx - J.mv[0]) <= J.range[0] && abs(ty - J.mv[1]) <= J.range[1]; };
  const int mvx = J.mv[0], mvy = J.mv[1];
  int tx = mvx, ty = mvy, t2x = 0, t2y = 0;                       // tmp (best), tmp2 (second best)
  int early = 1;
  const bool pgate = (J.flags & B2ME_EPZS_REFGT0_FRAME) != 0;
  epzs_visit(V, mvx, mvy);
  long long minc = mvc(mvx, mvy);
  minc += sad(mvx, mvy);
  bool done = pgate && J.prev_sad < (J.stop0 < minc ? J.stop0 : minc);            // :118
  if (!done && minc > J.stop0) {
    const long long stop = J.stop;
    if (minc < (stop >> 1)) done = true;                                          // :145
    if (!done) {
      bool checkMedian = false;
      long long second = BI_DISTBLK_MAX;
      const bool use[4] = {true, J.cond_host[1] && minc > stop, J.fixed_edge || (J.cond_host[2] && minc > 3 * stop),
                           (J.cond_host[3] & 1) && ((J.cond_host[3] & 2) || minc > 2 * stop)};
      // Predictor scan.  The raw predictors are fetched 32 at a time, one per lane, and handed round by shuffles.
      // EPZS_NBV > 0 (measured, not the default): which predictors are looked at (group enabled, in range, not visited yet) does
      // not depend on any distortion, so the scan can run in batches -- up to EPZS_NB admitted predictors collected, the
      // reference words of ALL of them requested before the first SAD is formed, the reference's decisions replayed in order on
      // the finished sums (a predictor whose vector cost is not below `second` at its turn is skipped as the reference skips
      // it).  1080p, 41 partitions, five predictors: 0.80 ms one at a time, 0.86 ms in batches of 4, 1.08 ms in batches of 8
      // (128 registers): the kernel is bound by its instruction count, the batches' bookkeeping costs more than the latency.
#ifndef EPZS_NBV
#define EPZS_NBV 0
#endif
      constexpr int EPZS_NB = EPZS_NBV > 0 ? EPZS_NBV : 1;      // EPZS_NBV == 0: no batches, one predictor at a time (still prefetched 32 at a time)
      const int ntot = J.npred[0] + J.npred[1] + J.npred[2] + J.npred[3];
      uint32_t pword = 0;
      int bx_[EPZS_NB], by_[EPZS_NB], nb = 0;
      for (int idx = 0; idx <= ntot; idx++) {
        bool admit = false; int px = 0, py = 0;
        if (idx < ntot) {
          if ((idx & 31) == 0) pword = (idx + lane < ntot) ? reinterpret_cast<const uint32_t *>(a.preds)[J.pred_first + idx + lane] : 0u;
          const uint32_t wv = __shfl_sync(0xffffffffu, pword, idx & 31);
          const int g = idx < J.npred[0] ? 0 : (idx < J.npred[0] + J.npred[1] ? 1 : (idx < J.npred[0] + J.npred[1] + J.npred[2] ? 2 : 3));
          px = (int)(short)((wv & 0xffffu) & 0xFFFCu); py = (int)(short)((wv >> 16) & 0xFFFCu);                       // set_integer_mv
          admit = use[g] && inrange(px, py) && epzs_visit(V, px, py);
          if (EPZS_NBV == 0) {
            if (admit) {
              long long mc = mvc(px, py);
              if (mc < second) {
                mc += sad(px, py);
                if (mc < minc) { t2x = tx; t2y = ty; tx = px; ty = py; second = minc; minc = mc; checkMedian = true; }
                else if (mc < second) { t2x = px; t2y = py; second = mc; checkMedian = true; }
              }
            }
            continue;
          }
          if (admit) {
#pragma unroll
            for (int e = 0; e < EPZS_NB; e++) if (e == nb) { bx_[e] = px; by_[e] = py; }
            nb++;
          }
        }
        if (nb == EPZS_NB || (idx == ntot && nb > 0)) {
          // ---- all reference words of the batch, then the sums ----
          long long bs_[EPZS_NB];
          uint32_t rw0[EPZS_NB], rw1[EPZS_NB];
          bool fast[EPZS_NB];
#pragma unroll
          for (int e = 0; e < EPZS_NB; e++) {
            fast[e] = false; rw0[e] = rw1[e] = 0u;
            if (e < nb && cur_aligned) {
              const int X = J.pos_x + (bx_[e] >> 2) + a.spad, Y = J.pos_y + (by_[e] >> 2) + a.spad;
              if (X >= 0 && Y >= 0 && X + bsx <= a.Wq && Y + bsy <= a.Hq) {
                fast[e] = true;
                const int c = X & 3;
                const uint8_t *pl = a.spl + ((size_t)(J.ref * 16 + c) * a.Hq + Y) * a.Wq + (X - c);
                if (lane < nw) rw0[e] = __ldg(reinterpret_cast<const uint32_t *>(pl + (size_t)(lane >> lw) * a.Wq + 4 * (lane & (wpr - 1))));
                if (lane + 32 < nw) rw1[e] = __ldg(reinterpret_cast<const uint32_t *>(pl + (size_t)((lane + 32) >> lw) * a.Wq + 4 * ((lane + 32) & (wpr - 1))));
              }
            }
          }
#pragma unroll
          for (int e = 0; e < EPZS_NB; e++) {
            bs_[e] = 0;
            if (e < nb && fast[e]) {
              uint32_t sv = lane < nw ? sad4(cw[0], rw0[e], 0u) : 0u;
              if (lane + 32 < nw) sv = sad4(cw[1], rw1[e], sv);
              bs_[e] = (long long)__reduce_add_sync(0xffffffffu, sv) << 5;
            }
          }
          // ---- the reference's decisions, in order ----
#pragma unroll
          for (int e = 0; e < EPZS_NB; e++) {
            if (e >= nb) continue;
            long long mc = mvc(bx_[e], by_[e]);
            if (mc < second) {
              if (fast[e]) { mc += bs_[e]; npts++; } else mc += sad(bx_[e], by_[e]);
              if (mc < minc) { t2x = tx; t2y = ty; tx = bx_[e]; ty = by_[e]; second = minc; minc = mc; checkMedian = true; }
              else if (mc < second) { t2x = bx_[e]; t2y = by_[e]; second = mc; checkMedian = true; }
            }
          }
          nb = 0;
        }
      }
      if ((J.flags & B2ME_EPZS_EARLY34) && minc < ((3 * stop) >> 2)) done = true;  // :586 (sub-macroblock variant)
      if (!done && minc > stop) {
        int pat = J.pat_init;
        if (J.flags & B2ME_EPZS_ADAPT) {
          if (minc < stop + ((3 * J.medthres) >> 1))
            pat = ((tx == 0 && ty == 0) || (abs(tx - mvx) < J.mv_range && abs(ty - mvy) < J.mv_range)) ? J.pat_sd : J.pat_sq;
          else pat = J.pat_else;
        }
        int cx = tx, cy = ty;
        int patternStop = 0, pointNumber = 0, nextLast = 0, motionDirection = 0;
        for (int guard = 0; guard < 4 && !ovf; guard++) {
          const b2me_epzs_pattern *P = &a.pats[pat];
          int totalCheckPts = P->npoints;
          int iter = 0;
          do {
            int checkPts = totalCheckPts;
            do {
              const int qx = cx + P->pt[pointNumber].dx, qy = cy + P->pt[pointNumber].dy;
              if (inrange(qx, qy) && epzs_visit(V, qx, qy)) {
                long long mc = mvc(qx, qy);
                if (mc < minc) {
                  mc += sad(qx, qy);
                  if (mc < minc) { tx = qx; ty = qy; minc = mc; motionDirection = pointNumber; }
                }
              }
              ++pointNumber;
              if (pointNumber >= P->npoints) pointNumber -= P->npoints;
              checkPts--;
            } while (checkPts > 0);
            if (nextLast || (tx == cx && ty == cy)) {
              patternStop = P->stop_search;
              const int np = P->next_pattern;
              if (np < 0 || np >= a.npats) { ovf = true; break; }
              P = &a.pats[np];
              totalCheckPts = P->npoints;
              nextLast = P->next_last;
              motionDirection = 0;
              pointNumber = 0;
            } else {
              totalCheckPts = P->pt[motionDirection].next_points;
              pointNumber = P->pt[motionDirection].start_nmbr;
              cx = tx; cy = ty;
            }
            if (++iter > 4096) { ovf = true; break; }
          } while (patternStop != 1);
          if (ovf) break;
          if (pgate && ((4 * J.prev_sad < minc) || ((3 * J.prev_sad < minc) && (J.prev_sad <= stop)))) { done = true; break; }   // :351
          if (!(checkMedian && (J.flags & B2ME_EPZS_DUAL) && minc > stop)) break;
          pointNumber = 0; patternStop = 0; motionDirection = 0; nextLast = 0;
          if ((tx == 0 && ty == 0) || (tx == mvx && ty == mvy))
            pat = (abs(tx - mvx) < J.mv_range && abs(ty - mvy) < J.mv_range) ? J.pat_sd : J.pat_sq;
          else pat = J.pat_dual;
          cx = t2x; cy = t2y;
          checkMedian = false;
        }
      }
    }
  }
  if (!done) early = 0;
  if (lane == 0) {
    if (ovf) { *a.errflag = 2; a.out[ji].cost = -1; }
    else a.out[ji].cost = minc;
    a.out[ji].mv[0] = (int16_t)tx; a.out[ji].mv[1] = (int16_t)ty; a.out[ji].early = (int16_t)early; a.out[ji].npoints = (int16_t)npts;
  }
}

// ---- list_prediction_cost, list 0 (JM/lencod/src/mode_decision.c:275-300; update_mcost :256-267; ref_cost mv_search.h:114) ---
// For the 21 (mode, block) entries of every macroblock: the reference minimising motion cost + lambda * refbits(ref),
// first minimum in reference order; the motion cost of an 8x8 quadrant in modes 5..7 is the sum over its sub-partitions
// (PartitionMotionSearch, mv_search.c:1601-1843).  One thread per (macroblock, entry); HBM-bound (nrefs * 41 * 8 B in per MB).
__constant__ signed char c_entry_parts[21][4] = {
  {0, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {3, -1, -1, -1}, {4, -1, -1, -1},
  {5, -1, -1, -1}, {6, -1, -1, -1}, {7, -1, -1, -1}, {8, -1, -1, -1},
  {9, 11, -1, -1}, {10, 12, -1, -1}, {13, 15, -1, -1}, {14, 16, -1, -1},
  {17, 18, -1, -1}, {19, 20, -1, -1}, {21, 22, -1, -1}, {23, 24, -1, -1},
  {25, 26, 29, 30}, {27, 28, 31, 32}, {33, 34, 37, 38}, {35, 36, 39, 40}};

// nrefs: references per macroblock in the cost array; lsize: listXsize of the list the costs belong to (the loop bound of
// list_prediction_cost and the "<= 1 reference: no reference bits" rule of ref_cost) -- list 1 of a B slice usually holds fewer.
__global__ void __launch_bounds__(256) k_select_refs(int nmb, int nrefs, int lsize, const long long *__restrict__ cost, int ref_lambda,
                                                      int8_t *__restrict__ best_ref, long long *__restrict__ best_cost)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmb * 21) return;
  const int mb = i / 21, e = i - mb * 21;
  long long bm = BI_DISTBLK_MAX; int br = 0;
  for (int r = 0; r < lsize; r++) {
    long long mc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int p = c_entry_parts[e][k]; if (p >= 0) mc += cost[((size_t)mb * nrefs + r) * NPART + p]; }
    if (mc < bm) {
      mc += lsize <= 1 ? 0 : (long long)ref_lambda * (65 - 2 * __clz(r + 1) - 2);    // refbits = 2 * floor(log2(ref + 1)) + 1
      if (mc < bm) { bm = mc; br = r; }
    }
  }
  best_ref[i] = (int8_t)br; best_cost[i] = bm;
}

// ---- compact frame search (b2me_search_frame_best): one predictor per (MB, ref) in, the mode decision's view out -------------
// pred_mb [nmb][nrefs][2] -> pred / centre [nmb][nrefs][41][2]: the 41 partitions share the predictor, the integer search centre
// is ((p + 2) >> 2) * 4 (JM_INT_DIVIDE, mv_search.c:931-932)
__global__ void __launch_bounds__(256) k_expand_pred(int n, const int16_t *__restrict__ pred_mb, int16_t *__restrict__ pred, int16_t *__restrict__ center)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (mb * nrefs + ref) * 41 + p
  if (i >= n) return;
  const int item = i / NPART;
  const int px = pred_mb[2 * item], py = pred_mb[2 * item + 1];
  pred[2 * i] = (int16_t)px; pred[2 * i + 1] = (int16_t)py;
  center[2 * i] = (int16_t)(((px + 2) >> 2) * 4); center[2 * i + 1] = (int16_t)(((py + 2) >> 2) * 4);
}

// the vector of every partition for the reference its (mode, block) entry chose; costs saturated to int32
__global__ void __launch_bounds__(256) k_gather_best(int nmb, int nrefs, const int16_t *__restrict__ mv, const int8_t *__restrict__ best_ref,
                                                     const long long *__restrict__ best_cost, int16_t *__restrict__ best_mv, int32_t *__restrict__ cost32)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nmb * 21) { const long long c = best_cost[i]; cost32[i] = c > 0x7fffffffll ? 0x7fffffff : (int32_t)c; }
  if (i >= nmb * NPART) return;
  const int mb = i / NPART, p = i - mb * NPART;
  int e = 0;
#pragma unroll 1
  for (int k = 0; k < 21; k++)
    for (int j = 0; j < 4; j++) if (c_entry_parts[k][j] == p) e = k;
  const int r = best_ref[mb * 21 + e];
  const size_t src = (((size_t)mb * nrefs + r) * NPART + p) * 2;
  best_mv[2 * i] = mv[src]; best_mv[2 * i + 1] = mv[src + 1];
}

cudaError_t launch_expand_pred(int n, const int16_t *pred_mb, int16_t *pred, int16_t *center, cudaStream_t s)
{
  k_expand_pred<<<(n + 255) / 256, 256, 0, s>>>(n, pred_mb, pred, center);
  return cudaGetLastError();
}
cudaError_t launch_select_gather(int nmb, int nrefs, const long long *cost, int ref_lambda, const int16_t *mv, int8_t *best_ref, long long *best_cost,
                                 int16_t *best_mv, int32_t *cost32, cudaStream_t s)
{
  k_select_refs<<<(nmb * 21 + 255) / 256, 256, 0, s>>>(nmb, nrefs, nrefs, cost, ref_lambda, best_ref, best_cost);
  k_gather_best<<<(nmb * NPART + 255) / 256, 256, 0, s>>>(nmb, nrefs, mv, best_ref, best_cost, best_mv, cost32);
  return cudaGetLastError();
}

}  // namespace b2

using namespace b2;

extern "C" int b2me_select_refs_list_dev(b2me_ctx *c, const int64_t *cost_dev, int list_size, int ref_lambda, int8_t *best_ref_dev, int64_t *best_cost_dev, void *stream)
{
  if (!c || !cost_dev || !best_ref_dev || !best_cost_dev || list_size < 1 || list_size > c->nrefs) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  const int n = c->nmb * 21;
  k_select_refs<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->nmb, c->nrefs, list_size, reinterpret_cast<const long long *>(cost_dev), ref_lambda,
                                                                  best_ref_dev, reinterpret_cast<long long *>(best_cost_dev));
  B2_CUDA_CHECK(c, cudaGetLastError());
  c->launches++;
  return B2ME_OK;
}

extern "C" int b2me_select_refs_dev(b2me_ctx *c, const int64_t *cost_dev, int ref_lambda, int8_t *best_ref_dev, int64_t *best_cost_dev, void *stream)
{
  if (!c || !cost_dev || !best_ref_dev || !best_cost_dev) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  const int n = c->nmb * 21;
  k_select_refs<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->nmb, c->nrefs, c->nrefs, reinterpret_cast<const long long *>(cost_dev), ref_lambda,
                                                                  best_ref_dev, reinterpret_cast<long long *>(best_cost_dev));
  B2_CUDA_CHECK(c, cudaGetLastError());
  c->launches++;
  return B2ME_OK;
}

extern "C" int b2me_distortion_candidates_dev(b2me_ctx *c, int metric, int test8x8, int n, const b2me_candidate *cands_dev,
                                              int64_t *out_dev, void *stream)
{
  if (!c || n < 0 || (n && (!cands_dev || !out_dev)) || metric < 0 || metric > 2) return B2ME_EINVAL;
  if (!n) return B2ME_OK;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (c->planes_pending && s != c->stream) B2_CUDA_CHECK(c, cudaStreamWaitEvent(s, c->ev_planes, 0));
  CandArgs a;
  a.cur = c->d_cur; a.cur_pitch = c->W; a.planes = c->d_planes; a.plane_size = c->plane_size;
  a.W = c->W; a.H = c->H; a.Wp = c->Wp; a.nrefs = c->nrefs; a.metric = metric; a.test8x8 = test8x8; a.n = n;
  a.cands = cands_dev; a.out = reinterpret_cast<long long *>(out_dev); a.errflag = c->d_errflag;
  k_cand_dist<<<(n + 3) / 4, 128, 0, s>>>(a);
  B2_CUDA_CHECK(c, cudaGetLastError());
  c->launches++;
  return B2ME_OK;
}

extern "C" int b2me_distortion_candidates(b2me_ctx *c, int metric, int test8x8, int n, const b2me_candidate *cands, int64_t *out)
{
  if (!c || n < 0 || (n && (!cands || !out)) || metric < 0 || metric > 2) return B2ME_EINVAL;
  if (!n) return B2ME_OK;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  b2me_candidate *dc = nullptr; int64_t *dout = nullptr;
  B2_CUDA_CHECK(c, cudaMallocAsync(&dc, sizeof(b2me_candidate) * n, c->stream));
  B2_CUDA_CHECK(c, cudaMallocAsync(&dout, sizeof(int64_t) * n, c->stream));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(dc, cands, sizeof(b2me_candidate) * n, cudaMemcpyHostToDevice, c->stream));
  int r = b2me_distortion_candidates_dev(c, metric, test8x8, n, dc, dout, c->stream);
  if (!r) {
    cudaError_t e = cudaMemcpyAsync(out, dout, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, c->stream);
    int flag = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "b2me_distortion_candidates: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
    else if (flag) {
      cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream);
      snprintf(c->err, sizeof(c->err), "b2me_distortion_candidates: a candidate is out of range (blocktype, reference slot or position)");
      r = B2ME_EINVAL;
    }
  }
  cudaFreeAsync(dc, c->stream); cudaFreeAsync(dout, c->stream);
  return r;
}

extern "C" int b2me_epzs_search_dev(b2me_ctx *c, int njobs, const b2me_epzs_job *jobs_dev, const int16_t *preds_dev,
                                    int npatterns, const b2me_epzs_pattern *patterns_dev, b2me_epzs_result *out_dev, void *stream)
{
  if (!c || njobs < 0 || npatterns < 1 || (njobs && (!jobs_dev || !patterns_dev || !out_dev))) return B2ME_EINVAL;
  if (!njobs) return B2ME_OK;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (c->planes_pending && s != c->stream) B2_CUDA_CHECK(c, cudaStreamWaitEvent(s, c->ev_planes, 0));
  EpzsArgs a;
  a.cur = c->d_cur; a.cur_pitch = c->W; a.planes = c->d_planes; a.plane_size = c->plane_size;
  a.W = c->W; a.H = c->H; a.Wp = c->Wp; a.nrefs = c->nrefs; a.njobs = njobs; a.npats = npatterns;
  a.jobs = jobs_dev; a.preds = preds_dev; a.pats = patterns_dev; a.out = out_dev; a.errflag = c->d_errflag;
  a.spl = c->d_spl; a.Wq = c->Wq; a.Hq = c->Hq; a.spad = c->spad;
  k_epzs<<<(njobs + 3) / 4, 128, 0, s>>>(a);
  B2_CUDA_CHECK(c, cudaGetLastError());
  c->launches++;
  return B2ME_OK;
}

extern "C" int b2me_epzs_search(b2me_ctx *c, int njobs, const b2me_epzs_job *jobs, int npreds_total, const int16_t *preds,
                                int npatterns, const b2me_epzs_pattern *patterns, b2me_epzs_result *out)
{
  if (!c || njobs < 0 || npreds_total < 0 || npatterns < 1 || (njobs && (!jobs || !patterns || !out)) || (npreds_total && !preds)) return B2ME_EINVAL;
  if (!njobs) return B2ME_OK;
  for (int i = 0; i < njobs; i++) {                   // every job's predictors lie inside the array handed over
    const int n = jobs[i].npred[0] + jobs[i].npred[1] + jobs[i].npred[2] + jobs[i].npred[3];
    if (jobs[i].pred_first < 0 || n < 0 || jobs[i].pred_first + n > npreds_total) return B2ME_EINVAL;
  }
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  const size_t bj = sizeof(b2me_epzs_job) * njobs, bp = ((sizeof(int16_t) * 2 * (size_t)npreds_total + 15) & ~(size_t)15) + 16;
  const size_t bt = sizeof(b2me_epzs_pattern) * npatterns, bo = sizeof(b2me_epzs_result) * njobs;
  uint8_t *d = nullptr;
  B2_CUDA_CHECK(c, cudaMallocAsync(&d, bj + bp + bt + bo, c->stream));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(d, jobs, bj, cudaMemcpyHostToDevice, c->stream));
  if (npreds_total) B2_CUDA_CHECK(c, cudaMemcpyAsync(d + bj, preds, sizeof(int16_t) * 2 * (size_t)npreds_total, cudaMemcpyHostToDevice, c->stream));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(d + bj + bp, patterns, bt, cudaMemcpyHostToDevice, c->stream));
  int r = b2me_epzs_search_dev(c, njobs, reinterpret_cast<const b2me_epzs_job *>(d), reinterpret_cast<const int16_t *>(d + bj), npatterns,
                               reinterpret_cast<const b2me_epzs_pattern *>(d + bj + bp), reinterpret_cast<b2me_epzs_result *>(d + bj + bp + bt), c->stream);
  if (!r) {
    int flag = 0;
    cudaError_t e = cudaMemcpyAsync(out, d + bj + bp + bt, bo, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "b2me_epzs_search: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
    else if (flag) {
      cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream);
      snprintf(c->err, sizeof(c->err), flag == 2 ? "b2me_epzs_search: a job visited more search points than the device list holds, or names a pattern out of range"
                                                 : "b2me_epzs_search: a job is out of range (blocktype, reference slot, position or pattern index)");
      r = flag == 2 ? B2ME_EUNSUPPORTED : B2ME_EINVAL;
    }
  }
  cudaFreeAsync(d, c->stream);
  return r;
}

extern "C" int b2me_bipred_distortion_candidates(b2me_ctx *c, int metric, int test8x8, int apply_weights, int luma_log_weight_denom, int n,
                                                 const b2me_bipred_job *cands, int64_t *out)
{
  if (!c || n < 0 || (n && (!cands || !out)) || metric < 0 || metric > 2 || luma_log_weight_denom < 0 || luma_log_weight_denom > 7) return B2ME_EINVAL;
  if (apply_weights && test8x8 && metric == 2) {
    snprintf(c->err, sizeof(c->err), "weighted bi-predictive SATD with the 8x8 Hadamard is not implemented (the reference's computeBiPredSATD2 8x8 branch "
                                     "reads past its source row, me_distortion.c:1167)");
    return B2ME_EUNSUPPORTED;
  }
  if (!n) return B2ME_OK;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  b2me_bipred_job *dj = nullptr; long long *dout = nullptr;
  B2_CUDA_CHECK(c, cudaMallocAsync(&dj, sizeof(b2me_bipred_job) * n, c->stream));
  B2_CUDA_CHECK(c, cudaMallocAsync(&dout, sizeof(long long) * n, c->stream));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(dj, cands, sizeof(b2me_bipred_job) * n, cudaMemcpyHostToDevice, c->stream));
  BiArgs a;
  memset(&a, 0, sizeof(a));
  a.cur = c->d_cur; a.cur_pitch = c->W; a.planes = c->d_planes; a.plane_size = c->plane_size;
  a.W = c->W; a.H = c->H; a.Wp = c->Wp; a.nrefs = c->nrefs; a.R = c->R; a.test8x8 = test8x8; a.wp = apply_weights; a.denom = luma_log_weight_denom;
  a.jobs = dj; a.njobs = n; a.errflag = c->d_errflag;
  k_bicand<<<n, 32, 0, c->stream>>>(a, metric, dout);
  int r = B2ME_OK, flag = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, sizeof(long long) * n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "b2me_bipred_distortion_candidates: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  else if (flag) {
    cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream);
    snprintf(c->err, sizeof(c->err), "b2me_bipred_distortion_candidates: a record is out of range (blocktype, reference slot or position)");
    r = B2ME_EINVAL;
  }
  c->launches++;
  cudaFreeAsync(dj, c->stream); cudaFreeAsync(dout, c->stream);
  return r;
}

static int bipred_check(b2me_ctx *c, int njobs, const void *jobs, const b2me_search_params *p, int apply_weights, int denom, int test8x8, const void *out)
{
  if (!c || njobs < 0 || (njobs && (!jobs || !out)) || !p) return B2ME_EINVAL;
  if (p->do_subpel && (p->metric_h < 0 || p->metric_h > 2 || p->metric_q < 0 || p->metric_q > 2)) return B2ME_EINVAL;
  if (denom < 0 || denom > 7) return B2ME_EINVAL;
  if (apply_weights && test8x8 && p->do_subpel && (p->metric_h == 2 || p->metric_q == 2)) {
    snprintf(c->err, sizeof(c->err), "weighted bi-predictive SATD with the 8x8 Hadamard is not implemented (the reference's "
                                     "computeBiPredSATD2 8x8 branch reads past its source row, me_distortion.c:1167)");
    return B2ME_EUNSUPPORTED;
  }
  return B2ME_OK;
}

extern "C" int b2me_bipred_search_dev(b2me_ctx *c, int njobs, const b2me_bipred_job *jobs_dev, const b2me_search_params *p,
                                      int apply_weights, int luma_log_weight_denom, int test8x8, b2me_bipred_result *out_dev, void *stream)
{
  int r = bipred_check(c, njobs, jobs_dev, p, apply_weights, luma_log_weight_denom, test8x8, out_dev);
  if (r || !njobs) return r;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (c->planes_pending && s != c->stream) B2_CUDA_CHECK(c, cudaStreamWaitEvent(s, c->ev_planes, 0));
  BiArgs a;
  a.cur = c->d_cur; a.cur_pitch = c->W; a.planes = c->d_planes; a.plane_size = c->plane_size;
  a.W = c->W; a.H = c->H; a.Wp = c->Wp; a.nrefs = c->nrefs; a.R = c->R;
  a.lambda[0] = p->lambda_factor[0]; a.lambda[1] = p->lambda_factor[1]; a.lambda[2] = p->lambda_factor[2];
  a.metric_h = p->metric_h; a.metric_q = p->metric_q; a.do_subpel = p->do_subpel; a.full81 = p->subpel_full ? 1 : 0; a.test8x8 = test8x8;
  a.wp = apply_weights; a.denom = luma_log_weight_denom; a.jobs = jobs_dev; a.out = out_dev; a.njobs = njobs; a.errflag = c->d_errflag;
  k_bipred<<<njobs, 128, 0, s>>>(a);
  B2_CUDA_CHECK(c, cudaGetLastError());
  c->launches++;
  return B2ME_OK;
}

extern "C" int b2me_bipred_search(b2me_ctx *c, int njobs, const b2me_bipred_job *jobs, const b2me_search_params *p,
                                  int apply_weights, int luma_log_weight_denom, int test8x8, b2me_bipred_result *out)
{
  int r = bipred_check(c, njobs, jobs, p, apply_weights, luma_log_weight_denom, test8x8, out);
  if (r || !njobs) return r;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  b2me_bipred_job *dj = nullptr; b2me_bipred_result *dr = nullptr;
  B2_CUDA_CHECK(c, cudaMallocAsync(&dj, sizeof(b2me_bipred_job) * njobs, c->stream));
  B2_CUDA_CHECK(c, cudaMallocAsync(&dr, sizeof(b2me_bipred_result) * njobs, c->stream));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(dj, jobs, sizeof(b2me_bipred_job) * njobs, cudaMemcpyHostToDevice, c->stream));
  r = b2me_bipred_search_dev(c, njobs, dj, p, apply_weights, luma_log_weight_denom, test8x8, dr, c->stream);
  if (!r) {
    cudaError_t e = cudaMemcpyAsync(out, dr, sizeof(b2me_bipred_result) * njobs, cudaMemcpyDeviceToHost, c->stream);
    int flag = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "b2me_bipred_search: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
    else if (flag) {
      cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream);
      snprintf(c->err, sizeof(c->err), "b2me_bipred_search: a job is out of range (blocktype, reference slot, position, search range or a sub-pel centre)");
      r = B2ME_EINVAL;
    }
  }
  cudaFreeAsync(dj, c->stream); cudaFreeAsync(dr, c->stream);
  return r;
}
