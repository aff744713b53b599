// subpel_refine.cu -- half- and quarter-pel refinement of all 41 partitions of a macroblock.
//
// Replaces, per (macroblock, reference): 41 calls of sub_pel_motion_estimation
// (JM/lencod/src/me_fullsearch.c:186-289) with computeSATD (me_distortion.c:745-825,
// HadamardSAD4x4 :175-258) or computeSAD (:349-426) as computePredHPel/QPel.
//
// One warp per (MB, ref).  A lane-job is one 4x4 tile of the macroblock and one candidate index
// (16 tiles x 9 half-pel, then 8 quarter-pel candidates) and walks the 7 block types whose partition
// contains the tile; partitions with equal motion vectors ask for the same reference tile, which is
// then computed once.  The lane loads the 4x4 reference tile from the quarter-pel plane
// [y&3][x&3] with the reference's own tile-origin clamp (UMVLine4X, refbuf.h:22-26), forms the
// difference against the current MB in shared memory, runs the 4x4 Hadamard in registers (tile_satd: dp4a rows) and
// adds (satd+1)>>1 into each partition's candidate accumulator.  When every active partition of the item carries the
// same vector (the common case) the tile depends on (tile, candidate) only: lane = (tile, candidate parity) and the
// 41 partition sums are formed through the block-size tree in shuffles (no atomics, no block-type walk).  41 threads then take the lexicographic (cost, position) minimum in
// spiral order with the reference's carried / reset min_mcost rules (mv_search.c:971-974,
// me_fullsearch.c:252-253).
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

constexpr long long DMAX = ((long long)0x7fffffff) << 5;

// spiral positions 0..8 (x,y) in units of the step (mv_search.c:406-442 with l = 1)
__constant__ signed char c_sp9[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

// spiral position c (0..80) -> (x, y); the first nine come from packed 2-bit tables (the hot case)
__device__ __forceinline__ void sp_xy(int c, int *x, int *y)
{
  if (c < 9) { *x = (int)((0x22215u >> (2 * c)) & 3u) - 1; *y = (int)((0x29421u >> (2 * c)) & 3u) - 1; }
  else spiral_xy(c, x, y);
}

// dot product of four unsigned bytes (a) with four signed bytes (b), accumulated into c
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ int hadamard4x4_abs(const int d[16])
{
  int m[16], s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
    const int p = a + b, q = a - b, r = c + e, t = c - e;
    m[4 * i] = p + r; m[4 * i + 1] = q + t; m[4 * i + 2] = p - r; m[4 * i + 3] = q - t;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = m[i], b = m[4 + i], c = m[8 + i], e = m[12 + i];
    const int p = a + b, q = a - b, r = c + e, t = c - e;
    s += abs(p + r) + abs(q + t) + abs(p - r) + abs(q - t);
  }
  return (s + 1) >> 1;
}

// partition of block type index bti (0..6) containing the 4x4 tile at (tx, ty)
__device__ __forceinline__ int part_of_tile(int bti, int tx, int ty)
{
  switch (bti) {
    case 0: return 0;
    case 1: return 1 + (ty >> 3);
    case 2: return 3 + (tx >> 3);
    case 3: return 5 + (ty >> 3) * 2 + (tx >> 3);
    case 4: return 9 + (ty >> 2) * 2 + (tx >> 3);
    case 5: return 17 + (ty >> 3) * 4 + (tx >> 2);
    default: return 25 + (ty >> 2) * 4 + (tx >> 2);
  }
}

// HadamardSAD4x4 (me_distortion.c:175-258) of the current tile rows cw[0..3] against the reference tile at rp, without
// unpacking a byte: a row's four horizontal coefficients are dot products of its packed bytes with +-1 patterns (dp4a,
// u8 x s8, on the FMA pipe: current row with +h, reference row with -h chained into the same accumulator); the vertical
// 4-point transform then works on scalars, and its last butterfly never materialises: |u + v| + |u - v| = 2 max(|u|, |v|),
// so (sum + 1) >> 1 is the sum of max(|u|, |v|) over the pairs.
__device__ __forceinline__ int tile_satd(const uint32_t (&cw)[4], const uint8_t *rp, int Wp)
{
  const int al = (int)(reinterpret_cast<size_t>(rp) & 3);      // two aligned words per row instead of four byte loads
  int R[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rp + (size_t)r * Wp - al);
    const uint32_t lo = w[0], hi = al ? w[1] : 0u;
    const uint32_t px = al ? __funnelshift_r(lo, hi, 8 * al) : lo;
    R[r][0] = dp4a_us(cw[r], 0x01010101u, dp4a_us(px, 0xFFFFFFFFu, 0));
    R[r][1] = dp4a_us(cw[r], 0xFF01FF01u, dp4a_us(px, 0x01FF01FFu, 0));
    R[r][2] = dp4a_us(cw[r], 0xFFFF0101u, dp4a_us(px, 0x0101FFFFu, 0));
    R[r][3] = dp4a_us(cw[r], 0x01FFFF01u, dp4a_us(px, 0xFF0101FFu, 0));
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int p = R[0][k] + R[1][k], q = R[0][k] - R[1][k], r = R[2][k] + R[3][k], t = R[2][k] - R[3][k];
    s += max(abs(p), abs(r)) + max(abs(q), abs(t));
  }
  return s;
}

// distortion of the 4x4 tile of the current MB at (tx, ty) against the reference tile `rp` (row pitch Wp)
__device__ __forceinline__ int tile_distortion(const uint8_t *cur, int tx, int ty, const uint8_t *rp, int Wp, int metric)
{
  const int al = (int)(reinterpret_cast<size_t>(rp) & 3);      // two aligned words per row instead of four byte loads
  if (metric == 2) {
    uint32_t cw[4];
#pragma unroll
    for (int r = 0; r < 4; r++) cw[r] = *reinterpret_cast<const uint32_t *>(&cur[(ty + r) * 16 + tx]);
    return tile_satd(cw, rp, Wp);
  }
  int d[16];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rp + (size_t)r * Wp - al);
    const uint32_t lo = w[0], hi = al ? w[1] : 0u;
    const uint32_t px = al ? __funnelshift_r(lo, hi, 8 * al) : lo;
    const uint32_t cw = *reinterpret_cast<const uint32_t *>(&cur[(ty + r) * 16 + tx]);
#pragma unroll
    for (int q = 0; q < 4; q++) d[r * 4 + q] = (int)((cw >> (8 * q)) & 255u) - (int)((px >> (8 * q)) & 255u);
  }
  if (metric == 2) return hadamard4x4_abs(d);
  int v = 0;
#pragma unroll
  for (int q = 0; q < 16; q++) v += metric == 1 ? d[q] * d[q] : abs(d[q]);      // SSE (computeSSE, me_distortion.c:1190) / SAD
  return v;
}

// One WARP per (MB, ref) item: nothing in the refinement of an item needs more than warp-level synchronisation,
// and CTA barriers between its short phases were the dominant stall.  NC = candidates per stage (9; 81 for
// full_sub_pel_motion_estimation); SP_WPC warps (items) per CTA.
template <int NC>
struct SpWarp {
  alignas(16) uint8_t cur[256];
  int dist[NPART][NC];
  short mv[NPART][2], prd[NPART][2];
  long long mincost[NPART];
};

template <int NC, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_subpel_refine(const SubArgs a)
{
  // Lane-job (k, c) = 4x4 tile k of the macroblock, candidate index c.  The seven partitions (one per block type)
  // that contain tile k usually carry the same motion vector and then ask for the SAME reference tile: the lane
  // walks the block types, computes a tile's distortion only when its (plane, x, y) differs from the ones it has
  // already computed, and adds the value to each partition's candidate accumulator.
  __shared__ SpWarp<NC> W[WPC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  SpWarp<NC> &S = W[w];
  for (int item = blockIdx.x * WPC + w; item < a.nitems; item += gridDim.x * WPC) {
    const int mb = a.mb_first + item / a.refs_per_mb, ref = a.ref_first + item % a.refs_per_mb;
    const int mbx = mb % a.mbw, mby = mb / a.mbw;
    const size_t base = (a.abs_index ? ((size_t)mb * a.nrefs + ref) : (size_t)item) * NPART;
    const uint8_t *planes = a.planes + (size_t)ref * 16 * a.plane_size;
    __syncwarp();
    for (int t = lane; t < 64; t += 32) reinterpret_cast<uint32_t *>(S.cur)[t] =
        *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + (t >> 2)) * a.cur_pitch + mbx * 16 + (t & 3) * 4);
    for (int p = lane; p < NPART; p += 32) {
      S.mv[p][0] = a.mv_int[(base + p) * 2]; S.mv[p][1] = a.mv_int[(base + p) * 2 + 1];
      S.prd[p][0] = a.pred[(base + p) * 2];  S.prd[p][1] = a.pred[(base + p) * 2 + 1];
      // BlockMotionSearch resets the bound when the metric changes between levels (mv_search.c:971-974)
      S.mincost[p] = a.use_bound ? a.cost_int[base + p] : DMAX;
    }
    const bool allp = a.part_mask == ((1ull << NPART) - 1);
    const int nstage = a.full81 ? 1 : 2;
    for (int stage = 0; stage < nstage; stage++) {
      // full81: full_sub_pel_motion_estimation (me_fullsearch.c:409-469): ONE stage, the 81 quarter-pel positions of
      // spiral_search (radius 4), computePredQPel, lambda[Q_PEL]
      const int step = (stage || a.full81) ? 1 : 2;
      const int metric = (stage || a.full81) ? a.metric_q : a.metric_h;
      const int first = a.full81 ? 0 : (stage ? a.start_qp : a.start_hp);
      const int lam = (stage || a.full81) ? a.lambda_q : a.lambda_h;
      const int ncand = a.full81 ? 81 : 9;
      // ---- uniform-vector path (SATD, nine candidates): when every active partition carries the SAME vector - static
      //      and panning content, the common case - the tile a partition asks for depends on (tile, candidate) only:
      //      144 tile SATDs, then the 41 partition sums of two candidates at a time through the 4x4 / 8x4 / 4x8 / 8x8 /
      //      16x8 / 8x16 / 16x16 tree in shuffles (no shared atomics, no per-block-type bookkeeping) ----
      bool fast = false;
      __syncwarp();
      if (NC == 9 && metric == 2 && a.part_mask != 0) {
        const int pf = __ffsll((long long)a.part_mask) - 1;
        const uint32_t m0 = *reinterpret_cast<const uint32_t *>(&S.mv[pf][0]);
        bool same = true;
        for (int p = lane; p < NPART; p += 32)
          if ((a.part_mask >> p) & 1ull) same = same && *reinterpret_cast<const uint32_t *>(&S.mv[p][0]) == m0;
        fast = __all_sync(0xffffffffu, same);
        if (fast) {
          // lane = (tile k, candidate parity): the lane that computes the SATD of tile k at candidate c is the lane that
          // needs it in the tree, so the value never leaves its register; the tile's current rows are loaded once
          const int mvx = (short)(m0 & 0xffffu), mvy = (short)(m0 >> 16);
          const int k = lane & 15, tx = k & 3, ty = k >> 2;
          uint32_t cw[4];
#pragma unroll
          for (int r = 0; r < 4; r++) cw[r] = *reinterpret_cast<const uint32_t *>(&S.cur[(4 * ty + r) * 16 + 4 * tx]);
          const int bx = 4 * (mbx * 16 + 4 * tx) + mvx, by = 4 * (mby * 16 + 4 * ty) + mvy;
#pragma unroll 1
          for (int pass = 0; pass < 5; pass++) {
            const int c = 2 * pass + (lane >> 4);                  // c == 9: no candidate (zero)
            int v = 0;
            if (c < 9 && c >= first) {
              const int spx = (int)((0x22215u >> (2 * c)) & 3u) - 1, spy = (int)((0x29421u >> (2 * c)) & 3u) - 1;
              const int qx = bx + step * spx, qy = by + step * spy;
              const int pl = (qy & 3) * 4 + (qx & 3);
              const int ox = iclamp(qx >> 2, -PADX, a.W + 15) + PADX, oy = iclamp(qy >> 2, -PADY, a.H + 3) + PADY;
              v = tile_satd(cw, planes + (size_t)pl * a.plane_size + (size_t)oy * a.Wp + ox, a.Wp);
            }
            const int h84 = v + __shfl_xor_sync(0xffffffffu, v, 1);
            const int v48 = v + __shfl_xor_sync(0xffffffffu, v, 4);
            const int e88 = h84 + __shfl_xor_sync(0xffffffffu, h84, 4);
            const int s168 = e88 + __shfl_xor_sync(0xffffffffu, e88, 2);
            const int s816 = e88 + __shfl_xor_sync(0xffffffffu, e88, 8);
            const int s1616 = s168 + __shfl_xor_sync(0xffffffffu, s168, 8);
            if (c < 9) {
              S.dist[25 + k][c] = v;
              if (!(tx & 1)) S.dist[9 + ty * 2 + (tx >> 1)][c] = h84;
              if (!(ty & 1)) S.dist[17 + (ty >> 1) * 4 + tx][c] = v48;
              if (!(tx & 1) && !(ty & 1)) S.dist[5 + (ty >> 1) * 2 + (tx >> 1)][c] = e88;
              if (tx == 0 && !(ty & 1)) S.dist[1 + (ty >> 1)][c] = s168;
              if (ty == 0 && !(tx & 1)) S.dist[3 + (tx >> 1)][c] = s816;
              if (k == 0) S.dist[0][c] = s1616;
            }
          }
        }
      }
      if (!fast) {
      for (int i = lane; i < NPART * NC; i += 32) (&S.dist[0][0])[i] = 0;
      __syncwarp();
      }
      for (int job = lane; job < (fast ? 0 : 16 * ncand); job += 32) {
        const int k = NC == 9 ? (job * 57) >> 9 : job / ncand, c = job - k * ncand;      // job / 9 for job < 144
        if (c < first) continue;
        const int tx = (k & 3) * 4, ty = (k >> 2) * 4;
        uint32_t keys[7]; int vals[7];
        int spx, spy; sp_xy(c, &spx, &spy);
        const int sx = step * spx, sy = step * spy;
#pragma unroll
        for (int bti = 0; bti < 7; bti++) {
          keys[bti] = 0xffffffffu; vals[bti] = 0;
          const int p = part_of_tile(bti, tx, ty);
          if (!allp && !((a.part_mask >> p) & 1ull)) continue;
          const uint32_t mvw = *reinterpret_cast<const uint32_t *>(&S.mv[p][0]);
          int v = -1;
          uint32_t kk;
          if (metric == 2) {          // SATD: per-tile origin clamp (me_distortion.c:771): the tile depends on the vector only
            kk = mvw;
#pragma unroll
            for (int b2 = 0; b2 < bti; b2++) if (keys[b2] == kk) v = vals[b2];
            if (v < 0) {
              const int qx = 4 * (mbx * 16 + tx) + S.mv[p][0] + sx, qy = 4 * (mby * 16 + ty) + S.mv[p][1] + sy;
              const int pl = (qy & 3) * 4 + (qx & 3);
              const int ox = iclamp(qx >> 2, -PADX, a.W + 15) + PADX, oy = iclamp(qy >> 2, -PADY, a.H + 3) + PADY;
              v = tile_distortion(S.cur, tx, ty, planes + (size_t)pl * a.plane_size + (size_t)oy * a.Wp + ox, a.Wp, 2);
            }
          } else {                    // SAD / SSE: block origin clamp (me_distortion.c:367, :1205): key = the tile actually read
            const PartGeom gm = part_geom(p);
            const int qx = 4 * (mbx * 16 + gm.ox) + S.mv[p][0] + sx, qy = 4 * (mby * 16 + gm.oy) + S.mv[p][1] + sy;
            const int pl = (qy & 3) * 4 + (qx & 3);
            const int ox = iclamp(qx >> 2, -PADX, a.W + 15) + PADX + (tx - gm.ox), oy = iclamp(qy >> 2, -PADY, a.H + 3) + PADY + (ty - gm.oy);
            kk = (uint32_t)pl | ((uint32_t)ox << 4) | ((uint32_t)oy << 18);     // ox, oy < 2^14
#pragma unroll
            for (int b2 = 0; b2 < bti; b2++) if (keys[b2] == kk) v = vals[b2];
            if (v < 0) v = tile_distortion(S.cur, tx, ty, planes + (size_t)pl * a.plane_size + (size_t)oy * a.Wp + ox, a.Wp, metric);
          }
          keys[bti] = kk; vals[bti] = v;
          atomicAdd(&S.dist[p][c], v);
        }
      }
      __syncwarp();
      for (int p = lane; p < NPART; p += 32) {
        if (!((a.part_mask >> p) & 1ull)) continue;
        long long best = S.mincost[p]; int best_pos = 0;
        if (NC == 9) {
          // the nine positions share three x and three y displacements: six mvbits instead of eighteen
          const int dx = S.mv[p][0] - S.prd[p][0], dy = S.mv[p][1] - S.prd[p][1];
          const int bxm = mvbits(dx - step), bx0 = mvbits(dx), bxp = mvbits(dx + step);
          const int bym = mvbits(dy - step), by0 = mvbits(dy), byp = mvbits(dy + step);
#pragma unroll
          for (int cc = 0; cc < 9; cc++) {
            if (cc < first) continue;
            const int spx = (int)((0x22215u >> (2 * cc)) & 3u) - 1, spy = (int)((0x29421u >> (2 * cc)) & 3u) - 1;   // compile-time per cc
            const int bits = (spx < 0 ? bxm : (spx > 0 ? bxp : bx0)) + (spy < 0 ? bym : (spy > 0 ? byp : by0));
            const long long cost = (long long)lam * bits + ((long long)S.dist[p][cc] << 5);
            if (cost < best) { best = cost; best_pos = cc; }
          }
        } else {
          for (int cc = first; cc < ncand; cc++) {
            int spx, spy; sp_xy(cc, &spx, &spy);
            const int mvx = S.mv[p][0] + step * spx, mvy = S.mv[p][1] + step * spy;
            const long long cost = (long long)lam * (mvbits(mvx - S.prd[p][0]) + mvbits(mvy - S.prd[p][1])) + ((long long)S.dist[p][cc] << 5);
            if (cost < best) { best = cost; best_pos = cc; }
          }
        }
        { int spx, spy; sp_xy(best_pos, &spx, &spy);
          S.mv[p][0] = (short)(S.mv[p][0] + step * spx); S.mv[p][1] = (short)(S.mv[p][1] + step * spy); }
        if (stage == 0 && !a.full81 && !a.start_qp) best = DMAX;     // me_fullsearch.c:252-253
        S.mincost[p] = best;
      }
      __syncwarp();
    }
    for (int p = lane; p < NPART; p += 32) {
      if (!((a.part_mask >> p) & 1ull)) continue;
      a.mv_sub[(base + p) * 2] = S.mv[p][0]; a.mv_sub[(base + p) * 2 + 1] = S.mv[p][1];
      a.cost_sub[base + p] = S.mincost[p];
    }
  }
}

cudaError_t launch_subpel_refine(const SubArgs &a, cudaStream_t s)
{
  if (a.full81) k_subpel_refine<81, 2><<<(a.nitems + 1) / 2, 64, 0, s>>>(a);
  else k_subpel_refine<9, 4><<<(a.nitems + 3) / 4, 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace b2


// ---- distortion4x4/8x8{SAD,SSE,SATD} on precomputed difference blocks (me_distortion.c:38-134) ----------------
// One thread per block; HBM-bound (32 or 128 bytes in, 8 out).
namespace b2 {

__device__ __forceinline__ int hadamard8x8_abs(const short *d)
{
  // HadamardSAD8x8 (me_distortion.c:266-341): separable 8-point Hadamard, (sum |coef| + 2) >> 2
  int m[64], s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = d[8 * i + j];
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int j = 0; j < 8; j++) if (!(j & len)) { const int x = v[j], y = v[j + len]; v[j] = x + y; v[j + len] = x - y; }
#pragma unroll
    for (int j = 0; j < 8; j++) m[8 * i + j] = v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = m[8 * i + j];
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int i = 0; i < 8; i++) if (!(i & len)) { const int x = v[i], y = v[i + len]; v[i] = x + y; v[i + len] = x - y; }
#pragma unroll
    for (int i = 0; i < 8; i++) s += abs(v[i]);
  }
  return (s + 2) >> 2;
}

__global__ void __launch_bounds__(128) k_distortion(int kind, int n, int nblk, const int16_t *__restrict__ diff, long long *__restrict__ out)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblk) return;
  const int m = n * n;
  short d[64];
  const uint4 *src = reinterpret_cast<const uint4 *>(diff + (size_t)b * m);
  for (int i = 0; i < m / 8; i++) reinterpret_cast<uint4 *>(d)[i] = src[i];
  long long v = 0;
  if (kind == 2) {
    if (n == 4) { int t[16]; for (int i = 0; i < 16; i++) t[i] = d[i]; v = hadamard4x4_abs(t); }
    else v = hadamard8x8_abs(d);
  } else {
    for (int i = 0; i < m; i++) v += kind == 0 ? (long long)abs((int)d[i]) : (long long)d[i] * d[i];
  }
  out[b] = v << 5;
}

cudaError_t launch_distortion(int kind, int n, int nblk, const int16_t *diff, long long *out, cudaStream_t s)
{
  k_distortion<<<(nblk + 127) / 128, 128, 0, s>>>(kind, n, nblk, diff, out);
  return cudaGetLastError();
}

}  // namespace b2


// ---- motion-compensated luma prediction of list 0 (luma_prediction / OneComponentLumaPrediction,
//      JM/lencod/src/mc_prediction.c:117-236), SURVEY 8(f)-1: keeps the vectors on the device between the search and
//      the transform.  One thread per 4x4 block: the block's partition (from the MB mode / 8x8 sub-modes), its
//      vector, UMVLine4X's whole-partition origin clamp, then 4 rows of 4 samples of quarter-pel plane [y&3][x&3];
//      writes prediction and original in the block-packed layout b2tq_4x4 takes.  HBM/L2-bound.
namespace b2 {

__global__ void __launch_bounds__(256) k_mc_luma(const McArgs a)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nmb * 16) return;
  const int m = t >> 4, k = t & 15, bx = k & 3, by = k >> 2, qd = (by >> 1) * 2 + (bx >> 1);
  const int mode = a.mb_mode[m], r = a.ref8[m * 4 + qd];
  int p, ox, oy;
  if (mode == 1) { p = 0; ox = 0; oy = 0; }
  else if (mode == 2) { p = 1 + (by >> 1); ox = 0; oy = 8 * (by >> 1); }
  else if (mode == 3) { p = 3 + (bx >> 1); ox = 8 * (bx >> 1); oy = 0; }
  else {
    const int sub = a.b8mode[m * 4 + qd];
    if (sub == 4) { p = 5 + qd; ox = 8 * (bx >> 1); oy = 8 * (by >> 1); }
    else if (sub == 5) { p = 9 + by * 2 + (bx >> 1); ox = 8 * (bx >> 1); oy = 4 * by; }
    else if (sub == 6) { p = 17 + (by >> 1) * 4 + bx; ox = 4 * bx; oy = 8 * (by >> 1); }
    else { p = 25 + by * 4 + bx; ox = 4 * bx; oy = 4 * by; }
  }
  const int mbx = m % a.mbw, mby = m / a.mbw;
  const int16_t *v = a.mv + (((size_t)m * a.nrefs + r) * NPART + p) * 2;
  const int qx = 4 * (mbx * 16 + ox) + v[0], qy = 4 * (mby * 16 + oy) + v[1];
  const int pl = (qy & 3) * 4 + (qx & 3);
  const int x = iclamp(qx >> 2, -PADX, a.W + 15) + PADX + (4 * bx - ox), y = iclamp(qy >> 2, -PADY, a.H + 3) + PADY + (4 * by - oy);
  const uint8_t *rp = a.planes + ((size_t)r * 16 + pl) * a.plane_size + (size_t)y * a.Wp + x;
  const int al = (int)(reinterpret_cast<size_t>(rp) & 3);
  uint32_t pw[4], cw[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rp + (size_t)i * a.Wp - al);
    const uint32_t lo = w[0], hi = al ? w[1] : 0u;
    pw[i] = al ? __funnelshift_r(lo, hi, 8 * al) : lo;
    cw[i] = *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + 4 * by + i) * a.cur_pitch + mbx * 16 + 4 * bx);
  }
  reinterpret_cast<uint4 *>(a.pred_blk)[t] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
  reinterpret_cast<uint4 *>(a.orig_blk)[t] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
}

// ---- luma + chroma prediction of 4:2:0 macroblocks from either list or both: luma_prediction (JM/lencod/src/mc_prediction.c:
//      144-236: p_dir 0 / 1 / 2, bi_prediction :82-99 = (l0 + l1 + 1) >> 1), chroma_prediction (:469-566) with the bilinear
//      eighth-sample interpolation of OneComponentChromaPrediction4x4_regenerate (:292-353: every chroma sample takes the vector
//      of the luma 4x4 block above it, coordinates clipped to the chroma plane per tap).  Unweighted.  24 threads per macroblock:
//      16 luma 4x4 blocks, 2 x 4 chroma 4x4 blocks; outputs in the block-packed layout of the transform entries. ----
__device__ __forceinline__ void mc_partition(int mode, const uint8_t *b8, int bx, int by, int &p, int &ox, int &oy)
{
  const int qd = (by >> 1) * 2 + (bx >> 1);
  if (mode == 1) { p = 0; ox = 0; oy = 0; }
  else if (mode == 2) { p = 1 + (by >> 1); ox = 0; oy = 8 * (by >> 1); }
  else if (mode == 3) { p = 3 + (bx >> 1); ox = 8 * (bx >> 1); oy = 0; }
  else {
    const int sub = b8[qd];
    if (sub == 4) { p = 5 + qd; ox = 8 * (bx >> 1); oy = 8 * (by >> 1); }
    else if (sub == 5) { p = 9 + by * 2 + (bx >> 1); ox = 8 * (bx >> 1); oy = 4 * by; }
    else if (sub == 6) { p = 17 + (by >> 1) * 4 + bx; ox = 4 * bx; oy = 8 * (by >> 1); }
    else { p = 25 + by * 4 + bx; ox = 4 * bx; oy = 4 * by; }
  }
}

__global__ void __launch_bounds__(192) k_mc_mb(const McMbArgs a)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nmb * 24) return;
  const int m = t / 24, k = t - 24 * m;
  const int mbx = m % a.mbw, mby = m / a.mbw, mode = a.mb_mode[m];
  const int Wc = a.W >> 1, Hc = a.H >> 1;
  if (k < 16) {
    const int bx = k & 3, by = k >> 2, qd = (by >> 1) * 2 + (bx >> 1), dir = a.pdir[m * 4 + qd];
    int p, ox, oy;
    mc_partition(mode, a.b8mode + m * 4, bx, by, p, ox, oy);
    int acc[16];
    for (int i = 0; i < 16; i++) acc[i] = 0;
    for (int L = 0; L < 2; L++) {
      if (!(dir == 2 || dir == L)) continue;
      const int r = a.ref8[(m * 2 + L) * 4 + qd];
      const int16_t *v = (L ? a.mv1 : a.mv0) + (((size_t)m * a.nrefs + r) * NPART + p) * 2;
      const int qx = 4 * (mbx * 16 + ox) + v[0], qy = 4 * (mby * 16 + oy) + v[1];
      const int x = iclamp(qx >> 2, -PADX, a.W + 15) + PADX + (4 * bx - ox), y = iclamp(qy >> 2, -PADY, a.H + 3) + PADY + (4 * by - oy);
      const uint8_t *rp = a.planes + ((size_t)r * 16 + (qy & 3) * 4 + (qx & 3)) * a.plane_size + (size_t)y * a.Wp + x;
      for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i * 4 + j] += rp[(size_t)i * a.Wp + j];
    }
    uint8_t *po = a.pred_y + ((size_t)m * 16 + k) * 16, *oo = a.orig_y + ((size_t)m * 16 + k) * 16;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        po[i * 4 + j] = (uint8_t)(dir == 2 ? (acc[i * 4 + j] + 1) >> 1 : acc[i * 4 + j]);
        oo[i * 4 + j] = a.cur[(size_t)(mby * 16 + 4 * by + i) * a.cur_pitch + mbx * 16 + 4 * bx + j];
      }
  } else {
    const int c = k - 16, pl = c >> 2, cb = c & 3, cbx = cb & 1, cby = cb >> 1, qd = cb, dir = a.pdir[m * 4 + qd];
    uint8_t *po = a.pred_c + (((size_t)m * 2 + pl) * 4 + cb) * 16, *oo = a.orig_c + (((size_t)m * 2 + pl) * 4 + cb) * 16;
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 4; i++) {
        const int ci = 4 * cbx + i, cj = 4 * cby + j;             // chroma sample inside the macroblock's 8x8 block
        int p, ox, oy, acc = 0;
        mc_partition(mode, a.b8mode + m * 4, ci >> 1, cj >> 1, p, ox, oy);
        for (int L = 0; L < 2; L++) {
          if (!(dir == 2 || dir == L)) continue;
          const int r = a.ref8[(m * 2 + L) * 4 + qd];
          const int16_t *v = (L ? a.mv1 : a.mv0) + (((size_t)m * a.nrefs + r) * NPART + p) * 2;
          const uint8_t *rc = a.refc + ((size_t)r * 2 + pl) * Wc * Hc;
          const int ii = 8 * (mbx * 8 + ci) + v[0], jj = 8 * (mby * 8 + cj) + v[1];
          const int x0 = iclamp(ii >> 3, 0, Wc - 1), x1 = iclamp((ii + 7) >> 3, 0, Wc - 1), y0 = iclamp(jj >> 3, 0, Hc - 1), y1 = iclamp((jj + 7) >> 3, 0, Hc - 1);
          const int fx = ii & 7, fy = jj & 7;
          acc += ((8 - fx) * (8 - fy) * rc[(size_t)y0 * Wc + x0] + fx * (8 - fy) * rc[(size_t)y0 * Wc + x1] +
                  (8 - fx) * fy * rc[(size_t)y1 * Wc + x0] + fx * fy * rc[(size_t)y1 * Wc + x1] + 32) >> 6;
        }
        po[j * 4 + i] = (uint8_t)(dir == 2 ? (acc + 1) >> 1 : acc);
        oo[j * 4 + i] = a.curc[(size_t)pl * Wc * Hc + (size_t)(mby * 8 + cj) * Wc + mbx * 8 + ci];
      }
  }
}

cudaError_t launch_mc_mb(const McMbArgs &a, cudaStream_t s)
{
  k_mc_mb<<<(a.nmb * 24 + 191) / 192, 192, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_mc_luma(const McArgs &a, cudaStream_t s)
{
  k_mc_luma<<<(a.nmb * 16 + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace b2
