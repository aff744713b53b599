// subpel_refine.cu -- half- and quarter-pel refinement of all 41 partitions of a macroblock.
//
// Replaces, per (macroblock, reference): 41 calls of sub_pel_motion_estimation
// (JM/lencod/src/me_fullsearch.c:186-289) with computeSATD (me_distortion.c:745-825,
// HadamardSAD4x4 :175-258) or computeSAD (:349-426) as computePredHPel/QPel.
//
// One CTA per (MB, ref).  A unit of work is one 4x4 tile of one candidate of one partition:
// 9 half-pel (then 8 quarter-pel) candidates x 7 block types x 16 tiles; each thread loads the
// 4x4 reference tile from the quarter-pel plane [y&3][x&3] with the reference's own
// tile-origin clamp (UMVLine4X, refbuf.h:22-26), forms the difference against the current MB
// in shared memory, runs the 4x4 Hadamard in registers and adds (satd+1)>>1 into the
// candidate's accumulator.  41 threads then take the lexicographic (cost, position) minimum in
// spiral order with the reference's carried / reset min_mcost rules (mv_search.c:971-974,
// me_fullsearch.c:252-253).
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

constexpr int SP_NT = 128;
constexpr long long DMAX = ((long long)0x7fffffff) << 5;

// spiral positions 0..8 (x,y) in units of the step (mv_search.c:406-442 with l = 1)
__constant__ signed char c_sp9[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

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

__global__ void __launch_bounds__(SP_NT) k_subpel_refine(const SubArgs a)
{
  __shared__ __align__(16) uint8_t cur[256];
  __shared__ int dist[NPART][9];
  __shared__ short mv[NPART][2], prd[NPART][2];
  __shared__ long long mincost[NPART];
  const int tid = threadIdx.x;
  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int mb = a.mb_first + item / a.refs_per_mb, ref = a.ref_first + item % a.refs_per_mb;
    const int mbx = mb % a.mbw, mby = mb / a.mbw;
    const size_t base = (a.abs_index ? ((size_t)mb * a.nrefs + ref) : (size_t)item) * NPART;
    const uint8_t *planes = a.planes + (size_t)ref * 16 * a.plane_size;
    __syncthreads();
    if (tid < 64) reinterpret_cast<uint32_t *>(cur)[tid] =
        *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + (tid >> 2)) * a.cur_pitch + mbx * 16 + (tid & 3) * 4);
    if (tid < NPART) {
      mv[tid][0] = a.mv_int[(base + tid) * 2]; mv[tid][1] = a.mv_int[(base + tid) * 2 + 1];
      prd[tid][0] = a.pred[(base + tid) * 2];  prd[tid][1] = a.pred[(base + tid) * 2 + 1];
      // BlockMotionSearch resets the bound when the metric changes between levels (mv_search.c:971-974)
      mincost[tid] = a.use_bound ? a.cost_int[base + tid] : DMAX;
    }
    for (int stage = 0; stage < 2; stage++) {
      const int step = stage ? 1 : 2;
      const int metric = stage ? a.metric_q : a.metric_h;
      const int first = stage ? a.start_qp : a.start_hp;
      const int lam = stage ? a.lambda_q : a.lambda_h;
      __syncthreads();
      for (int i = tid; i < NPART * 9; i += SP_NT) (&dist[0][0])[i] = 0;
      __syncthreads();
      // units: candidate c (0..8) x blocktype index (0..6) x tile k (0..15)
      for (int u = tid; u < 9 * 112; u += SP_NT) {
        const int c = u / 112, v = u - c * 112, bti = v >> 4, k = v & 15;
        if (c < first) continue;
        const int tx = (k & 3) * 4, ty = (k >> 2) * 4;
        int p;   // partition of type bti+1 containing tile k
        switch (bti) {
          case 0: p = 0; break;
          case 1: p = 1 + (ty >> 3); break;
          case 2: p = 3 + (tx >> 3); break;
          case 3: p = 5 + (ty >> 3) * 2 + (tx >> 3); break;
          case 4: p = 9 + (ty >> 2) * 2 + (tx >> 3); break;
          case 5: p = 17 + (ty >> 3) * 4 + (tx >> 2); break;
          default: p = 25 + (ty >> 2) * 4 + (tx >> 2); break;
        }
        if (!((a.part_mask >> p) & 1ull)) continue;
        const int mvx = mv[p][0] + step * c_sp9[c][0], mvy = mv[p][1] + step * c_sp9[c][1];
        int ox, oy, pl;
        if (metric == 2) {          // SATD: per-tile origin clamp (me_distortion.c:771)
          const int qx = 4 * (mbx * 16 + tx) + mvx, qy = 4 * (mby * 16 + ty) + mvy;
          pl = (qy & 3) * 4 + (qx & 3);
          ox = iclamp(qx >> 2, -PADX, a.W + 15) + PADX; oy = iclamp(qy >> 2, -PADY, a.H + 3) + PADY;
        } else {                    // SAD: block origin clamp (me_distortion.c:367)
          const PartGeom gm = part_geom(p);
          const int qx = 4 * (mbx * 16 + gm.ox) + mvx, qy = 4 * (mby * 16 + gm.oy) + mvy;
          pl = (qy & 3) * 4 + (qx & 3);
          ox = iclamp(qx >> 2, -PADX, a.W + 15) + PADX + (tx - gm.ox); oy = iclamp(qy >> 2, -PADY, a.H + 3) + PADY + (ty - gm.oy);
        }
        const uint8_t *rp = planes + (size_t)pl * a.plane_size + (size_t)oy * a.Wp + ox;
        const int al = (int)(reinterpret_cast<size_t>(rp) & 3);      // two aligned words per row instead of four byte loads
        int d[16];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const uint32_t *w = reinterpret_cast<const uint32_t *>(rp + (size_t)r * a.Wp - al);
          const uint32_t lo = w[0], hi = al ? w[1] : 0u;
          const uint32_t px = al ? __funnelshift_r(lo, hi, 8 * al) : lo;
          const uint32_t cw = *reinterpret_cast<const uint32_t *>(&cur[(ty + r) * 16 + tx]);
#pragma unroll
          for (int q = 0; q < 4; q++) d[r * 4 + q] = (int)((cw >> (8 * q)) & 255u) - (int)((px >> (8 * q)) & 255u);
        }
        int val;
        if (metric == 2) val = hadamard4x4_abs(d);
        else { val = 0;
#pragma unroll
          for (int q = 0; q < 16; q++) val += abs(d[q]); }
        atomicAdd(&dist[p][c], val);
      }
      __syncthreads();
      if (tid < NPART && ((a.part_mask >> tid) & 1ull)) {
        const int p = tid;
        long long best = mincost[p]; int best_pos = 0;
        for (int c = first; c < 9; c++) {
          const int mvx = mv[p][0] + step * c_sp9[c][0], mvy = mv[p][1] + step * c_sp9[c][1];
          const long long cost = (long long)lam * (mvbits(mvx - prd[p][0]) + mvbits(mvy - prd[p][1])) + ((long long)dist[p][c] << 5);
          if (cost < best) { best = cost; best_pos = c; }
        }
        mv[p][0] = (short)(mv[p][0] + step * c_sp9[best_pos][0]);
        mv[p][1] = (short)(mv[p][1] + step * c_sp9[best_pos][1]);
        if (stage == 0 && !a.start_qp) best = DMAX;     // me_fullsearch.c:252-253
        mincost[p] = best;
      }
    }
    __syncthreads();
    if (tid < NPART && ((a.part_mask >> tid) & 1ull)) {
      a.mv_sub[(base + tid) * 2] = mv[tid][0]; a.mv_sub[(base + tid) * 2 + 1] = mv[tid][1];
      a.cost_sub[base + tid] = mincost[tid];
    }
  }
}

cudaError_t launch_subpel_refine(const SubArgs &a, cudaStream_t s)
{
  k_subpel_refine<<<a.nitems, SP_NT, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace b2
