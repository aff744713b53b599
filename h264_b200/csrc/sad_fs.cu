// sad_fs.cu -- integer-pel full search for all 41 partitions of a macroblock.
//
// Replaces, per (macroblock, reference):  41 calls of full_search_motion_estimation
// (JM/lencod/src/me_fullsearch.c:39-103), each looping (2R+1)^2 times over computeSAD
// (JM/lencod/src/me_distortion.c:349-426) and mv_cost (JM/lencod/inc/mv_search.h:100-104).
//
// Formulation (the reference's own fast-full-search decomposition, me_fullfast.c:195-260):
// the SAD of every partition at a candidate is a sum of the sixteen 4x4 SADs of the MB at
// that candidate, so each pixel pair is compared ONCE per candidate (VABSDIFF4.U8.ACC on four
// packed bytes) and the 41 partition costs are packed-16-bit tree sums.
//
// Exactness notes
//  * UMVLine4X clamps the block ORIGIN to [-32,W+15]x[-20,H+3] (refbuf.h:25).  In the integer
//    plane the pad is pure edge replication and those bounds lie inside it, so the clamped
//    read equals a per-pixel coordinate clamp; the window is staged with per-pixel clamps and
//    the tree-sum identity holds at every candidate, frame borders included.
//  * argmin = min over (cost, spiral position) lexicographically with the caller's min_mcost
//    as initial bound (strict '<' in scan order, me_fullsearch.c:83-93).  The early exits of
//    the reference are result-neutral (SURVEY Q-J1).
//  * Partitions may have different search centres.  Partitions are grouped by centre; each
//    group is one pass over its own window.
//
// Work decomposition: one CTA per (MB, ref) item.  A task = one column dx of the window and K
// consecutive rows dy0..dy0+K-1 of candidates, processed in lock-step by one thread with a
// sliding window of K reference rows in registers (each reference row is loaded once per task
// and used by K candidates; each current row is one broadcast LDS.128).  After every 4 rows
// the 4x4 SADs of a block row are packed, tree-summed and compared against per-partition
// thresholds held in shared memory with one IADD3 per pair of partitions:
//        fail  <=>  sad_p + m(d) > B_p ,  m(d) = floor(lambda*minbits(d)/32), B_p = best_p>>5
// which is a necessary condition for (cost,pos) < best, so nothing that could win is dropped.
// Survivors (rare) are re-evaluated exactly and update best_p with a 64-bit atomicMin.
//
// Shared-memory window: 4 byte-shifted copies (copy c holds the window shifted left by c
// bytes) so that the 16 pixels of any candidate column are four aligned 32-bit words; copies
// are offset by 8 banks so that 32 consecutive dx hit 32 distinct banks.
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

constexpr int FS_K = 5;          // candidates per task (lock-step rows)
constexpr int FS_NT = 128;       // threads per CTA

struct FsSmemLayout {
  int wr, wpitch, copy_stride, win_bytes, total;
  int off_cur, off_bx, off_by, off_misc;
};
__host__ __device__ inline FsSmemLayout fs_layout(int R)
{
  FsSmemLayout L;
  L.wr = 2 * R + 16 + FS_K - 1;                       // rows (extra rows for the partial last group)
  L.wpitch = ((2 * R + 16 + 3) + 31) & ~31;           // bytes per row
  L.copy_stride = ((L.wr * L.wpitch + 127) & ~127) + 32;
  L.win_bytes = 4 * L.copy_stride;
  L.off_cur = L.win_bytes;
  L.off_bx = L.off_cur + 256;
  L.off_by = L.off_bx + ((2 * R + 1 + FS_K + 15) & ~15);
  L.off_misc = L.off_by + ((2 * R + 1 + FS_K + 15) & ~15);
  L.total = L.off_misc;
  return L;
}

struct FsShared {                 // static shared part
  unsigned long long best[NPART]; // (cost << 20) | pos
  uint32_t Cw[18];                // packed thresholds, C = 0x7fff - B  (two partitions per word)
  int Bs[5];                      // scalar thresholds B: 16x8 top, bottom, 8x16 left, right, 16x16
  short pcx[NPART], pcy[NPART];   // centre (relative MV, quarter-pel)
  short ppx[NPART], ppy[NPART];   // predictor (quarter-pel)
  short psr[NPART];               // per-partition search range (pel)
  signed char pgrp[NPART];        // centre group of the partition (-1 inactive)
  short gcx[NPART], gcy[NPART];   // centre of group g (pel)
  int ngroups;
  int err;
};

// index of partition p in the packed-threshold array (u16 view of Cw) or -1-s for scalar s
__device__ __forceinline__ int cidx(int p)
{
  if (p >= 25) return p - 25;            // 4x4   -> words 0..7
  if (p >= 17) return 24 + (p - 17);     // 4x8   -> words 12..15
  if (p >= 9)  return 16 + (p - 9);      // 8x4   -> words 8..11
  if (p >= 5)  return 32 + (p - 5);      // 8x8   -> words 16,17
  return p == 0 ? -5 : -p;               // scalars: p1->-1 p2->-2 p3->-3 p4->-4 p0->-5
}

__device__ __forceinline__ void set_threshold(FsShared &S, int p, unsigned long long key)
{
  unsigned long long b = (key >> 20) >> 5;
  int c = cidx(p);
  if (c >= 0) {
    uint32_t B = b > 0x7fffull ? 0x7fffu : (uint32_t)b;
    reinterpret_cast<volatile uint16_t *>(S.Cw)[c] = (uint16_t)(0x7fffu - B);
  } else {
    reinterpret_cast<volatile int *>(S.Bs)[-c - 1] = b > 0x3fffffffull ? 0x3fffffff : (int)b;
  }
}

// Exact evaluation of candidate (dx,dy) (window coordinates, 0..2R) for every partition of
// group g: recompute the sixteen 4x4 SADs, form each partition sum, exact mv cost, and
// atomically lower best[p].  Slow path: runs for survivors of the packed filter only.
__device__ __noinline__ void fs_exact_eval(FsShared &S, const uint8_t *smem, const FsSmemLayout L, int R, int g,
                                           int dx, int dy, int lambda_f, unsigned long long mask)
{
  const uint8_t *wb = smem + (dx & 3) * L.copy_stride + (dx >> 2) * 4 + dy * L.wpitch;
  const uint32_t *cur = reinterpret_cast<const uint32_t *>(smem + L.off_cur);
  uint32_t s[16];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t *rw = reinterpret_cast<const uint32_t *>(wb + (b * 4 + i) * L.wpitch);
      const uint32_t *cw = cur + (b * 4 + i) * 4;
      a0 = sad4(cw[0], rw[0], a0); a1 = sad4(cw[1], rw[1], a1);
      a2 = sad4(cw[2], rw[2], a2); a3 = sad4(cw[3], rw[3], a3);
    }
    s[b * 4 + 0] = a0; s[b * 4 + 1] = a1; s[b * 4 + 2] = a2; s[b * 4 + 3] = a3;
  }
  const int ox = dx - R, oy = dy - R;                       // displacement from the group centre (pel)
  const int ring = max(abs(ox), abs(oy));
  const int pos = spiral_index(ox, oy);
#pragma unroll
  for (int p = 0; p < NPART; p++) {
    if (!((mask >> p) & 1ull) || S.pgrp[p] != g) continue;
    const PartGeom gm = part_geom(p);
    uint32_t sum = 0;
#pragma unroll
    for (int by = 0; by < 4; by++)
#pragma unroll
      for (int bx = 0; bx < 4; bx++)
        if (bx * 4 >= gm.ox && bx * 4 < gm.ox + gm.w && by * 4 >= gm.oy && by * 4 < gm.oy + gm.h) sum += s[by * 4 + bx];
    if (ring > S.psr[p]) continue;
    const int mvx = S.pcx[p] + 4 * ox, mvy = S.pcy[p] + 4 * oy;
    const long long cost = ((long long)sum << 5) + (long long)lambda_f * (mvbits(mvx - S.ppx[p]) + mvbits(mvy - S.ppy[p]));
    const unsigned long long key = ((unsigned long long)cost << 20) | (unsigned)pos;
    if (key < *reinterpret_cast<volatile unsigned long long *>(&S.best[p])) {
      unsigned long long old = atomicMin(&S.best[p], key);
      set_threshold(S, p, old < key ? old : key);
    }
  }
}

// One task in "fine" mode: K candidates (dx, dy0..dy0+K-1), all 41 partitions filtered.
template <int K>
__device__ __forceinline__ void fs_task(FsShared &S, const uint8_t *smem, const FsSmemLayout L, int R, int g,
                                        int dx, int dy0, int lambda_f, unsigned long long mask)
{
  const uint8_t *wb = smem + (dx & 3) * L.copy_stride + (dx >> 2) * 4 + dy0 * L.wpitch;
  const uint4 *cur = reinterpret_cast<const uint4 *>(smem + L.off_cur);
  const uint8_t *tbx = smem + L.off_bx, *tby = smem + L.off_by;
  uint32_t rw[K][4];
  uint32_t acc[K][4];
  uint32_t m2[K], keepA01[K], keepA23[K], keepH[K], E0[K];
  uint32_t fail[K];
  const int bx = tbx[dx];
#pragma unroll
  for (int j = 0; j < K; j++) {
    uint32_t m = ((uint32_t)lambda_f * (uint32_t)(bx + tby[dy0 + j])) >> 5;
    m2[j] = m * 0x10001u;
    fail[j] = 0xffffffffu;
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0;
  }
#pragma unroll
  for (int j = 0; j < K - 1; j++) {
    const uint32_t *r = reinterpret_cast<const uint32_t *>(wb + j * L.wpitch);
    rw[j][0] = r[0]; rw[j][1] = r[1]; rw[j][2] = r[2]; rw[j][3] = r[3];
  }
#pragma unroll
  for (int i = 0; i < 16; i++) {
    {
      const uint32_t *r = reinterpret_cast<const uint32_t *>(wb + (i + K - 1) * L.wpitch);
      const int sl = (i + K - 1) % K;
      rw[sl][0] = r[0]; rw[sl][1] = r[1]; rw[sl][2] = r[2]; rw[sl][3] = r[3];
    }
    const uint4 c = cur[i];
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int sl = (i + j) % K;
      acc[j][0] = sad4(c.x, rw[sl][0], acc[j][0]);
      acc[j][1] = sad4(c.y, rw[sl][1], acc[j][1]);
      acc[j][2] = sad4(c.z, rw[sl][2], acc[j][2]);
      acc[j][3] = sad4(c.w, rw[sl][3], acc[j][3]);
    }
    if ((i & 3) == 3) {
      const int b = i >> 2;
      const volatile uint32_t *Cw = S.Cw;
      const uint32_t c0 = Cw[2 * b], c1 = Cw[2 * b + 1], c2 = Cw[8 + b];
      uint32_t c3 = 0, c4 = 0, c5 = 0;
      if (b & 1) { c3 = Cw[12 + (b >> 1) * 2]; c4 = Cw[13 + (b >> 1) * 2]; c5 = Cw[16 + (b >> 1)]; }
#pragma unroll
      for (int j = 0; j < K; j++) {
        const uint32_t A01 = acc[j][0] + (acc[j][1] << 16);
        const uint32_t A23 = acc[j][2] + (acc[j][3] << 16);
        const uint32_t Hh = (acc[j][0] + acc[j][1]) + ((acc[j][2] + acc[j][3]) << 16);
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0;
        fail[j] &= (A01 + m2[j] + c0) & (A23 + m2[j] + c1) & (Hh + m2[j] + c2);
        if (b & 1) {
          const uint32_t V01 = A01 + keepA01[j], V23 = A23 + keepA23[j], E = Hh + keepH[j];
          fail[j] &= (V01 + m2[j] + c3) & (V23 + m2[j] + c4) & (E + m2[j] + c5);
          if (b == 1) E0[j] = E;
          else {
            const int m = (int)(m2[j] & 0xffffu);
            const int e0l = E0[j] & 0xffffu, e0h = E0[j] >> 16, e1l = E & 0xffffu, e1h = E >> 16;
            const int top = e0l + e0h, bot = e1l + e1h, left = e0l + e1l, right = e0h + e1h, all = top + bot;
            const volatile int *Bs = S.Bs;
            const int sf = (Bs[0] - top - m) & (Bs[1] - bot - m) & (Bs[2] - left - m) & (Bs[3] - right - m) & (Bs[4] - all - m);
            const bool allfail = ((fail[j] & 0x80008000u) == 0x80008000u) && (sf < 0);
            if (!allfail && dy0 + j <= 2 * R) fs_exact_eval(S, smem, L, R, g, dx, dy0 + j, lambda_f, mask);
          }
        } else { keepA01[j] = A01; keepA23[j] = A23; keepH[j] = Hh; }
      }
    }
  }
}

__global__ void __launch_bounds__(FS_NT) k_sad_fs(const FsArgs a)
{
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ FsShared S;
  const FsSmemLayout L = fs_layout(a.R);
  const int tid = threadIdx.x;
  const int R = a.R, NC = 2 * R + 1;
  if (tid == 0) S.err = 0;

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int mb = a.mb_first + item / a.refs_per_mb, ref = a.ref_first + item % a.refs_per_mb;
    const int mbx = mb % a.mbw, mby = mb / a.mbw;
    const size_t base = (a.abs_index ? ((size_t)mb * a.nrefs + ref) : (size_t)item) * NPART;
    __syncthreads();   // previous item fully finished with S / smem
    // ---- per-partition parameters, current MB ----
    if (tid < NPART) {
      const int p = tid;
      const bool act = (a.part_mask >> p) & 1ull;
      const PartGeom gm = part_geom(p);
      S.pcx[p] = a.center[(base + p) * 2]; S.pcy[p] = a.center[(base + p) * 2 + 1];
      S.ppx[p] = a.pred[(base + p) * 2];   S.ppy[p] = a.pred[(base + p) * 2 + 1];
      S.psr[p] = (short)(a.restrict_mode < 0 ? a.sr_override : block_search_range(R, a.restrict_mode, ref, gm.bt));
      S.pgrp[p] = act ? 0 : -1;
      S.best[p] = ((unsigned long long)a.min_mcost << 20);
      if (act && ((S.pcx[p] | S.pcy[p]) & 3)) S.err = 1;       // sub-pel centres are not a full-search input
    }
    if (tid < 64) reinterpret_cast<uint32_t *>(smem + L.off_cur)[tid] =
        *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + (tid >> 2)) * a.cur_pitch + mbx * 16 + (tid & 3) * 4);
    __syncthreads();
    if (tid == 0) {                                           // group partitions by centre
      int ng = 0;
      for (int p = 0; p < NPART; p++) {
        if (S.pgrp[p] < 0) continue;
        int g = -1;
        for (int q = 0; q < ng; q++) if (S.gcx[q] == (S.pcx[p] >> 2) && S.gcy[q] == (S.pcy[p] >> 2)) { g = q; break; }
        if (g < 0) { g = ng++; S.gcx[g] = S.pcx[p] >> 2; S.gcy[g] = S.pcy[p] >> 2; }
        S.pgrp[p] = (signed char)g;
      }
      S.ngroups = ng;
    }
    __syncthreads();
    const int ngroups = S.ngroups;
    const uint8_t *plane = a.planes + (size_t)ref * 16 * a.plane_size;     // integer plane [0][0]

    for (int g = 0; g < ngroups; g++) {
      __syncthreads();
      // ---- stage the window: rows y0.., cols x0.. in padded-plane coordinates, clamped ----
      const int x0 = mbx * 16 + S.gcx[g] - R + PADX, y0 = mby * 16 + S.gcy[g] - R + PADY;
      const int wcols = 2 * R + 16 + 3;
      for (int i = tid; i < L.wr * L.wpitch; i += FS_NT) {
        const int r = i / L.wpitch, c = i - r * L.wpitch;
        uint8_t v = 0;
        if (c < wcols) v = plane[(size_t)iclamp(y0 + r, 0, a.Hp - 1) * a.Wp + iclamp(x0 + c, 0, a.Wp - 1)];
        smem[i] = v;
      }
      // thresholds: groups other than g never pass
      if (tid < 18) S.Cw[tid] = 0x80008000u;
      if (tid < 5) S.Bs[tid] = -1;
      // lower bound of mv bits over the partitions of the group, per window column / row
      for (int i = tid; i < 2 * (NC + FS_K); i += FS_NT) {
        const int isy = i >= NC + FS_K, d = isy ? i - (NC + FS_K) : i;
        int mn = 255;
        for (int p = 0; p < NPART; p++) {
          if (S.pgrp[p] != g) continue;
          const int v = isy ? mvbits(S.pcy[p] + 4 * (d - R) - S.ppy[p]) : mvbits(S.pcx[p] + 4 * (d - R) - S.ppx[p]);
          mn = min(mn, v);
        }
        (isy ? smem + L.off_by : smem + L.off_bx)[d] = (uint8_t)mn;
      }
      __syncthreads();
      // shifted copies 1..3 from copy 0
      for (int i = tid; i < 3 * L.wr * (L.wpitch / 4); i += FS_NT) {
        const int c = 1 + i / (L.wr * (L.wpitch / 4)), w = i % (L.wr * (L.wpitch / 4));
        const uint32_t *src = reinterpret_cast<const uint32_t *>(smem) + w;
        const uint32_t lo = src[0], hi = ((w + 1) % (L.wpitch / 4)) ? src[1] : 0u;
        reinterpret_cast<uint32_t *>(smem + c * L.copy_stride)[w] = __funnelshift_r(lo, hi, 8 * c);
      }
      __syncthreads();
      // ---- initial thresholds: exact cost of the centre candidate (spiral pos 0) ----
      if (tid == 0) fs_exact_eval(S, smem, L, R, g, R, R, a.lambda_f, a.part_mask);
      __syncthreads();
      // ---- main pass ----
      const int ngrp = (NC + FS_K - 1) / FS_K, ntask = NC * ngrp;
      for (int t = tid; t < ntask; t += FS_NT) {
        const int gy = t / NC, dx = t - gy * NC;
        fs_task<FS_K>(S, smem, L, R, g, dx, gy * FS_K, a.lambda_f, a.part_mask);
      }
    }
    __syncthreads();
    // ---- results ----
    if (tid < NPART && ((a.part_mask >> tid) & 1ull)) {
      const int p = tid;
      const unsigned long long key = S.best[p];
      const int pos = (int)(key & 0xfffffull);
      int sx, sy; spiral_xy(pos, &sx, &sy);
      a.mv_int[(base + p) * 2]     = (int16_t)(S.pcx[p] + 4 * sx);
      a.mv_int[(base + p) * 2 + 1] = (int16_t)(S.pcy[p] + 4 * sy);
      a.cost_int[base + p] = (long long)(key >> 20);
    }
    if (tid == 0 && S.err) { *a.errflag = 1; S.err = 0; }
  }
}

cudaError_t launch_sad_fs(const FsArgs &a, int sm_count, cudaStream_t s, int *smem_bytes_out)
{
  const FsSmemLayout L = fs_layout(a.R);
  static int configured = 0;
  if (configured < L.total) {
    cudaError_t e = cudaFuncSetAttribute(k_sad_fs, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return e;
    configured = L.total;
  }
  if (smem_bytes_out) *smem_bytes_out = L.total;
  int grid = a.nitems;
  k_sad_fs<<<grid, FS_NT, L.total, s>>>(a);
  return cudaGetLastError();
}

}  // namespace b2
