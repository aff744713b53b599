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
constexpr int FS_QCH = 32;       // queue chunk: candidates re-evaluated per cooperative pass
constexpr int FS_PRE = 1;        // radius of the exact pre-pass around the centres (initial bounds)
constexpr int FS_SMAX = 8;       // partitions whose centres lie within an 8-pel box share one window pass

struct FsSmemLayout {
  int wr, wpitch, copy_stride, win_bytes, total;
  int off_cur, off_bx, off_by, off_misc;
};
__host__ __device__ inline FsSmemLayout fs_layout(int R)
{
  FsSmemLayout L;
  L.wr = 2 * R + 16 + FS_K - 1 + FS_SMAX;             // rows (extra rows for the partial last group)
  L.wpitch = ((2 * R + 16 + 3 + FS_SMAX) + 31) & ~31; // bytes per row
  L.copy_stride = ((L.wr * L.wpitch + 127) & ~127) + 32;
  L.win_bytes = 4 * L.copy_stride;
  L.off_cur = L.win_bytes;
  L.off_bx = L.off_cur + 256;
  L.off_by = L.off_bx + ((2 * R + 1 + FS_SMAX + FS_K + 15) & ~15);
  L.off_misc = L.off_by + ((2 * R + 1 + FS_SMAX + FS_K + 15) & ~15);
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
  signed char pex[NPART], pey[NPART];   // centre of the partition relative to its group's box origin (pel)
  short gx0[NPART], gy0[NPART], gx1[NPART], gy1[NPART];   // centre bounding box of group g (pel)
  unsigned char dupx[NPART], dupy[NPART];   // predictor component already present at a lower partition of the group
  int red[2][4];                  // per-warp centre bounding box (min x, max x, min y, max y)
  int ngroups;
  int err;
  int nhits;
  int qn[2];                      // survivors queued for exact re-evaluation (double-buffered by round parity)
  unsigned short queue[FS_NT * FS_K];   // dx | dy << 8 (window coordinates)
  unsigned short s4[FS_QCH][16];  // 4x4 SADs of the candidates of the current queue chunk
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

// Cooperative exact re-evaluation of the queued window candidates (survivors of the packed
// filter, and the pre-pass neighbourhood of the centres) for every partition of group g:
// phase A computes the sixteen 4x4 SADs of each candidate (16 threads per candidate), phase B
// forms each partition's sum, its exact mv cost in the partition's own spiral and lowers
// best[p] with a 64-bit atomicMin (41 threads per candidate).  Called by the whole CTA.
__device__ __forceinline__ void fs_process_queue(FsShared &S, const uint8_t *smem, const FsSmemLayout L, int R, int g,
                                                 int lambda_f, int par)
{
  const int tid = threadIdx.x;
  const int n = S.qn[par];
  const uint32_t *cur = reinterpret_cast<const uint32_t *>(smem + L.off_cur);
  for (int c0 = 0; c0 < n; c0 += FS_QCH) {
    const int m = min(FS_QCH, n - c0);
    for (int idx = tid; idx < m * 16; idx += FS_NT) {
      const int c = idx >> 4, k = idx & 15, bx = k & 3, by = k >> 2;
      const unsigned q = S.queue[c0 + c];
      const int col = (int)(q & 255u) + 4 * bx, row = (int)(q >> 8) + 4 * by;
      const uint8_t *wb = smem + (col & 3) * L.copy_stride + (col >> 2) * 4 + row * L.wpitch;
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < 4; i++)
        acc = sad4(cur[(by * 4 + i) * 4 + bx], *reinterpret_cast<const uint32_t *>(wb + i * L.wpitch), acc);
      S.s4[c][k] = (unsigned short)acc;
    }
    __syncthreads();
    for (int idx = tid; idx < m * NPART; idx += FS_NT) {
      const int c = idx / NPART, p = idx - c * NPART;
      if (S.pgrp[p] != g) continue;
      const unsigned q = S.queue[c0 + c];
      const int dx = (int)(q & 255u), dy = (int)(q >> 8);
      const int ox = dx - R - S.pex[p], oy = dy - R - S.pey[p];   // displacement from the partition's own centre
      if (max(abs(ox), abs(oy)) > S.psr[p]) continue;
      const PartGeom gm = part_geom(p);
      uint32_t sum = 0;
      for (int by = gm.oy >> 2; by < ((gm.oy + gm.h) >> 2); by++)
        for (int bx = gm.ox >> 2; bx < ((gm.ox + gm.w) >> 2); bx++) sum += S.s4[c][by * 4 + bx];
      const int mvx = S.pcx[p] + 4 * ox, mvy = S.pcy[p] + 4 * oy;
      const long long cost = ((long long)sum << 5) + (long long)lambda_f * (mvbits(mvx - S.ppx[p]) + mvbits(mvy - S.ppy[p]));
      const unsigned long long key = ((unsigned long long)cost << 20) | (unsigned)spiral_index(ox, oy);
      if (key < *reinterpret_cast<volatile unsigned long long *>(&S.best[p])) {
        const unsigned long long old = atomicMin(&S.best[p], key);
        set_threshold(S, p, old < key ? old : key);
      }
    }
    __syncthreads();
  }
  if (tid == 0) { S.nhits += n; S.qn[par] = 0; }
  __syncthreads();
}

// One task in "fine" mode: K candidates (dx, dy0..dy0+K-1), all 41 partitions filtered.
template <int K>
__device__ __forceinline__ void fs_task(FsShared &S, const uint8_t *smem, const FsSmemLayout L, int R, int g,
                                        int dx, int dy0, int ncy, int lambda_f, int par)
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
            if (!allfail && dy0 + j < ncy) S.queue[atomicAdd(&S.qn[par], 1)] = (unsigned short)(dx | ((dy0 + j) << 8));
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
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = a.R;
  const int wpw = L.wpitch >> 2;                      // words per window row
  if (tid == 0) { S.err = 0; S.nhits = 0; S.qn[0] = S.qn[1] = 0; }

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
    // ---- cluster partitions whose centres fit in one FS_SMAX box (usually all of them) ----
    if (tid < 64) {
      const bool act = tid < NPART && S.pgrp[tid] >= 0;
      const int cx = act ? (S.pcx[tid] >> 2) : 0, cy = act ? (S.pcy[tid] >> 2) : 0;
      const int x0 = __reduce_min_sync(0xffffffffu, act ? cx : 0x7fff), x1 = __reduce_max_sync(0xffffffffu, act ? cx : -0x7fff);
      const int y0 = __reduce_min_sync(0xffffffffu, act ? cy : 0x7fff), y1 = __reduce_max_sync(0xffffffffu, act ? cy : -0x7fff);
      if (lane == 0) { S.red[warp][0] = x0; S.red[warp][1] = x1; S.red[warp][2] = y0; S.red[warp][3] = y1; }
    }
    __syncthreads();
    {
      const int bx0 = min(S.red[0][0], S.red[1][0]), bx1 = max(S.red[0][1], S.red[1][1]);
      const int by0 = min(S.red[0][2], S.red[1][2]), by1 = max(S.red[0][3], S.red[1][3]);
      if (bx1 - bx0 <= FS_SMAX && by1 - by0 <= FS_SMAX) {
        if (tid == 0) { S.gx0[0] = bx0; S.gx1[0] = bx1; S.gy0[0] = by0; S.gy1[0] = by1; S.ngroups = bx1 >= bx0 ? 1 : 0; }
      } else if (tid == 0) {        // general case: greedy clustering (serial, rare)
        int ng = 0;
        for (int p = 0; p < NPART; p++) {
          if (S.pgrp[p] < 0) continue;
          const int cx = S.pcx[p] >> 2, cy = S.pcy[p] >> 2;
          int g = -1;
          for (int q = 0; q < ng; q++) {
            const int nx0 = min((int)S.gx0[q], cx), nx1 = max((int)S.gx1[q], cx), ny0 = min((int)S.gy0[q], cy), ny1 = max((int)S.gy1[q], cy);
            if (nx1 - nx0 <= FS_SMAX && ny1 - ny0 <= FS_SMAX) { g = q; S.gx0[q] = nx0; S.gx1[q] = nx1; S.gy0[q] = ny0; S.gy1[q] = ny1; break; }
          }
          if (g < 0) { g = ng++; S.gx0[g] = S.gx1[g] = cx; S.gy0[g] = S.gy1[g] = cy; }
          S.pgrp[p] = (signed char)g;
        }
        S.ngroups = ng;
      }
    }
    __syncthreads();
    const int ngroups = S.ngroups;
    const uint8_t *plane = a.planes + (size_t)ref * 16 * a.plane_size;     // integer plane [0][0]

    for (int g = 0; g < ngroups; g++) {
      if (g) __syncthreads();
      const int spanx = S.gx1[g] - S.gx0[g], spany = S.gy1[g] - S.gy0[g];
      const int ncx = 2 * R + 1 + spanx, ncy = 2 * R + 1 + spany;
      // ---- stage the window (4 byte-shifted copies), padded-plane coordinates ----
      const int x0 = mbx * 16 + S.gx0[g] - R + PADX, y0 = mby * 16 + S.gy0[g] - R + PADY;
      const int nrows = 2 * R + 16 + FS_K - 1 + spany, ncols = 2 * R + 16 + 3 + spanx;
      uint32_t *c0 = reinterpret_cast<uint32_t *>(smem), *c1 = reinterpret_cast<uint32_t *>(smem + L.copy_stride),
               *c2 = reinterpret_cast<uint32_t *>(smem + 2 * L.copy_stride), *c3 = reinterpret_cast<uint32_t *>(smem + 3 * L.copy_stride);
      if (x0 >= 0 && y0 >= 0 && x0 + ncols <= a.Wp && y0 + nrows <= a.Hp) {
        const int al = x0 & 3;
#pragma unroll 4
        for (int r = warp; r < nrows; r += FS_NT / 32) {
          const uint32_t *grow = reinterpret_cast<const uint32_t *>(plane + (size_t)(y0 + r) * a.Wp + (x0 - al));
          for (int j = lane; j < wpw; j += 32) {
            const uint32_t g0 = grow[j], g1 = grow[j + 1], g2 = grow[j + 2];
            // bytes al.. of (g0,g1,g2): copy c starts at byte al + c
            const uint32_t lo = al ? __funnelshift_r(g0, g1, 8 * al) : g0;      // window word j
            const uint32_t hi = al ? __funnelshift_r(g1, g2, 8 * al) : g1;      // window word j+1
            c0[r * wpw + j] = lo;
            c1[r * wpw + j] = __funnelshift_r(lo, hi, 8);
            c2[r * wpw + j] = __funnelshift_r(lo, hi, 16);
            c3[r * wpw + j] = __funnelshift_r(lo, hi, 24);
          }
        }
      } else {                      // window leaves the padded plane: per-pixel coordinate clamp
        for (int r = warp; r < nrows; r += FS_NT / 32) {
          const uint8_t *grow = plane + (size_t)iclamp(y0 + r, 0, a.Hp - 1) * a.Wp;
          for (int j = lane; j < wpw; j += 32) {
            uint32_t b[7];
#pragma unroll
            for (int k = 0; k < 7; k++) b[k] = grow[iclamp(x0 + 4 * j + k, 0, a.Wp - 1)];
            c0[r * wpw + j] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
            c1[r * wpw + j] = b[1] | (b[2] << 8) | (b[3] << 16) | (b[4] << 24);
            c2[r * wpw + j] = b[2] | (b[3] << 8) | (b[4] << 16) | (b[5] << 24);
            c3[r * wpw + j] = b[3] | (b[4] << 8) | (b[5] << 16) | (b[6] << 24);
          }
        }
      }
      // thresholds: partitions of other groups never pass
      if (tid < 18) S.Cw[tid] = 0x80008000u;
      if (tid < 5) S.Bs[tid] = -1;
      if (tid < NPART && S.pgrp[tid] == g) { S.pex[tid] = (signed char)((S.pcx[tid] >> 2) - S.gx0[g]); S.pey[tid] = (signed char)((S.pcy[tid] >> 2) - S.gy0[g]); }
      if (tid >= 64 && tid < 64 + NPART && S.pgrp[tid - 64] == g) {     // predictor component seen at a lower partition?
        const int p = tid - 64;
        bool fx = false, fy = false;
        for (int q = 0; q < p; q++)
          if (S.pgrp[q] == g) { fx |= (S.ppx[q] == S.ppx[p]); fy |= (S.ppy[q] == S.ppy[p]); }
        S.dupx[p] = fx; S.dupy[p] = fy;
      }
      __syncthreads();
      // lower bound of the mv bits over the predictors of the group, per window column / row:
      // every partition of the group sees the same absolute displacement 4*(g0 + d - R)
      for (int i = tid; i < ncx + ncy + FS_K; i += FS_NT) {
        const bool isy = i >= ncx;
        const int d = isy ? i - ncx : i;
        const int mv = 4 * ((isy ? S.gy0[g] : S.gx0[g]) + d - R);
        const unsigned char *dup = isy ? S.dupy : S.dupx;
        const short *pp = isy ? S.ppy : S.ppx;
        int mn = 255;
        for (int q = 0; q < NPART; q++)
          if (S.pgrp[q] == g && !dup[q]) mn = min(mn, mvbits(mv - pp[q]));
        (isy ? smem + L.off_by : smem + L.off_bx)[d] = (uint8_t)mn;
      }
      // ---- initial bounds: exact pre-pass over the neighbourhood of the group's centres ----
      {
        const int xlo = max(0, R - FS_PRE), ylo = max(0, R - FS_PRE);
        const int pw = min(ncx - 1, R + spanx + FS_PRE) - xlo + 1, ph = min(ncy - 1, R + spany + FS_PRE) - ylo + 1;
        for (int i = tid; i < pw * ph; i += FS_NT) {
          const int dx = xlo + i % pw, dy = ylo + i / pw;
          S.queue[i] = (unsigned short)(dx | (dy << 8));
        }
        if (tid == 0) S.qn[0] = pw * ph;
      }
      __syncthreads();
      fs_process_queue(S, smem, L, R, g, a.lambda_f, 0);
      // ---- main pass: rounds of FS_NT tasks, survivors re-evaluated after each round ----
      const int ngrp = (ncy + FS_K - 1) / FS_K, ntask = ncx * ngrp;
      for (int t0 = 0, par = 0; t0 < ntask; t0 += FS_NT, par ^= 1) {
        const int t = t0 + tid;
        if (t < ntask) {
          const int gy = t / ncx, dx = t - gy * ncx;
          fs_task<FS_K>(S, smem, L, R, g, dx, gy * FS_K, ncy, a.lambda_f, par);
        }
        __syncthreads();
        if (S.qn[par]) fs_process_queue(S, smem, L, R, g, a.lambda_f, par);
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
    if (tid == 0 && a.stats) { atomicAdd(&a.stats[0], (unsigned long long)S.nhits); atomicAdd(&a.stats[1], (unsigned long long)ngroups); atomicAdd(&a.stats[2], 1ull); S.nhits = 0; }
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
