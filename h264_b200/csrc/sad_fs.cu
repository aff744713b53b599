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
// Work decomposition.  One CTA (4 warps) works on FS_NSLOT (MB, ref) items at a time, each with
// its own window in shared memory.  A warp-task = 64 window columns x K rows of candidates of one
// item: lane l owns the column pair (dx, dx+4), dx = 8*(l>>2) + (l&3), and walks K consecutive
// rows dy0..dy0+K-1 in lock step with a sliding window of K reference rows in registers, so a
// reference row costs three LDS.64 per 2K candidates and a current row one broadcast LDS.128.
// Warps pull tasks from a shared counter and never meet at a barrier inside an item.
//
// Filter.  After every 4 rows the 4x4 SADs of a block row are packed two per register (IMAD on the
// FMA pipe), tree-summed, and folded into one running minimum per candidate with one
// VIADDMNMX.S16x2 per pair of partitions:   run = min(run, sad_p - B_p - 1),  B_p = best_p >> 5.
// At the end  run + m < 0  (m = floor(lambda*minbits/32), a lower bound of the candidate's mv
// cost over the partitions' predictors) is a NECESSARY condition for (cost,pos) < best of some
// partition, so nothing that could win is dropped.  Survivors (rare) are re-evaluated exactly by
// the warp itself (16 lanes per candidate for the 4x4 SADs, a summed-area table, then one lane
// per (candidate, partition)) and lower best_p with a 64-bit atomicMin on (cost << 20 | spiral pos).
//
// Shared-memory window: 4 byte-shifted copies (copy c holds the window shifted left by c bytes) so
// that the 16+4 pixels of a column pair are three aligned 64-bit words; copy c is stored c rows
// lower so that, with a row pitch of 24 (mod 32) words, the four copies start 8 banks apart while
// every copy stays 128-byte aligned (TMA destination rule).
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

constexpr int FS_K = 5;          // candidate rows per task (lock step)
constexpr int FS_NW = 4;         // warps per CTA
constexpr int FS_NT = FS_NW * 32;
constexpr int FS_PRE = 0;        // radius of the exact pre-pass around the centres (initial bounds)
constexpr int FS_SMAX = 8;       // partitions whose centres lie within an 8-pel box share one window pass
constexpr int FS_BMAX = 12;      // a trailing column block of at most this width is walked row-major (type B)
constexpr int FS_MCAP = 2047;    // cap of each half of the mv-cost lower bound (keeps packed sums in range)

struct FsGeom {                  // window geometry of a search range (host + device)
  int pitch;                     // bytes per window row: == 32 or 96 (mod 128)
  int rows;                      // physical rows per copy (logical rows + 3)
  int copy_bytes;                // bytes per copy (multiple of 128)
  int slot_bytes;                // 4 copies
  int nslot;                     // items per CTA
  int total;                     // dynamic shared memory
};
__host__ __device__ inline FsGeom fs_geom(int R)
{
  FsGeom G;
  const int ncmax = 2 * R + 1 + FS_SMAX;
  int need = ncmax - 1 + 24;                       // last a-column + 6 words
  int p = (need + 15) & ~15;
  while ((p & 127) != 32 && (p & 127) != 96) p += 16;
  G.pitch = p;
  const int ngy = (ncmax + FS_K - 1) / FS_K;
  G.rows = FS_K * ngy + 15 + 3;
  G.copy_bytes = (G.rows * G.pitch + 127) & ~127;
  G.slot_bytes = 4 * G.copy_bytes;
  G.nslot = G.slot_bytes <= 40 * 1024 ? 2 : 1;
  G.total = G.nslot * G.slot_bytes;
  return G;
}

struct __align__(16) FsSlot {     // per-item state
  uint32_t cur[64];               // current MB, 16 rows x 4 words
  unsigned long long best[NPART]; // (cost << 20) | pos
  uint32_t Cw[20];                // packed filter constants, see cmap()
  int C16;                        // 16x16: -B-1
  unsigned short mxs[160], mys[160];   // per window column / row: lower bound of lambda*bits >> 5
  short pcx[NPART], pcy[NPART];   // centre (relative MV, quarter-pel)
  short ppx[NPART], ppy[NPART];   // predictor (quarter-pel)
  short psr[NPART];               // per-partition search range (pel)
  signed char pgrp[NPART];        // centre group of the partition (-1 inactive)
  signed char pex[NPART], pey[NPART];   // centre of the partition relative to its group's box origin (pel)
  short gx0[NPART], gy0[NPART], gx1[NPART], gy1[NPART];   // centre bounding box of group g (pel)
  int red[2][4];
  int ngroups;
  int active;                     // slot holds an item
  int mb, ref; unsigned base_lo, base_hi;
  // current group
  int ncx, ncy, ngy, gc, ncbA, ntaskA, npb, ntask;
  int ppxmin, ppxmax, ppymin, ppymax;
};

struct FsWarp { unsigned short sat[2][28]; };

// partition p -> index into the u16 view of Cw (word*2 + half); p == 0 -> -1 (scalar C16)
__device__ __forceinline__ int cmap(int p)
{
  if (p >= 25) { const int k = p - 25, b = k >> 2, x = k & 3; return 2 * (3 * b + (x & 1)) + (x >> 1); }      // CX[b] = (x0,x2), CY[b] = (x1,x3)
  if (p >= 17) { const int k = p - 17, bb = k >> 2, x = k & 3; return 2 * (12 + 3 * bb + (x & 1)) + (x >> 1); }
  if (p >= 9)  { const int k = p - 9; return 2 * (3 * (k >> 1) + 2) + (k & 1); }                              // CH[b] = (left,right)
  if (p >= 5)  { const int k = p - 5; return 2 * (14 + 3 * (k >> 1)) + (k & 1); }                             // CE[bb]
  if (p >= 3)  return 2 * 19 + (p - 3);                                                                       // CLR
  if (p >= 1)  return 2 * 18 + (p - 1);                                                                       // CTB
  return -1;
}

__device__ __forceinline__ void set_threshold(FsSlot &S, int p, unsigned long long key)
{
  const unsigned long long b = (key >> 20) >> 5;
  const int c = cmap(p);
  if (c >= 0) {
    const int B = b > 32765ull ? 32765 : (int)b;
    // low half: -B-1, high half two lower (absorbs the carries of the two 32-bit adds, see addmin2)
    reinterpret_cast<volatile uint16_t *>(S.Cw)[c] = (uint16_t)(-B - 1 - 2 * (c & 1));
  } else {
    *reinterpret_cast<volatile int *>(&S.C16) = b > 0x3fffffffull ? -0x40000000 : -(int)b - 1;
  }
}

// run = min(run, a + c) per signed halfword, as a 32-bit IMAD (a * one + c, `one` a run-time 1 so that it
// stays on the FMA pipe) followed by VIMNMX.S16x2.  Measured (b2me_ubench): VIADDMNMX.S16x2 occupies the
// ALU pipe as long as a VABSDIFF4, VIMNMX.S16x2 half as long, IMAD issues beside VABSDIFF4 for free.  The
// 32-bit add may carry from the low into the high half; the high halves of the constants are two lower to
// absorb that carry and the one of the final "+ m" (set_threshold), which keeps the filter conservative.
__device__ __forceinline__ uint32_t addmin2(uint32_t a, uint32_t c, uint32_t run, uint32_t one)
{
  const uint32_t t = a * one + c;
  uint32_t r;
  asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(t), "r"(run));
  return r;
}
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b)
{
  uint32_t t;
  asm("add.s16x2 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
  return t;
}

__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }

// Exact evaluation of two window candidates (one per half warp; the second may be absent) of slot S
// for every partition of group g.  Called by a whole warp.
template <int PITCH>
__device__ __noinline__ void fs_exact2(FsSlot &S, FsWarp &ws, const uint8_t *win, int copy_bytes, const uint32_t *pgt,
                                       int dx0, int dy0, int dx1, int dy1, bool valid1, int R, int g, int lambda_f)
{
  const int lane = threadIdx.x & 31, half = lane >> 4, k = lane & 15, bx = k & 3, by = k >> 2;
  {
    const int col = (half ? dx1 : dx0) + 4 * bx, row = (half ? dy1 : dy0) + 4 * by;
    const int c = col & 3;
    const uint8_t *p = win + c * copy_bytes + (row + c) * PITCH + (col >> 2) * 4;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) v = sad4(S.cur[(by * 4 + i) * 4 + bx], *reinterpret_cast<const uint32_t *>(p + i * PITCH), v);
    uint32_t t;
    t = __shfl_up_sync(0xffffffffu, v, 1, 16); if (bx >= 1) v += t;
    t = __shfl_up_sync(0xffffffffu, v, 2, 16); if (bx >= 2) v += t;
    t = __shfl_up_sync(0xffffffffu, v, 4, 16); if (by >= 1) v += t;
    t = __shfl_up_sync(0xffffffffu, v, 8, 16); if (by >= 2) v += t;
    ws.sat[half][(by + 1) * 5 + bx + 1] = (unsigned short)v;
  }
  __syncwarp();
  const int n = valid1 ? 2 * NPART : NPART;
  for (int idx = lane; idx < n; idx += 32) {
    const int h = idx >= NPART ? 1 : 0, p = idx - NPART * h;
    if (S.pgrp[p] != g) continue;
    const int ox = (h ? dx1 : dx0) - R - S.pex[p], oy = (h ? dy1 : dy0) - R - S.pey[p];   // displacement from the partition's own centre
    if (max(abs(ox), abs(oy)) > S.psr[p]) continue;
    const uint32_t q = pgt[p];
    const int x0 = q & 7, y0 = (q >> 4) & 7, x1 = (q >> 8) & 7, y1 = (q >> 12) & 7;
    const unsigned short *sat = ws.sat[h];
    const int sum = (int)sat[y1 * 5 + x1] - (int)sat[y0 * 5 + x1] - (int)sat[y1 * 5 + x0] + (int)sat[y0 * 5 + x0];
    const int mvx = S.pcx[p] + 4 * ox, mvy = S.pcy[p] + 4 * oy;
    const long long cost = ((long long)sum << 5) + (long long)lambda_f * (mvbits(mvx - S.ppx[p]) + mvbits(mvy - S.ppy[p]));
    const unsigned long long bestv = *reinterpret_cast<volatile unsigned long long *>(&S.best[p]);
    if ((unsigned long long)cost > (bestv >> 20)) continue;
    const unsigned long long key = ((unsigned long long)cost << 20) | (unsigned)spiral_index(ox, oy);
    if (key < bestv) {
      const unsigned long long old = atomicMin(&S.best[p], key);
      set_threshold(S, p, old < key ? old : key);
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void ld3(uint32_t (&r)[6], const uint8_t *p)
{
  const uint2 a = *reinterpret_cast<const uint2 *>(p), b = *reinterpret_cast<const uint2 *>(p + 8), c = *reinterpret_cast<const uint2 *>(p + 16);
  r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y; r[4] = c.x; r[5] = c.y;
}

// One lane-job: candidates (dx, dy0..dy0+K-1) and (dx+4, same rows); wb = lane address of window row dy0.
// Returns the pass mask: bit j = column a row j, bit K+j = column b row j.
template <int K, int PITCH>
__device__ __forceinline__ uint32_t fs_task(const FsSlot &S, const uint8_t *wb, uint32_t mxa, uint32_t mxb, const unsigned short *mys, uint32_t one)
{
  const uint4 *cur = reinterpret_cast<const uint4 *>(S.cur);
  uint32_t rw[K][6];
  uint32_t acc[2][K][4];
  uint32_t run[2][K], X0[2][K], Y0[2][K], E0[2][K];
  uint32_t pass = 0;
#pragma unroll
  for (int j = 0; j < K - 1; j++) ld3(rw[j], wb + j * PITCH);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    ld3(rw[(i + K - 1) % K], wb + (i + K - 1) * PITCH);
    const uint4 c = cur[i];
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int sl = (i + j) % K;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if ((i & 3) == 0) {
          acc[q][j][0] = sad4(c.x, rw[sl][q + 0], 0u);
          acc[q][j][1] = sad4(c.y, rw[sl][q + 1], 0u);
          acc[q][j][2] = sad4(c.z, rw[sl][q + 2], 0u);
          acc[q][j][3] = sad4(c.w, rw[sl][q + 3], 0u);
        } else {
          acc[q][j][0] = sad4(c.x, rw[sl][q + 0], acc[q][j][0]);
          acc[q][j][1] = sad4(c.y, rw[sl][q + 1], acc[q][j][1]);
          acc[q][j][2] = sad4(c.z, rw[sl][q + 2], acc[q][j][2]);
          acc[q][j][3] = sad4(c.w, rw[sl][q + 3], acc[q][j][3]);
        }
      }
    }
    if ((i & 3) == 3) {
      const int b = i >> 2;
      const uint32_t cx = ld_vol(&S.Cw[3 * b]), cy = ld_vol(&S.Cw[3 * b + 1]), ch = ld_vol(&S.Cw[3 * b + 2]);
      uint32_t cxv = 0, cyv = 0, ce = 0, ctb = 0, clr = 0; int c16 = 0;
      if (b & 1) { cxv = ld_vol(&S.Cw[12 + 3 * (b >> 1)]); cyv = ld_vol(&S.Cw[13 + 3 * (b >> 1)]); ce = ld_vol(&S.Cw[14 + 3 * (b >> 1)]); }
      if (b == 3) { ctb = ld_vol(&S.Cw[18]); clr = ld_vol(&S.Cw[19]); c16 = *reinterpret_cast<const volatile int *>(&S.C16); }
#pragma unroll
      for (int j = 0; j < K; j++) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const uint32_t X = acc[q][j][2] * 65536u + acc[q][j][0];
          const uint32_t Y = acc[q][j][3] * 65536u + acc[q][j][1];
          const uint32_t H = add2(X, Y);
          uint32_t r = (b == 0) ? 0x7fff7fffu : run[q][j];
          r = addmin2(X, cx, r, one);
          r = addmin2(Y, cy, r, one);
          r = addmin2(H, ch, r, one);
          if (b & 1) {
            const uint32_t XV = add2(X, X0[q][j]), YV = add2(Y, Y0[q][j]);
            const uint32_t E = add2(XV, YV);
            r = addmin2(XV, cxv, r, one);
            r = addmin2(YV, cyv, r, one);
            r = addmin2(E, ce, r, one);
            if (b == 1) E0[q][j] = E;
            else {
              const uint32_t top = __dp2a_lo(E0[q][j], 0x0101u, 0u), bot = __dp2a_lo(E, 0x0101u, 0u);
              const uint32_t TB = bot * 65536u + top, LR = add2(E0[q][j], E);
              r = addmin2(TB, ctb, r, one);
              r = addmin2(LR, clr, r, one);
              const uint32_t m = (q ? mxb : mxa) + mys[j];
              const int s = (int)(top + bot) + c16 + (int)m;
              const uint32_t t = r + m * 0x10001u;
              if ((t & 0x80008000u) != 0u || s < 0) pass |= 1u << (q * K + j);
            }
          } else { X0[q][j] = X; Y0[q][j] = Y; }
          run[q][j] = r;
        }
      }
    }
  }
  return pass;
}

template <int PITCH>
__global__ void __launch_bounds__(FS_NT, 3) k_sad_fs(const FsArgs a)
{
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ FsSlot SS[2];
  __shared__ FsWarp WS[FS_NW];
  __shared__ uint32_t pgt[NPART];
  __shared__ int s_next, s_err, s_nhits;
  constexpr int K = FS_K;
  const FsGeom G = fs_geom(a.R);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = a.R;
  const int nslot = G.nslot;
  const int wpw = PITCH >> 2;                      // words per window row

  if (tid < NPART) {
    const PartGeom gm = part_geom(tid);
    pgt[tid] = (gm.ox >> 2) | ((gm.oy >> 2) << 4) | (((gm.ox + gm.w) >> 2) << 8) | (((gm.oy + gm.h) >> 2) << 12);
  }
  for (int i = tid; i < FS_NW * 2 * 28; i += FS_NT) (&WS[0].sat[0][0])[i] = 0;
  if (tid == 0) { s_err = 0; s_nhits = 0; }

  const int nunits = (a.nitems + nslot - 1) / nslot;
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    __syncthreads();   // previous unit fully finished
    // ---- per-partition parameters, current MB (threads 0..63 -> slot 0, 64..127 -> slot 1) ----
    {
      const int sl = tid >> 6, t = tid & 63;
      FsSlot &S = SS[sl];
      const int item = unit * nslot + sl;
      const bool on = sl < nslot && item < a.nitems;
      if (t == 0) { S.active = on; S.ngroups = 0; S.ntask = 0; }
      if (on) {
        const int mb = a.mb_first + item / a.refs_per_mb, ref = a.ref_first + item % a.refs_per_mb;
        const int mbx = mb % a.mbw, mby = mb / a.mbw;
        const size_t base = (a.abs_index ? ((size_t)mb * a.nrefs + ref) : (size_t)item) * NPART;
        if (t == 0) { S.mb = mb; S.ref = ref; S.base_lo = (unsigned)base; S.base_hi = (unsigned)((unsigned long long)base >> 32); }
        if (t < NPART) {
          const int p = t;
          const bool act = (a.part_mask >> p) & 1ull;
          const PartGeom gm = part_geom(p);
          S.pcx[p] = a.center[(base + p) * 2]; S.pcy[p] = a.center[(base + p) * 2 + 1];
          S.ppx[p] = a.pred[(base + p) * 2];   S.ppy[p] = a.pred[(base + p) * 2 + 1];
          S.psr[p] = (short)(a.restrict_mode < 0 ? a.sr_override : block_search_range(R, a.restrict_mode, ref, gm.bt));
          S.pgrp[p] = act ? 0 : -1;
          const long long mm = a.min_mcost < 0 ? 0 : (a.min_mcost > (1ll << 42) ? (1ll << 42) : a.min_mcost);
          S.best[p] = ((unsigned long long)mm << 20);
          if (act && ((S.pcx[p] | S.pcy[p]) & 3)) s_err = 1;       // sub-pel centres are not a full-search input
        }
        S.cur[t] = *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + (t >> 2)) * a.cur_pitch + mbx * 16 + (t & 3) * 4);
      }
    }
    __syncthreads();
    // ---- cluster partitions whose centres fit in one FS_SMAX box (usually all of them) ----
    {
      const int sl = tid >> 6, t = tid & 63;
      FsSlot &S = SS[sl];
      if (S.active) {
        const bool act = t < NPART && S.pgrp[t] >= 0;
        const int cx = act ? (S.pcx[t] >> 2) : 0, cy = act ? (S.pcy[t] >> 2) : 0;
        const int x0 = __reduce_min_sync(0xffffffffu, act ? cx : 0x7fff), x1 = __reduce_max_sync(0xffffffffu, act ? cx : -0x7fff);
        const int y0 = __reduce_min_sync(0xffffffffu, act ? cy : 0x7fff), y1 = __reduce_max_sync(0xffffffffu, act ? cy : -0x7fff);
        if (lane == 0) { S.red[t >> 5][0] = x0; S.red[t >> 5][1] = x1; S.red[t >> 5][2] = y0; S.red[t >> 5][3] = y1; }
      }
    }
    __syncthreads();
    if ((tid & 63) == 0 && SS[tid >> 6].active) {
      FsSlot &S = SS[tid >> 6];
      const int bx0 = min(S.red[0][0], S.red[1][0]), bx1 = max(S.red[0][1], S.red[1][1]);
      const int by0 = min(S.red[0][2], S.red[1][2]), by1 = max(S.red[0][3], S.red[1][3]);
      if (bx1 - bx0 <= FS_SMAX && by1 - by0 <= FS_SMAX) {
        S.gx0[0] = bx0; S.gx1[0] = bx1; S.gy0[0] = by0; S.gy1[0] = by1; S.ngroups = bx1 >= bx0 ? 1 : 0;
      } else {        // general case: greedy clustering (serial, rare)
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
    const int maxg = max(SS[0].ngroups, SS[1].active ? SS[1].ngroups : 0);

    for (int g = 0; g < maxg; g++) {
      if (g) __syncthreads();
      // ---- group geometry + filter constants ----
      {
        const int sl = tid >> 6, t = tid & 63;
        FsSlot &S = SS[sl];
        const bool on = S.active && g < S.ngroups;
        if (on) {
          const bool ing = t < NPART && S.pgrp[t] == g;
          const int px = ing ? S.ppx[t] : 0, py = ing ? S.ppy[t] : 0;
          const int x0 = __reduce_min_sync(0xffffffffu, ing ? px : 0x7fff), x1 = __reduce_max_sync(0xffffffffu, ing ? px : -0x7fff);
          const int y0 = __reduce_min_sync(0xffffffffu, ing ? py : 0x7fff), y1 = __reduce_max_sync(0xffffffffu, ing ? py : -0x7fff);
          if (lane == 0) { S.red[t >> 5][0] = x0; S.red[t >> 5][1] = x1; S.red[t >> 5][2] = y0; S.red[t >> 5][3] = y1; }
          if (ing) { S.pex[t] = (signed char)((S.pcx[t] >> 2) - S.gx0[g]); S.pey[t] = (signed char)((S.pcy[t] >> 2) - S.gy0[g]); }
          if (t < 20) S.Cw[t] = 0;                 // partitions of other groups never pass
          if (t == 20) S.C16 = 0;
          if (t == 0) {
            const int spanx = S.gx1[g] - S.gx0[g], spany = S.gy1[g] - S.gy0[g];
            const int ncx = 2 * R + 1 + spanx, ncy = 2 * R + 1 + spany;
            const int ngy = (ncy + K - 1) / K;
            const int ncb = (ncx + 63) >> 6, wlast = ncx - 64 * (ncb - 1);
            const int ncbA = wlast <= FS_BMAX ? ncb - 1 : ncb;
            int npb = 0;
            if (ncbA < ncb) npb = wlast <= 4 ? wlast : (wlast <= 8 ? 4 : wlast - 4);
            S.ncx = ncx; S.ncy = ncy; S.ngy = ngy; S.gc = min(ngy - 1, (R + (spany >> 1)) / K);
            S.ncbA = ncbA; S.ntaskA = ncbA * ngy; S.npb = npb;
            S.ntask = ncbA * ngy + (npb * ngy + 31) / 32;
          }
        } else if (t == 0) S.ntask = 0;
      }
      __syncthreads();
      {
        const int sl = tid >> 6, t = tid & 63;
        FsSlot &S = SS[sl];
        const bool on = S.active && g < S.ngroups;
        if (on) {
          if (t < NPART && S.pgrp[t] == g) set_threshold(S, t, S.best[t]);
          if (t == 0) {
            S.ppxmin = min(S.red[0][0], S.red[1][0]); S.ppxmax = max(S.red[0][1], S.red[1][1]);
            S.ppymin = min(S.red[0][2], S.red[1][2]); S.ppymax = max(S.red[0][3], S.red[1][3]);
          }
        }
      }
      // ---- stage the windows (4 byte-shifted copies, copy c stored c rows lower) ----
      for (int sl = 0; sl < nslot; sl++) {
        FsSlot &S = SS[sl];
        if (!(S.active && g < S.ngroups)) continue;
        const int mbx = S.mb % a.mbw, mby = S.mb / a.mbw;
        const uint8_t *plane = a.planes + (size_t)S.ref * 16 * a.plane_size;     // integer plane [0][0]
        const int x0 = mbx * 16 + S.gx0[g] - R + PADX, y0 = mby * 16 + S.gy0[g] - R + PADY;
        const int nrows = G.rows - 3;
        const int needw = S.ncx + 15;
        uint8_t *wbase = smem + sl * G.slot_bytes;
        uint32_t *c0 = reinterpret_cast<uint32_t *>(wbase), *c1 = reinterpret_cast<uint32_t *>(wbase + G.copy_bytes + PITCH),
                 *c2 = reinterpret_cast<uint32_t *>(wbase + 2 * G.copy_bytes + 2 * PITCH), *c3 = reinterpret_cast<uint32_t *>(wbase + 3 * G.copy_bytes + 3 * PITCH);
        if (x0 >= 0 && y0 >= 0 && x0 + needw <= a.Wp && y0 + nrows <= a.Hp) {
          const int al = x0 & 3;
#pragma unroll 2
          for (int r = warp; r < nrows; r += FS_NW) {
            const uint32_t *grow = reinterpret_cast<const uint32_t *>(plane + (size_t)(y0 + r) * a.Wp + (x0 - al));
            for (int j = lane; j < wpw; j += 32) {
              const uint32_t g0 = grow[j], g1 = grow[j + 1], g2 = grow[j + 2];
              const uint32_t lo = al ? __funnelshift_r(g0, g1, 8 * al) : g0;      // window word j
              const uint32_t hi = al ? __funnelshift_r(g1, g2, 8 * al) : g1;      // window word j+1
              c0[r * wpw + j] = lo;
              c1[r * wpw + j] = __funnelshift_r(lo, hi, 8);
              c2[r * wpw + j] = __funnelshift_r(lo, hi, 16);
              c3[r * wpw + j] = __funnelshift_r(lo, hi, 24);
            }
          }
        } else {                      // window leaves the padded plane: per-pixel coordinate clamp
          for (int r = warp; r < nrows; r += FS_NW) {
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
      }
      __syncthreads();
      // ---- lower bound of the mv cost per window column / row: every partition of the group sees the
      //      same displacement 4*(g0 + d - R); bits is monotone in |mv - pred|, so the distance to the
      //      predictors' [min,max] interval bounds every partition's term from below ----
      for (int sl = 0; sl < nslot; sl++) {
        FsSlot &S = SS[sl];
        if (!(S.active && g < S.ngroups)) continue;
        const int ncx = S.ncx, ncy = S.ncy;
        for (int i = tid; i < ncx + ncy + K; i += FS_NT) {
          const bool isy = i >= ncx;
          const int d = isy ? i - ncx : i;
          const int mv = 4 * ((isy ? S.gy0[g] : S.gx0[g]) + d - R);
          const int lo = isy ? S.ppymin : S.ppxmin, hi = isy ? S.ppymax : S.ppxmax;
          const int dist = max(0, max(lo - mv, mv - hi));
          const long long v = ((long long)a.lambda_f * mvbits(dist)) >> 5;
          (isy ? S.mys : S.mxs)[d] = (unsigned short)(v > FS_MCAP ? FS_MCAP : v);
        }
      }
      if (tid == 0) s_next = 0;
      __syncthreads();
      // ---- initial bounds: exact pre-pass over the centres' box ----
      {
        int total[2] = {0, 0}, pw[2] = {1, 1}, xlo[2] = {0, 0}, ylo[2] = {0, 0};
        for (int sl = 0; sl < nslot; sl++) {
          FsSlot &S = SS[sl];
          if (!(S.active && g < S.ngroups)) continue;
          xlo[sl] = max(0, R - FS_PRE); ylo[sl] = max(0, R - FS_PRE);
          pw[sl] = min(S.ncx - 1, R + S.ncx - (2 * R + 1) + FS_PRE) - xlo[sl] + 1;
          const int ph = min(S.ncy - 1, R + S.ncy - (2 * R + 1) + FS_PRE) - ylo[sl] + 1;
          total[sl] = pw[sl] * ph;
        }
        const int npair0 = (total[0] + 1) >> 1, npair1 = (total[1] + 1) >> 1;
        for (int pr = warp; pr < npair0 + npair1; pr += FS_NW) {
          const int sl = pr >= npair0 ? 1 : 0, i0 = 2 * (pr - (sl ? npair0 : 0)), i1 = i0 + 1;
          const bool v1 = i1 < total[sl];
          const int dx0 = xlo[sl] + i0 % pw[sl], dy0 = ylo[sl] + i0 / pw[sl];
          const int dx1 = v1 ? xlo[sl] + i1 % pw[sl] : dx0, dy1 = v1 ? ylo[sl] + i1 / pw[sl] : dy0;
          fs_exact2<PITCH>(SS[sl], WS[warp], smem + sl * G.slot_bytes, G.copy_bytes, pgt, dx0, dy0, dx1, dy1, v1, R, g, a.lambda_f);
        }
      }
      __syncthreads();
      // ---- main pass: warps pull tasks until both items are exhausted ----
      const int nt0 = SS[0].ntask, nt1 = nslot > 1 ? SS[1].ntask : 0;
      int nh = 0;
      for (;;) {
        int T = 0;
        if (lane == 0) T = atomicAdd(&s_next, 1);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= nt0 + nt1) break;
        // interleave the two items so that both advance from their centres outwards
        int sl, t;
        if (T < 2 * min(nt0, nt1)) { sl = T & 1; t = T >> 1; }
        else { sl = nt0 > nt1 ? 0 : 1; t = T - min(nt0, nt1); }
        FsSlot &S = SS[sl];
        const uint8_t *win = smem + sl * G.slot_bytes;
        int dxa, dy0; bool va, vb;
        if (t < S.ntaskA) {
          const int cb = t / S.ngy, k = t - cb * S.ngy;
          const int gc = S.gc, lo = gc, hi = S.ngy - 1 - gc, mn = min(lo, hi);
          const int gy = k <= 2 * mn ? ((k & 1) ? gc + ((k + 1) >> 1) : gc - (k >> 1)) : (lo > hi ? gc - (k - hi) : gc + (k - lo));
          dxa = 64 * cb + 8 * (lane >> 2) + (lane & 3); dy0 = gy * K;
          va = dxa < S.ncx; vb = dxa + 4 < S.ncx;
        } else {
          const int job = (t - S.ntaskA) * 32 + lane;
          const int i = job / S.ngy, gy = job - i * S.ngy;
          dxa = 64 * S.ncbA + 8 * (i >> 2) + (i & 3); dy0 = gy * K;
          va = i < S.npb && dxa < S.ncx; vb = va && dxa + 4 < S.ncx;
          if (!va) dy0 = 0;
        }
        if (!va) dxa = 0;
        const int c = dxa & 3;
        const uint8_t *wb = win + c * G.copy_bytes + (dy0 + c) * PITCH + (dxa >> 2) * 4;
        uint32_t pass = fs_task<K, PITCH>(S, wb, S.mxs[dxa], S.mxs[dxa + 4], S.mys + dy0, (uint32_t)a.one);
        uint32_t vm = 0;
#pragma unroll
        for (int j = 0; j < K; j++) if (dy0 + j < S.ncy) vm |= (va ? 1u << j : 0u) | (vb ? 1u << (K + j) : 0u);
        pass &= vm;
        if (__any_sync(0xffffffffu, pass != 0)) {
          for (int b = 0; b < 2 * K; b++) {
            uint32_t m = __ballot_sync(0xffffffffu, (pass >> b) & 1u);
            while (m) {
              const int l0 = __ffs(m) - 1; m &= m - 1;
              const bool v1 = m != 0;
              const int l1 = v1 ? __ffs(m) - 1 : l0; m &= m - 1;
              const int ddx = b >= K ? 4 : 0, ddy = b >= K ? b - K : b;
              const int ex0 = __shfl_sync(0xffffffffu, dxa, l0) + ddx, ey0 = __shfl_sync(0xffffffffu, dy0, l0) + ddy;
              const int ex1 = __shfl_sync(0xffffffffu, dxa, l1) + ddx, ey1 = __shfl_sync(0xffffffffu, dy0, l1) + ddy;
              fs_exact2<PITCH>(S, WS[warp], win, G.copy_bytes, pgt, ex0, ey0, ex1, ey1, v1, R, g, a.lambda_f);
              nh += v1 ? 2 : 1;
            }
          }
        }
      }
      if (lane == 0 && nh) atomicAdd(&s_nhits, nh);
    }
    __syncthreads();
    // ---- results ----
    {
      const int sl = tid >> 6, p = tid & 63;
      FsSlot &S = SS[sl];
      if (S.active && p < NPART && ((a.part_mask >> p) & 1ull)) {
        const size_t base = ((size_t)S.base_hi << 32) | S.base_lo;
        const unsigned long long key = S.best[p];
        const int pos = (int)(key & 0xfffffull);
        int sx, sy; spiral_xy(pos, &sx, &sy);
        a.mv_int[(base + p) * 2]     = (int16_t)(S.pcx[p] + 4 * sx);
        a.mv_int[(base + p) * 2 + 1] = (int16_t)(S.pcy[p] + 4 * sy);
        long long cost = (long long)(key >> 20);
        if (pos == 0 && cost == (1ll << 42) && a.min_mcost > (1ll << 42)) cost = a.min_mcost;   // bound never beaten
        a.cost_int[base + p] = cost;
      }
    }
    if (tid == 0 && s_err) { *a.errflag = 1; s_err = 0; }
    if (tid == 0 && a.stats) {
      atomicAdd(&a.stats[0], (unsigned long long)s_nhits);
      atomicAdd(&a.stats[1], (unsigned long long)(SS[0].ngroups + (SS[1].active ? SS[1].ngroups : 0)));
      atomicAdd(&a.stats[2], (unsigned long long)(1 + (SS[1].active ? 1 : 0)));
      s_nhits = 0;
    }
  }
}

cudaError_t launch_sad_fs(const FsArgs &a, int sm_count, cudaStream_t s, int *smem_bytes_out)
{
  const FsGeom G = fs_geom(a.R);
  static int configured96 = 0, configured160 = 0;
  const int nunits = (a.nitems + G.nslot - 1) / G.nslot;
  cudaError_t e;
  if (G.pitch == 96) {
    if (configured96 < G.total) {
      e = cudaFuncSetAttribute(k_sad_fs<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, G.total);
      if (e != cudaSuccess) return e;
      configured96 = G.total;
    }
    k_sad_fs<96><<<nunits, FS_NT, G.total, s>>>(a);
  } else if (G.pitch == 160) {
    if (configured160 < G.total) {
      e = cudaFuncSetAttribute(k_sad_fs<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, G.total);
      if (e != cudaSuccess) return e;
      configured160 = G.total;
    }
    k_sad_fs<160><<<nunits, FS_NT, G.total, s>>>(a);
  } else {
    return cudaErrorInvalidValue;
  }
  if (smem_bytes_out) *smem_bytes_out = G.total;
  (void)sm_count;
  return cudaGetLastError();
}

}  // namespace b2
