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
// Work decomposition.  Persistent CTAs, two window buffers each.  Buffer b of CTA i walks the items
// i + (2k+b)*gridDim; a "unit" is one (item, centre group) and lives in one buffer: its window (TMA),
// per-partition state and filter constants.  A warp-task = 64 window columns x K rows of candidates:
// lane l owns the column pair (dx, dx+4), dx = 8*(l>>2) + (l&3), and walks K consecutive rows
// dy0..dy0+K-1 in lock step with a sliding window of K reference rows in registers, so a reference row
// costs three LDS.64 per 2K candidates and a current row one broadcast LDS.128.  Worker warps pull tasks of
// whichever buffer is ready from a shared counter (the claim of their NEXT tasks is issued before they run the
// current ones) and never meet at a CTA barrier; a dedicated producer warp writes the results of a finished unit
// and sets up the buffer's next unit (parameters, TMA of the window, tables, exact pre-pass) while the workers
// drain the other buffer.
//
// Filter.  After every 4 rows the 4x4 SADs of a block row are packed two per register (IMAD on the
// FMA pipe), tree-summed (VIADD.16x2), and folded into one running minimum per candidate:
//        run = min(run, sad_p - B_p - 1),  B_p = best_p >> 5      (IMAD add + VIMNMX.S16x2)
// At the end  run + m < 0  (m = floor(lambda*minbits/32), a lower bound of the candidate's mv
// cost over the partitions' predictors) is a NECESSARY condition for (cost,pos) < best of some
// partition, so nothing that could win is dropped.  Survivors (rare) are re-evaluated exactly by
// the warp itself (16 lanes per candidate for the 4x4 SADs, a summed-area table, then one lane
// per (candidate, partition)) and lower best_p with a 64-bit atomicMin on (cost << 20 | spiral pos).
//
// Shared-memory window: 4 byte-shifted copies (copy c holds the window shifted left by c bytes) so
// that the 16+4 pixels of a column pair are three aligned 64-bit words; copy c is stored c rows
// lower so that, with a row pitch of 24 (mod 32) words, the four copies start 8 banks apart while
// every copy stays 128-byte aligned (TMA destination rule).  Each copy is one cp.async.bulk.tensor
// box.  A TMA box must START on a 16-byte boundary in global memory (a box at an odd byte column raises
// "illegal instruction", tools/tma_probe), so the reference's search plane (edge-replicated integer
// picture with a pad of R+32, k_search_plane) is kept in HBM as 16 byte-shifted planes,
// plane_s[y][x] = P[y][x+s]: copy c of a window at column x0 is the box at column (x0+c)&~15 of plane
// (x0+c)&15.  Windows that leave the search plane (centres more than 32 pel outside the picture) are
// staged by the setting-up warp with per-pixel clamps.
#include "b2_common.cuh"
#include <cstdlib>
#include "b2_ctx.h"

namespace b2 {

#ifndef FS_KV
#define FS_KV 4
#endif
constexpr int FS_K = FS_KV;         // candidate rows per task (lock step); fs_task4 is written for 4
#ifndef FS_PREV
#define FS_PREV -1
#endif
constexpr int FS_PRE = FS_PREV;       // >= 0: radius of the exact pre-pass around the centres' box; -1: one centre only (initial bounds)
#ifndef FS_SMAXV
#define FS_SMAXV 7
#endif
constexpr int FS_SMAX = FS_SMAXV;       // partitions whose centres lie within a box of this many pel share one window pass
#ifndef FS_COLDPFV
#define FS_COLDPFV 1
#endif
constexpr bool FS_COLDPF = FS_COLDPFV != 0;     // fs_cold_lane: software prefetch of the next candidate's packed sums
#ifndef FS_ROLL4V
#define FS_ROLL4V 0
#endif
constexpr bool FS_ROLL4 = FS_ROLL4V != 0;       // fs_task4's loop rolled over 4-row block rows (instruction-cache footprint)
#ifndef FS_REDUCEDV
#define FS_REDUCEDV 0
#endif
constexpr bool FS_REDUCED = FS_REDUCEDV != 0;   // reduced task routines for the odd column / odd row of a +-R window (fs_task4<1, 4>, <2, 1>)
constexpr int FS_BMAX = 12;      // a trailing column block of at most this width is walked row-major (type B)
constexpr int FS_MCAP = 2047;    // cap of each half of the mv-cost lower bound (keeps packed sums in range)

__host__ __device__ inline FsGeom fs_geom(int R)
{
  FsGeom G;
  const int ncmax = 2 * R + 1 + FS_SMAX;
  int need = ncmax - 1 + 24;                       // last a-column + 6 words
  int p = (need + 15) & ~15;
  while ((p & 127) != 32 && (p & 127) != 96) p += 16;
  G.pitch = p;
  const int ngy = (ncmax + FS_K - 1) / FS_K;
  G.rows = ncmax + 15 + 3;                         // rows that valid candidates read (+3 for the copies' row shift)
  G.copy_bytes = (G.rows * G.pitch + 127) & ~127;
  G.slot_bytes = 4 * G.copy_bytes;
  // masked candidates of the last row group read up to FS_K*ngy+15+3 rows: harmless garbage from the next
  // copy / buffer; the slack keeps those reads inside the allocation at the very end
  const int over = (FS_K * ngy + 15 + 3 - G.rows) * G.pitch;
  G.total = 2 * G.slot_bytes + ((over + 127) & ~127);
  G.threads = G.pitch == 96 ? 160 : 416;    // workers + one producer warp
  return G;
}

template <int PITCH>
struct __align__(16) FsSlotT {    // one unit = (item, centre group): everything but the window
  static constexpr int NTAB = PITCH == 96 ? 88 : 152;
  uint32_t cur[64];               // current MB, 16 rows x 4 words
  unsigned long long best[NPART]; // (cost << 20) | pos
  uint32_t Cw[20];                // packed filter constants, see cmap()
  int C16;                        // 16x16: -B-1
  unsigned short mxs[NTAB], mys[NTAB];   // per window column / row: lower bound of lambda*bits >> 5 (also clustering scratch)
  short pcx[NPART], pcy[NPART];   // centre (relative MV, quarter-pel)
  short ppx[NPART], ppy[NPART];   // predictor (quarter-pel)
  short psr[NPART];               // per-partition search range (pel)
  signed char pgrp[NPART];        // centre group of the partition (-1 inactive)
  signed char pex[NPART], pey[NPART];   // centre of the partition relative to its group's box origin (pel)
  int item, mb, ref, g, ngroups; unsigned base_lo, base_hi;
  int seedx, seedy;               // window coordinates of the seed candidate the pre-pass evaluated (-1: none)
  int prex, prey, prew, preh;     // window rectangle the pre-pass evaluated exactly
  int gx0, gy0, spanx, spany;     // the group's centre box (pel)
  int wx0, wy0, inside;           // window origin in the search plane; inside: the TMA path applies
  int ncx, ncy, ngy, gc, ncbA, ntaskA, npb, ntask;
  static constexpr int NTT = PITCH == 96 ? 40 : 96;
  unsigned short ttab[NTT];       // type-A task t -> dy0 | column block << 8, centre rows first
};

struct FsCtl {                    // per window buffer: the pipeline between the producer warp and the workers
  unsigned long long mbar;        // TMA completion barrier
  int ready_epoch;                // epoch whose window + unit are complete (written last, read first)
  int finished_epoch;             // epoch whose tasks are all done (written by the worker that completed the last one)
  int next;                       // (epoch << 20) | tasks claimed
  int done;                       // tasks completed in the current epoch
  int ntask;                      // tasks of the current epoch
  int slot;                       // unit slot attached to this buffer
  int ended;                      // no more units for this buffer
  int epoch, tma_uses;
};

#ifndef FS_DENSEV
#define FS_DENSEV 32
#endif
constexpr int FS_DENSE = FS_DENSEV;     // candidates of a warp passing the cold path's test at once from which the dense collapse runs
constexpr int FS_RCAP = 24;      // survivor records per warp and task
struct __align__(16) FsWarp {
  unsigned short sat[2][28];     // fs_exact2's summed-area tables (row 0 / column 0 stay zero)
  uint32_t rec[FS_RCAP];         // records of the dense collapse; ALSO words 0..23 of the sparse cold path's staging (X, Y, X0 of one lane)
  int nrec;                      // records appended since the last flush (may exceed FS_RCAP: overflow)
};
#ifndef FS_SPARSEV
#define FS_SPARSEV 8
#endif
constexpr int FS_SPARSE = FS_SPARSEV;     // at most this many lanes of a warp pass the group filter at once: their packed sums cross through shared memory

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

template <class SLOT>
__device__ __forceinline__ void set_threshold(SLOT &S, int p, unsigned long long key)
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
template <int PITCH, class SLOT>
__device__ __noinline__ void fs_exact2(SLOT &S, FsWarp &ws, const uint8_t *win, int copy_bytes, const uint32_t *pgt,
                                       int dx0, int dy0, int dx1, int dy1, bool valid1, int R, int g, int lambda_f, const FsArgs *ga = nullptr)
{
  const int lane = threadIdx.x & 31, half = lane >> 4, k = lane & 15, bx = k & 3, by = k >> 2;
  {
    const int col = (half ? dx1 : dx0) + 4 * bx, row = (half ? dy1 : dy0) + 4 * by;
    uint32_t v = 0;
    if (ga) {                                        // straight from the search planes (the producer's pre-pass, before the window exists):
      const int X = S.wx0 + col, Y = S.wy0 + row;    // plane X & 3 holds the four bytes from column X as one aligned word
      const int c = X & 3;
      const uint8_t *p = ga->spl + ((size_t)(S.ref * 16 + c) * ga->Hq + Y) * ga->Wq + (X - c);
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = __ldg(reinterpret_cast<const uint32_t *>(p + (size_t)i * ga->Wq));
#pragma unroll
      for (int i = 0; i < 4; i++) v = sad4(S.cur[(by * 4 + i) * 4 + bx], w[i], v);
    } else {
      const int c = col & 3;
      const uint8_t *p = win + c * copy_bytes + (row + c) * PITCH + (col >> 2) * 4;
#pragma unroll
      for (int i = 0; i < 4; i++) v = sad4(S.cur[(by * 4 + i) * 4 + bx], *reinterpret_cast<const uint32_t *>(p + i * PITCH), v);
    }
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

// ---- survivor records ---------------------------------------------------------------------------------------
// Records of the dense collapse (fs_dense_collapse): a (candidate, partition) pair whose SAD is known,
//   word = sad (16) | partition << 16 | (q * 4 + j) << 22 | lane << 25        (q: column of the lane's pair, j: row of the task)
// evaluated exactly (mv cost, spiral position, 64-bit atomicMin) by fs_process_records, one lane per record.
__device__ __noinline__ void fs_rec(FsWarp &ws, uint32_t word)
{
  const int i = atomicAdd(&ws.nrec, 1);
  if (i < FS_RCAP) ws.rec[i] = word;
}

// Exact evaluation of the task's survivor records, one lane per record: mv cost with the partition's own predictor, spiral
// position, strict (cost, position) order through a 64-bit atomicMin (me_fullsearch.c:83-93), new filter constant.
template <class SLOT>
__device__ __noinline__ void fs_process_records(SLOT &S, FsWarp &ws, int n, int t, int R, int g, int lambda_f, int first, int stride)
{
  for (int i = first; i < n; i += stride) {
    const uint32_t w = ws.rec[i];
    const int sad = w & 0xffff, p = (w >> 16) & 63, qj = (w >> 22) & 7, l = w >> 25;
    int dxa, dy0;
    if (t < S.ntaskA) { const int e = S.ttab[t]; dxa = 64 * (e >> 8) + 8 * (l >> 2) + (l & 3); dy0 = e & 255; }
    else { const int job = (t - S.ntaskA) * 32 + l; const int ii = job / S.ngy, gy = job - ii * S.ngy; dxa = 64 * S.ncbA + 8 * (ii >> 2) + (ii & 3); dy0 = gy * FS_K; }
    const int dx = dxa + 4 * (qj >> 2), dy = dy0 + (qj & 3);
    if (S.pgrp[p] != g) continue;
    const int ox = dx - R - S.pex[p], oy = dy - R - S.pey[p];   // displacement from the partition's own centre
    if (max(abs(ox), abs(oy)) > S.psr[p]) continue;
    const int mvx = S.pcx[p] + 4 * ox, mvy = S.pcy[p] + 4 * oy;
    const long long cost = ((long long)sad << 5) + (long long)lambda_f * (mvbits(mvx - S.ppx[p]) + mvbits(mvy - S.ppy[p]));
    const unsigned long long bestv = *reinterpret_cast<volatile unsigned long long *>(&S.best[p]);
    if ((unsigned long long)cost > (bestv >> 20)) continue;
    const unsigned long long key = ((unsigned long long)cost << 20) | (unsigned)spiral_index(ox, oy);
    if (key < bestv) {
      const unsigned long long old = atomicMin(&S.best[p], key);
      set_threshold(S, p, old < key ? old : key);
    }
  }
}

struct FsTaskCtx { int t, R, g, lambda_f; };

__device__ __forceinline__ uint32_t minu2(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// Dense cold path (many lanes of the warp passed at once: the bounds are still loose, e.g. the first tasks of an item whose
// predictor is wrong).  Instead of recording every passing (candidate, partition) pair -- hundreds -- the passing lanes find,
// per partition of this level, the ONE candidate of the task with the smallest SAD + mv-cost bound (a minimum per lane that
// carries the candidate's index, a warp min-reduction per partition), record it and evaluate those <= 23 records exactly at once.  The bounds drop to about
// their final values for this task's candidates, and the ordinary per-candidate test that follows finds few survivors.
// Called by the lanes of `mask` (converged); vmm: the lane's candidates that are valid and outside the pre-pass box.
template <class SLOT>
__device__ __noinline__ void fs_dense_collapse(SLOT &S, FsWarp &ws, const uint32_t *buf, int b, int dxa, int dy0, uint32_t vmm, uint32_t lanebits,
                                               FsTaskCtx tc, uint32_t mask)
{
  const int lane = threadIdx.x & 31;
  const int nk = b == 3 ? 11 : 9;
  uint32_t M[11]; uint32_t M16 = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < 11; k++) M[k] = 0xffffffffu;
  // per lane: minimum over its candidates of ((SAD + mv-cost bound) << 3 | candidate) for every partition of this level
  uint32_t Mh[11];
#pragma unroll
  for (int k = 0; k < 11; k++) Mh[k] = 0xffffffffu;
  uint32_t nX = buf[0], nY = buf[8], nX0 = buf[16], nY0 = buf[24], nE0 = buf[32];      // next candidate's words, requested one iteration ahead (see fs_cold_lane)
#pragma unroll 1
  for (int c = 0; c < 8; c++) {
    uint32_t X = nX, Y = nY, X0 = nX0, Y0 = nY0, E0 = nE0;
    if (FS_COLDPF) { if (c < 7) { nX = buf[c + 1]; nY = buf[9 + c]; nX0 = buf[17 + c]; nY0 = buf[25 + c]; nE0 = buf[33 + c]; } }
    else { X = buf[c]; Y = buf[8 + c]; X0 = buf[16 + c]; Y0 = buf[24 + c]; E0 = buf[32 + c]; }
    if (!((vmm >> c) & 1u)) continue;
    const uint32_t m = (uint32_t)S.mxs[dxa + 4 * (c >> 2)] + (uint32_t)S.mys[dy0 + (c & 3)];
    uint32_t Q[11];
    Q[0] = X0; Q[1] = Y0; Q[2] = X0 + Y0; Q[3] = X; Q[4] = Y; Q[5] = X + Y; Q[6] = X + X0; Q[7] = Y + Y0; Q[8] = Q[6] + Q[7];
    const uint32_t top = (E0 & 0xffffu) + (E0 >> 16), bot = (Q[8] & 0xffffu) + (Q[8] >> 16);
    Q[9] = (bot << 16) | top; Q[10] = E0 + Q[8];
#pragma unroll
    for (int k = 0; k < 11; k++) {
      if (k < nk) {
        M[k] = min(M[k], ((((Q[k] & 0xffffu) + m) << 3) | (uint32_t)c));
        Mh[k] = min(Mh[k], ((((Q[k] >> 16) + m) << 3) | (uint32_t)c));
      }
    }
    if (b == 3) M16 = min(M16, ((top + bot + m) << 3) | (uint32_t)c);
  }
  // per partition: the warp's minimum; its owner records the candidate (only partitions whose bound the minimum can reach)
#pragma unroll
  for (int k = 0; k < 11; k++) {
    if (k < nk) {
      const int a = b - 1, bb = b >> 1;
      const int cidx = k < 3 ? 3 * a + k : (k < 6 ? 3 * b + (k - 3) : (k < 9 ? 12 + 3 * bb + (k - 6) : 18 + (k - 9)));
      const int plo = k == 0 ? 25 + 4 * a : k == 1 ? 26 + 4 * a : k == 2 ? 9 + 2 * a : k == 3 ? 25 + 4 * b : k == 4 ? 26 + 4 * b : k == 5 ? 9 + 2 * b
                    : k == 6 ? 17 + 4 * bb : k == 7 ? 18 + 4 * bb : k == 8 ? 5 + 2 * bb : k == 9 ? 1 : 3;
      const int phi = (k == 2 || k == 5 || k == 8 || k >= 9) ? plo + 1 : plo + 2;
      const uint32_t cc = ld_vol(&S.Cw[cidx]);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const uint32_t mine = h ? Mh[k] : M[k];
        const uint32_t w = __reduce_min_sync(mask, mine > 0x7ffffffu ? 0xffffffffu : ((mine << 5) | (uint32_t)lane));
        const uint32_t v = w >> 8, th = h ? cc >> 16 : cc & 0xffffu;          // v = SAD + bound;  th = (-B - 1 [- 2]) mod 2^16
        if (w != 0xffffffffu && (((v + th) & 0x8000u) != 0u) && (w & 31u) == (uint32_t)lane) {
          const int c = mine & 7;
          const uint32_t m = (uint32_t)S.mxs[dxa + 4 * (c >> 2)] + (uint32_t)S.mys[dy0 + (c & 3)];
          fs_rec(ws, ((mine >> 3) - m) | ((uint32_t)(h ? phi : plo) << 16) | lanebits | ((uint32_t)c << 22));
        }
      }
    }
  }
  if (b == 3) {
    const uint32_t w = __reduce_min_sync(mask, M16 > 0x7ffffffu ? 0xffffffffu : ((M16 << 5) | (uint32_t)lane));
    if (w != 0xffffffffu && (int)(w >> 8) + *reinterpret_cast<const volatile int *>(&S.C16) < 0 && (w & 31u) == (uint32_t)lane) {
      const int c = M16 & 7;
      const uint32_t m = (uint32_t)S.mxs[dxa + 4 * (c >> 2)] + (uint32_t)S.mys[dy0 + (c & 3)];
      fs_rec(ws, ((M16 >> 3) - m) | lanebits | ((uint32_t)c << 22));
    }
  }
  __syncwarp(mask);
  {
    const int n1 = *reinterpret_cast<volatile int *>(&ws.nrec);
    if (n1 > 0 && n1 <= FS_RCAP) {
      fs_process_records(S, ws, n1, tc.t, tc.R, tc.g, tc.lambda_f, __popc(mask & ((1u << lane) - 1u)), __popc(mask));
      __syncwarp(mask);
      if ((mask & ((1u << lane) - 1u)) == 0u) *reinterpret_cast<volatile int *>(&ws.nrec) = 0;
      __syncwarp(mask);
    }
  }
}

// The cold path of one lane at the end of the odd block row b: a quick packed test per candidate (the hot path's filter without
// the minimum over the candidates), the records of those that pass.  buf[5][8]: X, Y, X0, Y0, E0 per candidate q*4+j.
// Returns the pass bits of the candidates.
template <class SLOT>
__device__ __noinline__ uint32_t fs_cold_lane(SLOT &S, FsWarp &ws, const uint32_t *buf, int b, int dxa, int dy0, uint32_t vm, uint32_t lanebits, FsTaskCtx tc)
{
  const int a = b - 1, bb = b >> 1, R = tc.R;
  uint32_t vmm = 0;                                  // valid candidates outside the centres' box (that box was evaluated exactly by the producer's pre-pass)
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int x = dxa + 4 * (c >> 2), y = dy0 + (c & 3);
    if (((vm >> c) & 1u) && !(x >= S.prex && x < S.prex + S.prew && y >= S.prey && y < S.prey + S.preh) && !(x == S.seedx && y == S.seedy)) vmm |= 1u << c;
  }
  const uint32_t mask = __activemask();              // the lanes in the cold path (converged here)
  uint32_t pass = 0;
#pragma unroll 1
  for (int round = 0; round < 2; round++) {
    const uint32_t cx0 = ld_vol(&S.Cw[3 * a]), cy0 = ld_vol(&S.Cw[3 * a + 1]), ch0 = ld_vol(&S.Cw[3 * a + 2]);
    const uint32_t cx = ld_vol(&S.Cw[3 * b]), cy = ld_vol(&S.Cw[3 * b + 1]), ch = ld_vol(&S.Cw[3 * b + 2]);
    const uint32_t cxv = ld_vol(&S.Cw[12 + 3 * bb]), cyv = ld_vol(&S.Cw[13 + 3 * bb]), ce = ld_vol(&S.Cw[14 + 3 * bb]);
    const uint32_t ctb = ld_vol(&S.Cw[18]), clr = ld_vol(&S.Cw[19]);
    const int c16 = *reinterpret_cast<const volatile int *>(&S.C16);
    pass = 0;
    // buf is local memory, and with the SM's shared-memory carve-out local lines miss L1: every read is an L2 round trip.  The five
    // words of candidate c + 1 are requested before candidate c is tested (FS_COLDPF), so the loop pays the latency once, not 8 times.
    uint32_t nX = buf[0], nY = buf[8], nX0 = buf[16], nY0 = buf[24], nE0 = buf[32];
#pragma unroll 1
    for (int c = 0; c < 8; c++) {
      uint32_t X = nX, Y = nY, X0 = nX0, Y0 = nY0, E0 = nE0;
      if (FS_COLDPF) { if (c < 7) { nX = buf[c + 1]; nY = buf[9 + c]; nX0 = buf[17 + c]; nY0 = buf[25 + c]; nE0 = buf[33 + c]; } }
      else { X = buf[c]; Y = buf[8 + c]; X0 = buf[16 + c]; Y0 = buf[24 + c]; E0 = buf[32 + c]; }
      if (!((vmm >> c) & 1u)) continue;
      const uint32_t m = (uint32_t)S.mxs[dxa + 4 * (c >> 2)] + (uint32_t)S.mys[dy0 + (c & 3)];
      const uint32_t XV = X + X0, YV = Y + Y0, E = XV + YV;
      uint32_t t = __vimin3_s16x2(X0 + cx0, Y0 + cy0, X0 + Y0 + ch0);
      t = __vimin3_s16x2(t, X + cx, Y + cy);
      t = __vimin3_s16x2(t, X + Y + ch, XV + cxv);
      t = __vimin3_s16x2(t, YV + cyv, E + ce);
      int s16 = 0;
      if (b == 3) {
        const uint32_t top = (E0 & 0xffffu) + (E0 >> 16), bot = (E & 0xffffu) + (E >> 16);
        t = __vimin3_s16x2(t, ((bot << 16) | top) + ctb, E0 + E + clr);
        s16 = (int)(top + bot) + c16 + (int)m;
      }
      if ((((t + m * 0x10001u) & 0x80008000u) != 0u) || s16 < 0) pass |= 1u << c;
    }
    __syncwarp(mask);
    // many candidates of the warp pass at once: the bounds are still loose (the first tasks of an item whose predictor is
    // wrong).  Collapse them (one exact evaluation per partition), then test again against the new bounds.
    if (round || (int)__reduce_add_sync(mask, (unsigned)__popc(pass)) < FS_DENSE) break;
    fs_dense_collapse(S, ws, buf, b, dxa, dy0, vmm, lanebits, tc, mask);
  }
  return pass;
}

// Sparse cold path (the usual one: one to FS_SPARSE lanes of the warp passed the group filter).  The passing lane's packed sums
// (X, Y, X0, Y0, E0 of its eight candidates) were staged in shared memory (ws.rec: 24 words, stg: 16 words); lane c of every
// group of eight tests candidate c -- the hot path's filter without the minimum over the candidates -- so the eight tests run
// side by side instead of one after the other (measured: the serial version read its operands back from LOCAL memory, which
// with 229 KB of the SM carved out as shared memory misses L1: 8 260 warp-cycles per entry, 11 % of the kernel's warp time).
// dxa, dy0, vm: the passing lane's (uniform).  Returns the pass bits of its candidates (uniform).
template <class SLOT>
__device__ __noinline__ uint32_t fs_cold_test8(SLOT &S, const FsWarp &ws, const uint32_t *stg, int b, int dxa, int dy0, uint32_t vm, int R)
{
  const int c = threadIdx.x & 7;
  const int a = b - 1, bb = b >> 1;
  const int x = dxa + 4 * (c >> 2), y = dy0 + (c & 3);
  // valid candidates outside the rectangle / seed the producer's pre-pass evaluated exactly
  const bool v = ((vm >> c) & 1u) && !(x >= S.prex && x < S.prex + S.prew && y >= S.prey && y < S.prey + S.preh) && !(x == S.seedx && y == S.seedy);
  const uint32_t X = ld_vol(&ws.rec[c]), Y = ld_vol(&ws.rec[8 + c]), X0 = ld_vol(&ws.rec[16 + c]), Y0 = ld_vol(&stg[c]), E0 = ld_vol(&stg[8 + c]);
  const uint32_t m = (uint32_t)S.mxs[x] + (uint32_t)S.mys[y];
  const uint32_t XV = X + X0, YV = Y + Y0, E = XV + YV;
  uint32_t t = __vimin3_s16x2(X0 + ld_vol(&S.Cw[3 * a]), Y0 + ld_vol(&S.Cw[3 * a + 1]), X0 + Y0 + ld_vol(&S.Cw[3 * a + 2]));
  t = __vimin3_s16x2(t, X + ld_vol(&S.Cw[3 * b]), Y + ld_vol(&S.Cw[3 * b + 1]));
  t = __vimin3_s16x2(t, X + Y + ld_vol(&S.Cw[3 * b + 2]), XV + ld_vol(&S.Cw[12 + 3 * bb]));
  t = __vimin3_s16x2(t, YV + ld_vol(&S.Cw[13 + 3 * bb]), E + ld_vol(&S.Cw[14 + 3 * bb]));
  int s16 = 0;
  if (b == 3) {
    const uint32_t top = (E0 & 0xffffu) + (E0 >> 16), bot = (E & 0xffffu) + (E >> 16);
    t = __vimin3_s16x2(t, ((bot << 16) | top) + ld_vol(&S.Cw[18]), E0 + E + ld_vol(&S.Cw[19]));
    s16 = (int)(top + bot) + *reinterpret_cast<const volatile int *>(&S.C16) + (int)m;
  }
  const bool ps = v && ((((t + m * 0x10001u) & 0x80008000u) != 0u) || s16 < 0);
  return __ballot_sync(0xffffffffu, ps) & 0xffu;
}

__device__ __forceinline__ uint32_t min2s(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// per-halfword minimum of the lane's eight candidates (3 VIMNMX3 + 1 VIMNMX)
__device__ __forceinline__ uint32_t gmin8(const uint32_t (&v)[2][4])
{
  const uint32_t a = __vimin3_s16x2(v[0][0], v[0][1], v[0][2]);
  const uint32_t b = __vimin3_s16x2(v[0][3], v[1][0], v[1][1]);
  const uint32_t c = __vimin3_s16x2(v[1][2], v[1][3], a);
  return min2s(b, c);
}

// The lane's packed sums go to the cold path through local memory: one rolled copy of the per-candidate test, and a small
// call site (the spill is a handful of STL.128).
template <class SLOT>
__device__ __forceinline__ uint32_t fs_cold_spill(SLOT &S, FsWarp &ws, const uint32_t (&X)[2][4], const uint32_t (&Y)[2][4], const uint32_t (&X0)[2][4],
                                                  const uint32_t (&Y0)[2][4], const uint32_t (&E0)[2][4], int b, int dxa, int dy0, uint32_t vm, uint32_t lanebits, FsTaskCtx tc)
{
  uint32_t buf[5][8];
#pragma unroll
  for (int q = 0; q < 2; q++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      buf[0][q * 4 + j] = X[q][j]; buf[1][q * 4 + j] = Y[q][j];
      buf[2][q * 4 + j] = X0[q][j]; buf[3][q * 4 + j] = Y0[q][j]; buf[4][q * 4 + j] = E0[q][j];
    }
  }
  return fs_cold_lane(S, ws, &buf[0][0], b, dxa, dy0, vm, lanebits, tc);
}

// One lane-job: candidates (dxa, dy0..dy0+3) and (dxa+4, same rows); wb = lane address of window row dy0.  The two 8-row
// halves of the macroblock run as a rolled loop (the register window slot of reference row i+j is (i+j)&3 in both halves, so
// the body is the same code; 12 desynchronised warps per SM stream through it).
//
// Filter (round 2).  Measured: the task loop is bound by the ALU pipe (VABSDIFF4 occupies it for two cycles; so do VIMNMX3,
// LOP3, SHF), not by issue slots, so what counts is the ALU work beside the 512 VABSDIFF4 of a task.  The sixteen 4x4 SADs of
// a candidate are packed two per register on the FMA pipe (IMAD), and the ONLY per-candidate ALU work left is the minimum of
// each packed 4x4 SAD over the lane's eight candidates (7 two-input minima per 8 candidates and register).  The 41-partition
// tree is then formed ONCE per lane on those minima:  sum over the 4x4 blocks k of P of min_c sad_k(c)  <=  min_c sad_P(c),
// so   LB_P - B_P - 1 + min_c m(c) < 0   is a necessary condition for any of the eight candidates to beat partition P's best
// (B_P = best_P >> 5, m = lower bound of the mv cost >> 5).  Two warp votes per task (after 8 and 16 rows) decide whether some
// lane passed; only then do the passing lanes test their candidates one by one (fs_cold_lane: exact sums, still in
// registers) and record the (candidate, partition, SAD) triples that pass.  Returns the pass bits of the lane's candidates
// (bit q*4+j), used only when the record list overflows.
//
// QN, JN: the candidates the routine computes -- columns q < QN of the pair, rows j < JN of the four.  <2, 4> is the full job;
// <1, 4> serves the trailing column block when it is at most 4 columns wide (the second column of every pair lies outside the
// window: the +-R window has an odd number of columns) and <2, 1> the last row group when only its first row exists (an odd
// number of rows): half / a quarter of the VABSDIFF4 of a full job.  The packed sums of the other candidates are a constant that
// never wins a minimum; their vm bits are clear, so the cold paths never look at them.
template <int PITCH, int QN, int JN, class SLOT>
__device__ __forceinline__ uint32_t fs_task4(SLOT &S, FsWarp &ws, uint32_t *stg, const uint8_t *wb, int dxa, int dy0, uint32_t vm, uint32_t mmin, uint32_t one, FsTaskCtx tc, long long &ncold)
{
  constexpr int K = 4;
  constexpr uint32_t BIG = 0x1fff1fffu;            // packed sums of a candidate that is not computed
  const uint4 *cur = reinterpret_cast<const uint4 *>(S.cur);
  const uint32_t lanebits = (uint32_t)(threadIdx.x & 31) << 25;
  const uint32_t mm2 = mmin * 0x10001u;
  uint32_t rw[K][6];
  uint32_t acc[2][K][4];
  uint32_t X0[2][K], Y0[2][K], E0[2][K];
  uint32_t pass = 0, ME0 = 0, MX0 = 0, MY0 = 0;
#pragma unroll
  for (int j = 0; j < K; j++) { E0[0][j] = E0[1][j] = 0; X0[0][j] = X0[1][j] = 0; Y0[0][j] = Y0[1][j] = 0; }
#pragma unroll
  for (int j = 0; j < K - 1; j++) ld3(rw[j], wb + j * PITCH);
  // FS_ROLL4: the loop is rolled over the four block rows (even / odd block rows take a warp-uniform branch after their four
  // rows) instead of over the two 8-row halves: the body is ~4.5 KB of SASS instead of ~7 KB.  The instruction caches are small
  // (L0 ~6 KB per scheduler, L1.5 32 KB per SM) and the SM's warps walk the loop desynchronised.
  constexpr int RPI = FS_ROLL4 ? 4 : 8;           // rows per rolled iteration
#pragma unroll 1
  for (int it = 0; it < 16 / RPI; it++) {
    const int bb = FS_ROLL4 ? it >> 1 : it;
    const uint8_t *wr = wb + it * RPI * PITCH;
    const uint32_t *Cb = S.Cw + 6 * bb;
#pragma unroll
    for (int r = 0; r < RPI; r++) {
      ld3(rw[(r + K - 1) & 3], wr + (r + K - 1) * PITCH);
      const uint4 c = cur[it * RPI + r];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int sl = (r + j) & 3;
#pragma unroll
        for (int q = 0; q < 2; q++) {
          if (q >= QN || j >= JN) continue;
          if ((r & 3) == 0) {
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
      if (FS_ROLL4 ? (r == 3 && !(it & 1)) : r == 3) {   // even block row 2*bb: pack, minima over the lane's candidates
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            X0[q][j] = (q < QN && j < JN) ? acc[q][j][2] * 65536u + acc[q][j][0] : BIG;
            Y0[q][j] = (q < QN && j < JN) ? acc[q][j][3] * 65536u + acc[q][j][1] : BIG;
          }
        }
        MX0 = gmin8(X0); MY0 = gmin8(Y0);
      }
      if (FS_ROLL4 ? (r == 3 && (it & 1)) : r == 7) {    // odd block row 2*bb+1: the tree on the minima, one vote
        uint32_t X[2][K], Y[2][K];
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            X[q][j] = (q < QN && j < JN) ? acc[q][j][2] * 65536u + acc[q][j][0] : BIG;
            Y[q][j] = (q < QN && j < JN) ? acc[q][j][3] * 65536u + acc[q][j][1] : BIG;
          }
        }
        const uint32_t MX1 = gmin8(X), MY1 = gmin8(Y);
        const uint32_t cx0 = ld_vol(&Cb[0]), cy0 = ld_vol(&Cb[1]), ch0 = ld_vol(&Cb[2]);
        const uint32_t cx1 = ld_vol(&Cb[3]), cy1 = ld_vol(&Cb[4]), ch1 = ld_vol(&Cb[5]);
        const uint32_t cxv = ld_vol(&S.Cw[12 + 3 * bb]), cyv = ld_vol(&S.Cw[13 + 3 * bb]), ce = ld_vol(&S.Cw[14 + 3 * bb]);
        const uint32_t MXV = add2(MX0, MX1), MYV = add2(MY0, MY1), ME = add2(MXV, MYV);
        uint32_t rr = __vimin3_s16x2(MX0 * one + cx0, MY0 * one + cy0, add2(MX0, MY0) * one + ch0);
        rr = __vimin3_s16x2(rr, MX1 * one + cx1, MY1 * one + cy1);
        rr = __vimin3_s16x2(rr, add2(MX1, MY1) * one + ch1, MXV * one + cxv);
        rr = __vimin3_s16x2(rr, MYV * one + cyv, ME * one + ce);
        int s16 = 0;
        if (bb) {                                     // whole-MB partitions
          const uint32_t ctb = ld_vol(&S.Cw[18]), clr = ld_vol(&S.Cw[19]);
          const int c16 = *reinterpret_cast<const volatile int *>(&S.C16);
          const uint32_t top = __dp2a_lo(ME0, 0x0101u, 0u), bot = __dp2a_lo(ME, 0x0101u, 0u);
          rr = __vimin3_s16x2(rr, (bot * 65536u + top) * one + ctb, add2(ME0, ME) * one + clr);
          s16 = (int)(top + bot) + c16 + (int)mmin;
        } else {
          ME0 = ME;
#pragma unroll
          for (int j = 0; j < K; j++) {               // the upper half's two 8x8 sums per candidate: read by the cold path at b == 3
#pragma unroll
            for (int q = 0; q < 2; q++) E0[q][j] = (q < QN && j < JN) ? add2(add2(X0[q][j], Y0[q][j]), add2(X[q][j], Y[q][j])) : BIG;
          }
        }
        const bool lp = vm != 0u && ((((rr + mm2) & 0x80008000u) != 0u) || s16 < 0);
        uint32_t bal = __ballot_sync(0xffffffffu, lp);
        if (bal) {                                    // cold: a candidate of some lane may beat a partition's best
#ifdef FS_PROFILE
          const long long tc0_ = clock64();
#endif
          if (__popc(bal) <= FS_SPARSE) {             // the usual case: the passing lanes one at a time, eight candidates side by side
            const int lane = threadIdx.x & 31;
#pragma unroll 1
            while (bal) {
              const int L = __ffs(bal) - 1; bal &= bal - 1;
              if (lane == L) {
                uint4 *d = reinterpret_cast<uint4 *>(ws.rec), *e = reinterpret_cast<uint4 *>(stg);
                d[0] = make_uint4(X[0][0], X[0][1], X[0][2], X[0][3]); d[1] = make_uint4(X[1][0], X[1][1], X[1][2], X[1][3]);
                d[2] = make_uint4(Y[0][0], Y[0][1], Y[0][2], Y[0][3]); d[3] = make_uint4(Y[1][0], Y[1][1], Y[1][2], Y[1][3]);
                d[4] = make_uint4(X0[0][0], X0[0][1], X0[0][2], X0[0][3]); d[5] = make_uint4(X0[1][0], X0[1][1], X0[1][2], X0[1][3]);
                e[0] = make_uint4(Y0[0][0], Y0[0][1], Y0[0][2], Y0[0][3]); e[1] = make_uint4(Y0[1][0], Y0[1][1], Y0[1][2], Y0[1][3]);
                e[2] = make_uint4(E0[0][0], E0[0][1], E0[0][2], E0[0][3]); e[3] = make_uint4(E0[1][0], E0[1][1], E0[1][2], E0[1][3]);
              }
              __syncwarp();
              const uint32_t bits = fs_cold_test8(S, ws, stg, 2 * bb + 1, __shfl_sync(0xffffffffu, dxa, L), __shfl_sync(0xffffffffu, dy0, L),
                                                  __shfl_sync(0xffffffffu, vm, L), tc.R);
              if (lane == L) pass |= bits;
              __syncwarp();
            }
          } else {
            if (lp) pass |= fs_cold_spill(S, ws, X, Y, X0, Y0, E0, 2 * bb + 1, dxa, dy0, vm, lanebits, tc);
            __syncwarp();
          }
#ifdef FS_PROFILE
          ncold += (1ll << 40) + (clock64() - tc0_);           // entries in the high bits, warp cycles in the low 40
#endif
        }
      }
    }
  }
  return pass;
}

// Survivors of a task (out of line: rare): the candidates whose pass bit is set are evaluated exactly for all 41 partitions at once
// (fs_exact2: 16 lanes per candidate for the 4x4 SADs, one lane per (candidate, partition)).  Returns the number of exact
// evaluations (lane 0).
template <int PITCH, class SLOT>
__device__ __noinline__ int fs_survivors(SLOT &S, FsWarp &ws, const uint8_t *win, int copy_bytes, const uint32_t *pgt, uint32_t pass,
                                         int dxa, int dy0, int R, int g, int lambda_f)
{
  constexpr int K = FS_K;
  const int lane = threadIdx.x & 31;
  int nh = 0;
  for (int bb = 0; bb < 2 * K; bb++) {
    uint32_t m = __ballot_sync(0xffffffffu, (pass >> bb) & 1u);
    while (m) {
      const int l0 = __ffs(m) - 1; m &= m - 1;
      const bool v1 = m != 0;
      const int l1 = v1 ? __ffs(m) - 1 : l0; m &= m - 1;
      const int ddx = bb >= K ? 4 : 0, ddy = bb >= K ? bb - K : bb;
      const int ex0_ = __shfl_sync(0xffffffffu, dxa, l0) + ddx, ey0 = __shfl_sync(0xffffffffu, dy0, l0) + ddy;
      const int ex1_ = __shfl_sync(0xffffffffu, dxa, l1) + ddx, ey1 = __shfl_sync(0xffffffffu, dy0, l1) + ddy;
      fs_exact2<PITCH>(S, ws, win, copy_bytes, pgt, ex0_, ey0, ex1_, ey1, v1, R, g, lambda_f);
      if (lane == 0) nh += v1 ? 2 : 1;
    }
  }
  return nh;
}

// ---- mbarrier / TMA (PTX) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(void *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// waits (hardware-suspended, no issue slots) until the phase of parity `parity` completes or about `ns` pass
__device__ __forceinline__ bool mbar_try_wait_ns(void *bar, uint32_t parity, uint32_t ns)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(void *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, void *bar, int x, int y, int z)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// Ordering between the shared-memory accesses of the claim / completion protocol: CTA-scope acquire-release
// fence (cheaper than the sequentially-consistent __threadfence_block()).
#define FS_FENCE() asm volatile("fence.acq_rel.cta;" ::: "memory")

#ifdef FS_PROFILE
#define FS_CLOCK() clock64()
#else
#define FS_CLOCK() 0ll
#endif

struct FsCtaStats { int err, nhits, ngroups, nitems; unsigned long long cyc[9]; };   // cyc: task, exact, producer total, worker idle, worker total, producer busy (warp cycles)

// ---- producer-warp routines (called by ONE warp) -------------------------------------------------------------

// Results of a finished item (me_fullsearch.c:95-102: mv += spiral[best], return min_mcost).
template <class SLOT>
__device__ __forceinline__ void fs_write_results(const SLOT &S, const FsArgs &a, FsCtaStats *st, int *lastmv)
{
  const int lane = threadIdx.x & 31;
  if (S.item < 0 || S.ngroups <= 0) return;
  const size_t base = ((size_t)S.base_hi << 32) | S.base_lo;
  for (int p = lane; p < NPART; p += 32) {
    if (!((a.part_mask >> p) & 1ull)) continue;
    const unsigned long long key = S.best[p];
    const int pos = (int)(key & 0xfffffull);
    int sx, sy; spiral_xy(pos, &sx, &sy);
    a.mv_int[(base + p) * 2]     = (int16_t)(S.pcx[p] + 4 * sx);
    a.mv_int[(base + p) * 2 + 1] = (int16_t)(S.pcy[p] + 4 * sy);
    if (p == 0) lastmv[S.ref & 15] = ((S.pcx[p] + 4 * sx) & 0xffff) | ((S.pcy[p] + 4 * sy) << 16);   // seed of this CTA's next item of that reference
    long long cost = (long long)(key >> 20);
    if (pos == 0 && cost == (1ll << 42) && a.min_mcost > (1ll << 42)) cost = a.min_mcost;   // bound never beaten
    a.cost_int[base + p] = cost;
  }
  if (lane == 0) { atomicAdd(&st->ngroups, S.ngroups); atomicAdd(&st->nitems, 1); }
}

// Everything of centre group g that does not need the window: box, geometry, window origin, filter constants
// (from the partitions' current bounds), mv-cost tables.
template <int PITCH, class SLOT>
__device__ __forceinline__ void fs_setup_group(SLOT &S, int g, const FsArgs &a)
{
  constexpr int K = FS_K;
  const int lane = threadIdx.x & 31, R = a.R;
  int gx0 = 0x7fff, gx1 = -0x7fff, gy0 = 0x7fff, gy1 = -0x7fff, qx0 = 0x7fff, qx1 = -0x7fff, qy0 = 0x7fff, qy1 = -0x7fff;
  for (int p = lane; p < NPART; p += 32) {
    if (S.pgrp[p] != g) continue;
    const int cx = S.pcx[p] >> 2, cy = S.pcy[p] >> 2, px = S.ppx[p], py = S.ppy[p];
    gx0 = min(gx0, cx); gx1 = max(gx1, cx); gy0 = min(gy0, cy); gy1 = max(gy1, cy);
    qx0 = min(qx0, px); qx1 = max(qx1, px); qy0 = min(qy0, py); qy1 = max(qy1, py);
  }
  gx0 = __reduce_min_sync(0xffffffffu, gx0); gx1 = __reduce_max_sync(0xffffffffu, gx1);
  gy0 = __reduce_min_sync(0xffffffffu, gy0); gy1 = __reduce_max_sync(0xffffffffu, gy1);
  qx0 = __reduce_min_sync(0xffffffffu, qx0); qx1 = __reduce_max_sync(0xffffffffu, qx1);
  qy0 = __reduce_min_sync(0xffffffffu, qy0); qy1 = __reduce_max_sync(0xffffffffu, qy1);
  const int spanx = gx1 - gx0, spany = gy1 - gy0;
  const int ncx = 2 * R + 1 + spanx, ncy = 2 * R + 1 + spany;
  const int ngy = (ncy + K - 1) / K;
  const int ncb = (ncx + 63) >> 6, wlast = ncx - 64 * (ncb - 1);
  const int ncbA = wlast <= FS_BMAX ? ncb - 1 : ncb;
  int npb = 0;
  if (ncbA < ncb) npb = wlast <= 4 ? wlast : (wlast <= 8 ? 4 : wlast - 4);
  const int mbx = S.mb % a.mbw, mby = S.mb / a.mbw;
  const int x0 = mbx * 16 + gx0 - R + a.spad, y0 = mby * 16 + gy0 - R + a.spad;
  // ---- filter constants: partitions of other groups never pass ----
  if (lane < 20) S.Cw[lane] = 0;
  if (lane == 20) S.C16 = 0;
  __syncwarp();
  for (int p = lane; p < NPART; p += 32) {
    if (S.pgrp[p] != g) continue;
    S.pex[p] = (signed char)((S.pcx[p] >> 2) - gx0); S.pey[p] = (signed char)((S.pcy[p] >> 2) - gy0);
    set_threshold(S, p, S.best[p]);
  }
  // ---- lower bound of the mv cost per window column / row: every partition of the group sees the same
  //      displacement 4*(g0 + d - R); bits is monotone in |mv - pred|, so the distance to the predictors'
  //      [min,max] interval bounds every partition's term from below ----
  for (int i = lane; i < ncx + ncy + K + 4; i += 32) {
    const bool isy = i >= ncx + 4;
    const int d = isy ? i - ncx - 4 : i;
    const int mv = 4 * ((isy ? gy0 : gx0) + d - R);
    const int lo = isy ? qy0 : qx0, hi = isy ? qy1 : qx1;
    const int dist = max(0, max(lo - mv, mv - hi));
    const long long v = ((long long)a.lambda_f * mvbits(dist)) >> 5;
    (isy ? S.mys : S.mxs)[d] = (unsigned short)(v > FS_MCAP ? FS_MCAP : v);
  }
  {                                                // type-A tasks: row groups from the centre outwards, per column block
    const int gc = min(ngy - 1, (R + (spany >> 1)) / K), lo = gc, hi = ngy - 1 - gc, mn = min(lo, hi);
    for (int t = lane; t < ncbA * ngy && t < SLOT::NTT; t += 32) {
      const int cb = t / ngy, k = t - cb * ngy;
      const int gy = k <= 2 * mn ? ((k & 1) ? gc + ((k + 1) >> 1) : gc - (k >> 1)) : (lo > hi ? gc - (k - hi) : gc + (k - lo));
      S.ttab[t] = (unsigned short)((gy * K) | (cb << 8));
    }
  }
  if (lane == 0) {
    S.g = g; S.gx0 = gx0; S.gy0 = gy0; S.spanx = spanx; S.spany = spany;
    S.wx0 = x0; S.wy0 = y0;
    S.inside = (x0 >= 0 && y0 >= 0 && x0 + ncx + 15 <= a.Wq && y0 + ncy + 15 <= a.Hq && !(a.flags & 1)) ? 1 : 0;
    S.ncx = ncx; S.ncy = ncy; S.ngy = ngy; S.gc = min(ngy - 1, (R + (spany >> 1)) / K);
    S.ncbA = ncbA; S.ntaskA = ncbA * ngy; S.npb = npb; S.ntask = ncbA * ngy + ((a.flags & 8) ? 0 : (npb * ngy + 31) / 32);
  }
  __syncwarp();
}

// Initial bounds of a unit (exact evaluation of the first centre and of the seed candidate).  global: read the candidates from
// the search planes in global memory -- the producer does that while the unit is still STAGED (fs_prepare), so that attaching
// it to a window buffer later is only the window copy: the pre-pass is off the buffer's turn-around path.  Windows that leave
// the search plane (clamped staging) keep the pre-pass on the staged window (fs_activate).
template <int PITCH, class SLOT>
__device__ __forceinline__ void fs_prepass(SLOT &S, FsWarp &ws, const uint8_t *win, const uint32_t *pgt, const FsArgs &a, const int *lastmv, bool global)
{
  const FsGeom G = fs_geom(a.R);
  const int lane = threadIdx.x & 31, R = a.R;
  // ---- initial bounds: exact pre-pass over the centres' box, plus ONE seed candidate: the 16x16 vector this CTA's last finished
  //      item of the same reference ended on.  Any candidate may be evaluated first without changing the result (the argmin
  //      is exact); when the predictor is wrong but the motion is coherent the seed brings the bounds down before the first
  //      task instead of after the row group that holds the optimum (measured: 6.6 ms -> see DESIGN 4.1 for predictors off by
  //      +-8 pel) ----
  {
    // FS_PRE < 0 (default): ONE candidate instead of the centres' box -- the centre of the group's first partition.  With one
    // predictor per partition the box holds up to (1 + span)^2 candidates and their exact evaluation saturated the producer
    // warp (measured: per-partition predictors 1.84 ms against 1.23 ms with shared ones); the other centres are ordinary
    // candidates of the hot path.
    int xlo, ylo, pw, ph;
    if (FS_PRE >= 0) {
      xlo = max(0, R - FS_PRE); ylo = max(0, R - FS_PRE);
      pw = min(S.ncx - 1, R + S.spanx + FS_PRE) - xlo + 1; ph = min(S.ncy - 1, R + S.spany + FS_PRE) - ylo + 1;
    } else {
      int p0 = 0;
      while (p0 < NPART - 1 && S.pgrp[p0] != S.g) p0++;
      xlo = R + S.pex[p0]; ylo = R + S.pey[p0]; pw = 1; ph = 1;
    }
    int total = pw * ph;
    int sdx = -1, sdy = -1;
    {
      const int sm = lastmv[S.ref & 15];
      if (sm != 0x7fff7fff && !(a.flags & 4)) {
        const int x = ((int)(short)(sm & 0xffff) >> 2) - S.gx0 + R, y = ((sm >> 16) >> 2) - S.gy0 + R;
        const bool inbox = x >= xlo && x < xlo + pw && y >= ylo && y < ylo + ph;
        if (x >= 0 && x < S.ncx && y >= 0 && y < S.ncy && !inbox) { sdx = x; sdy = y; }
      }
    }
    if (lane == 0) { S.seedx = sdx; S.seedy = sdy; S.prex = xlo; S.prey = ylo; S.prew = pw; S.preh = ph; }
    if (sdx >= 0) total++;
    for (int i0 = 0; i0 < total; i0 += 2) {
      const int i1 = i0 + 1;
      const bool v1 = i1 < total;
      const bool s0 = sdx >= 0 && i0 == total - 1, s1 = sdx >= 0 && i1 == total - 1;
      const int dx0 = s0 ? sdx : xlo + i0 % pw, dy0 = s0 ? sdy : ylo + i0 / pw;
      const int dx1 = !v1 ? dx0 : (s1 ? sdx : xlo + i1 % pw), dy1 = !v1 ? dy0 : (s1 ? sdy : ylo + i1 / pw);
      fs_exact2<PITCH>(S, ws, win, G.copy_bytes, pgt, dx0, dy0, dx1, dy1, v1, R, S.g, a.lambda_f, global ? &a : nullptr);
    }
  }
  __syncwarp();
}

// Loads item `item` (handed out in order by a global counter: load balance) into slot S and prepares its first
// centre group; fetches the index of the following item into `item` while the loads are in flight.  Returns false
// when the items are exhausted.  Touches no window: runs while the workers are busy.
template <int PITCH, class SLOT>
__device__ __noinline__ bool fs_prepare(SLOT &S, int &item_io, const FsArgs &a, FsCtaStats *st, FsWarp &ws, const uint32_t *pgt, const int *lastmv)
{
  const int lane = threadIdx.x & 31, R = a.R;
  for (;;) {
    const int item = item_io;
    if (item >= a.nitems) { if (lane == 0) { S.item = -1; S.ngroups = 0; } __syncwarp(); return false; }
    int nxt = 0;
    if (lane == 0) nxt = atomicAdd(a.work_counter, 1);
    // ---- per-partition parameters, current MB: every load is issued before the first use ----
    const int mb = a.mb_first + item / a.refs_per_mb, ref = a.ref_first + item % a.refs_per_mb;
    const int mbx = mb % a.mbw, mby = mb / a.mbw;
    const size_t base = (a.abs_index ? ((size_t)mb * a.nrefs + ref) : (size_t)item) * NPART;
    const uint32_t *cen32 = reinterpret_cast<const uint32_t *>(a.center) + base, *prd32 = reinterpret_cast<const uint32_t *>(a.pred) + base;
    const bool has1 = lane + 32 < NPART;
    const uint32_t c0 = cen32[lane], q0 = prd32[lane], c1 = has1 ? cen32[lane + 32] : 0u, q1 = has1 ? prd32[lane + 32] : 0u;
    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + (lane >> 2)) * a.cur_pitch + mbx * 16 + (lane & 3) * 4);
    const uint32_t w1 = *reinterpret_cast<const uint32_t *>(a.cur + (size_t)(mby * 16 + 8 + (lane >> 2)) * a.cur_pitch + mbx * 16 + (lane & 3) * 4);
    const long long mm = a.min_mcost < 0 ? 0 : (a.min_mcost > (1ll << 42) ? (1ll << 42) : a.min_mcost);
    int bx0 = 0x7fff, bx1 = -0x7fff, by0 = 0x7fff, by1 = -0x7fff;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int p = lane + 32 * h;
      if (p >= NPART) continue;
      const uint32_t cw = h ? c1 : c0, qw = h ? q1 : q0;
      const int cx = (short)(cw & 0xffffu), cy = (short)(cw >> 16);
      const bool act = (a.part_mask >> p) & 1ull;
      const PartGeom gm = part_geom(p);
      S.pcx[p] = (short)cx; S.pcy[p] = (short)cy;
      S.ppx[p] = (short)(qw & 0xffffu); S.ppy[p] = (short)(qw >> 16);
      S.psr[p] = (short)(a.restrict_mode < 0 ? a.sr_override : block_search_range(R, a.restrict_mode, ref, gm.bt));
      S.pgrp[p] = act ? 0 : -1;
      S.best[p] = ((unsigned long long)mm << 20);
      if (act) {
        if ((cx | cy) & 3) st->err = 1;            // sub-pel centres are not a full-search input
        bx0 = min(bx0, cx >> 2); bx1 = max(bx1, cx >> 2); by0 = min(by0, cy >> 2); by1 = max(by1, cy >> 2);
      }
    }
    S.cur[lane] = w0; S.cur[lane + 32] = w1;
    item_io = __shfl_sync(0xffffffffu, nxt, 0);
    bx0 = __reduce_min_sync(0xffffffffu, bx0); bx1 = __reduce_max_sync(0xffffffffu, bx1);
    by0 = __reduce_min_sync(0xffffffffu, by0); by1 = __reduce_max_sync(0xffffffffu, by1);
    int ng = bx1 >= bx0 ? 1 : 0;
    __syncwarp();
    // ---- cluster partitions whose centres fit in one FS_SMAX box (usually all of them) ----
    if (ng && (bx1 - bx0 > FS_SMAX || by1 - by0 > FS_SMAX)) {      // general case: greedy clustering (serial, rare)
      if (lane == 0) {
        short *gx0 = reinterpret_cast<short *>(S.mxs), *gx1 = gx0 + NPART, *gy0 = gx1 + NPART, *gy1 = gy0 + NPART;   // scratch
        ng = 0;
        for (int p = 0; p < NPART; p++) {
          if (S.pgrp[p] < 0) continue;
          const int cx = S.pcx[p] >> 2, cy = S.pcy[p] >> 2;
          int gg = -1;
          for (int q = 0; q < ng; q++) {
            const int nx0 = min((int)gx0[q], cx), nx1 = max((int)gx1[q], cx), ny0 = min((int)gy0[q], cy), ny1 = max((int)gy1[q], cy);
            if (nx1 - nx0 <= FS_SMAX && ny1 - ny0 <= FS_SMAX) { gg = q; gx0[q] = nx0; gx1[q] = nx1; gy0[q] = ny0; gy1[q] = ny1; break; }
          }
          if (gg < 0) { gg = ng++; gx0[gg] = gx1[gg] = cx; gy0[gg] = gy1[gg] = cy; }
          S.pgrp[p] = (signed char)gg;
        }
      }
      ng = __shfl_sync(0xffffffffu, ng, 0);
    }
    if (lane == 0) {
      S.item = item; S.mb = mb; S.ref = ref; S.base_lo = (unsigned)base; S.base_hi = (unsigned)((unsigned long long)base >> 32);
      S.ngroups = ng; S.g = 0;
    }
    __syncwarp();
    if (ng == 0) continue;                         // no active partition: nothing to search, nothing to write
    fs_setup_group<PITCH>(S, 0, a);
    if (S.inside) {                                // initial bounds now, from global memory: attaching the unit later is only the window copy
      fs_prepass<PITCH>(S, ws, nullptr, pgt, a, lastmv, true);
      if (lane == 0) S.inside = 2;
      __syncwarp();
    }
    return true;
  }
}

// Attaches the prepared unit S to window buffer C: window (TMA, or clamped staging at the picture's far
// outside), exact pre-pass over the centres' box (initial bounds), then publishes the unit: ready_epoch first,
// task counter second (see the worker loop).
template <int PITCH, class SLOT>
__device__ __noinline__ void fs_activate(FsCtl &C, SLOT &S, int slot_index, FsWarp &ws, uint8_t *win, const uint32_t *pgt,
                                         const FsArgs &a, const CUtensorMap *tm, FsCtaStats *st, int *lastmv, const SLOT *O = nullptr)
{
  constexpr int K = FS_K;
  const FsGeom G = fs_geom(a.R);
  const int lane = threadIdx.x & 31, R = a.R;
  const int ep = C.epoch + 1;
  const int x0 = S.wx0, y0 = S.wy0;
  if (S.inside) {                                  // ---- window: four TMA boxes (copy c starts c bytes right, c rows up) ----
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&C.mbar, 4u * PITCH * G.rows);
#pragma unroll
      for (int c = 0; c < 4; c++)      // a TMA box must start on a 16-byte boundary: take plane (x0+c)&15 at the aligned column
        tma_load_3d(win + c * G.copy_bytes, tm, &C.mbar, (x0 + c) & ~15, y0 - c, S.ref * 16 + ((x0 + c) & 15));
    }
    __syncwarp();
    if (O) { fs_write_results(*O, a, st, lastmv); __syncwarp(); }     // the finished unit's results, under the window copy
    const uint32_t parity = (uint32_t)C.tma_uses & 1u;
    while (!mbar_try_wait_ns(&C.mbar, parity, 2000u)) { if (a.flags & 2) __nanosleep(400); }
    __syncwarp();
    if (lane == 0) C.tma_uses++;
  } else {                                         // window leaves the search plane: per-pixel coordinate clamp
    if (O) { fs_write_results(*O, a, st, lastmv); __syncwarp(); }
    const uint8_t *plane = a.spl + (size_t)S.ref * 16 * a.Wq * a.Hq;     // shift-0 plane
    const int nrows = min(K * S.ngy + 15, G.rows - 3), wpw = PITCH >> 2;
    for (int r = 0; r < nrows; r++) {
      const uint8_t *grow = plane + (size_t)iclamp(y0 + r, 0, a.Hq - 1) * a.Wq;
      for (int j = lane; j < wpw; j += 32) {
        uint32_t b[7];
#pragma unroll
        for (int k = 0; k < 7; k++) b[k] = grow[iclamp(x0 + 4 * j + k, 0, a.Wq - 1)];
#pragma unroll
        for (int c = 0; c < 4; c++)
          reinterpret_cast<uint32_t *>(win + c * G.copy_bytes + (r + c) * PITCH)[j] = b[c] | (b[c + 1] << 8) | (b[c + 2] << 16) | (b[c + 3] << 24);
      }
    }
  }
  __syncwarp();
  if (S.inside != 2) fs_prepass<PITCH>(S, ws, win, pgt, a, lastmv, false);
  __syncwarp();
  if (lane == 0) {
    // seqlock: the negative epoch marks "being re-armed" BEFORE any field changes, so that a worker that still holds
    // a stale claim on the previous epoch's (exhausted) counter cannot pair it with the new unit's ntask / slot
    *reinterpret_cast<volatile int *>(&C.ready_epoch) = -ep;
    __threadfence_block();
    C.slot = slot_index; C.ntask = S.ntask; C.done = 0; C.epoch = ep;
    __threadfence_block();
    *reinterpret_cast<volatile int *>(&C.ready_epoch) = ep;
    __threadfence_block();
    atomicExch(&C.next, (ep & 0x7ff) << 20);
  }
  __syncwarp();
}

template <int PITCH, int NWORK, int MINB>
#ifdef FS_MAXNREG
__global__ void __maxnreg__(FS_MAXNREG) k_sad_fs(
#else
__global__ void __launch_bounds__((NWORK + 1) * 32, MINB) k_sad_fs(
#endif
    const CUtensorMap *__restrict__ tmap, const __grid_constant__ FsArgs a)
{
  using SLOT = FsSlotT<PITCH>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ SLOT SS[3];                           // two attached to the window buffers, one being prepared
  __shared__ FsCtl CB[2];
  __shared__ unsigned long long evt;               // completes a phase whenever a worker finishes a unit: the producer sleeps on it
  constexpr int NT = (NWORK + 1) * 32;
  __shared__ FsWarp WS[NWORK + 1];
  __shared__ __align__(16) uint32_t STG[NWORK][16];   // words 24..39 of the sparse cold path's staging (Y0, E0), per worker warp
  __shared__ uint32_t pgt[NPART];
  __shared__ FsCtaStats st;
  __shared__ int lastmv[16];                       // per reference slot: the 16x16 vector of the CTA's last finished item (0x7fff7fff: none)
  constexpr int K = FS_K;
  const FsGeom G = fs_geom(a.R);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = a.R;

  if (tid < NPART) {
    const PartGeom gm = part_geom(tid);
    pgt[tid] = (gm.ox >> 2) | ((gm.oy >> 2) << 4) | (((gm.ox + gm.w) >> 2) << 8) | (((gm.oy + gm.h) >> 2) << 12);
  }
  for (int i = tid; i < (int)((NWORK + 1) * sizeof(FsWarp) / 4); i += NT) reinterpret_cast<uint32_t *>(&WS[0])[i] = 0;
  if (tid == 0) { st.err = 0; st.nhits = 0; st.ngroups = 0; st.nitems = 0; for (int i = 0; i < 9; i++) st.cyc[i] = 0; }
  if (tid < 2) {
    FsCtl &C = CB[tid];
    C.ready_epoch = 0; C.finished_epoch = 0; C.next = 0; C.done = 0; C.ended = 0; C.epoch = 0; C.tma_uses = 0; C.ntask = 0; C.slot = tid;
    mbar_init(&C.mbar, 1);
    if (tid == 0) mbar_init(&evt, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 3) { SS[tid].item = -1; SS[tid].ngroups = 0; SS[tid].g = 0; }
  if (tid < 16) lastmv[tid] = 0x7fff7fff;
  __syncthreads();

  if (warp == NWORK) {
    // ---- producer warp: keeps one unit prepared ahead; when the workers complete a buffer's unit it attaches
    //      the prepared unit to that buffer, then writes the finished item's results and prepares the next ----
    const long long t_p0 = clock64();
    int seen0 = 0, seen1 = 0, stage = 2, item = 0;
    long long p_busy = 0;                            // cycles spent attaching / writing / preparing (FS_PROFILE)
    if (lane == 0) item = atomicAdd(a.work_counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    bool staged = fs_prepare<PITCH>(SS[0], item, a, &st, WS[warp], pgt, lastmv);
    for (int b = 0; b < 2; b++) {
      if (staged) {
        fs_activate<PITCH>(CB[b], SS[b], b, WS[warp], smem + b * G.slot_bytes, pgt, a, tmap, &st, lastmv);
        staged = fs_prepare<PITCH>(SS[b == 0 ? 1 : 2], item, a, &st, WS[warp], pgt, lastmv);
      } else if (lane == 0) { __threadfence_block(); *reinterpret_cast<volatile int *>(&CB[b].ended) = 1; }
    }
    for (;;) {
      int nend = 0; bool any = false;
      // parity of evt's current phase, read BEFORE the scan: a completion after this read ends the wait below at once
      const uint32_t evp = mbar_test_wait(&evt, 0) ? 1u : 0u;
      const long long tb0 = FS_CLOCK();
#pragma unroll 1
      for (int b = 0; b < 2; b++) {
        FsCtl &C = CB[b];
        if (*reinterpret_cast<volatile int *>(&C.ended)) { nend++; continue; }
        const int fin = *reinterpret_cast<volatile int *>(&C.finished_epoch);
        if (fin == (b ? seen1 : seen0)) continue;
        __threadfence_block();
        if (b) seen1 = fin; else seen0 = fin;
        any = true;
        const int old = C.slot;
        SLOT &O = SS[old];
        if (O.g + 1 < O.ngroups) {                   // next centre group of the same item, in place (rare)
          fs_setup_group<PITCH>(O, O.g + 1, a);
          if (O.inside) { fs_prepass<PITCH>(O, WS[warp], nullptr, pgt, a, lastmv, true); if (lane == 0) O.inside = 2; __syncwarp(); }
          fs_activate<PITCH>(C, O, old, WS[warp], smem + b * G.slot_bytes, pgt, a, tmap, &st, lastmv);
        } else if (staged) {
          fs_activate<PITCH>(C, SS[stage], stage, WS[warp], smem + b * G.slot_bytes, pgt, a, tmap, &st, lastmv, &O);
          stage = old;
          staged = fs_prepare<PITCH>(SS[stage], item, a, &st, WS[warp], pgt, lastmv);
        } else {
          fs_write_results(O, a, &st, lastmv);
          __syncwarp();
          if (lane == 0) { O.item = -1; __threadfence_block(); *reinterpret_cast<volatile int *>(&C.ended) = 1; }
        }
      }
      if (any) p_busy += FS_CLOCK() - tb0;
      if (nend == 2) break;
      if (!any) {
        // sleep on the event barrier; the hardware ends a try_wait after ~0.4 us whatever the hint says, so only every 16th
        // time-out goes back to the scan (two completions inside one scan flip the parity twice: they are seen at that rescan)
#pragma unroll 1
        for (int k = 0; k < 16 && !mbar_try_wait_ns(&evt, evp, 20000u); k++) { if (a.flags & 2) __nanosleep(1000); }
      }
    }
    if (lane == 0) { atomicAdd(&st.cyc[2], (unsigned long long)(clock64() - t_p0)); atomicAdd(&st.cyc[5], (unsigned long long)p_busy); }
  } else {
  // ---- worker warps.  They drain ONE buffer at a time (pref) and move to the other only when pref has no task
  //      left to claim, so the two units finish staggered and the producer's work overlaps the other buffer ----
  int ex0 = 0, ex1 = 0, nh = 0, pref = 0;
  long long ncold = 0;
  long long c_task = 0, c_exact = 0;
  const long long t_begin = clock64();
  unsigned long long g_begin; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_begin));
  long long c_idle = 0, c_claim = 0, c_dec = 0, c_post = 0;
  for (;;) {
    bool any = false; int nend = 0;
    const long long ti0 = FS_CLOCK();
#pragma unroll 1
    for (int k = 0; k < 2 && !any; k++) {
      const int b = pref ^ k;
      FsCtl &C = CB[b];
      const int ep0 = *reinterpret_cast<volatile int *>(&C.ready_epoch);
      if (ep0 == (b ? ex1 : ex0)) { if (*reinterpret_cast<volatile int *>(&C.ended)) nend++; continue; }
      int c = 0;
      if (lane == 0) c = atomicAdd(&C.next, a.claim);
      c = __shfl_sync(0xffffffffu, c, 0);
      const int ep = *reinterpret_cast<volatile int *>(&C.ready_epoch);
      if (ep <= 0 || (c >> 20) != (ep & 0x7ff)) continue;       // counter not (yet) armed for the published epoch
      FS_FENCE();
      const int ntask = *reinterpret_cast<volatile int *>(&C.ntask);
      const int slot = *reinterpret_cast<volatile int *>(&C.slot);
      FS_FENCE();
      if (*reinterpret_cast<volatile int *>(&C.ready_epoch) != ep) continue;   // re-armed meanwhile: ntask / slot may be the next unit's
      if ((c & 0xfffff) >= ntask) { if (b) ex1 = ep; else ex0 = ep; continue; }
      pref = b;
      any = true;
      SLOT &S = SS[slot];
      uint8_t *win = smem + b * G.slot_bytes;
      const int g = S.g;
      int t0 = c & 0xfffff;
      c_claim += FS_CLOCK() - ti0;
      // A warp that holds unfinished tasks of (b, ep) knows the unit cannot complete, so the buffer cannot be re-armed:
      // its NEXT claim is issued before the tasks (the shared-memory atomic's latency hides under them) and needs no
      // validation, and its completion count is a fire-and-forget add unless that next claim came back exhausted.
#pragma unroll 1
      for (;;) {
      const int t1 = min(t0 + a.claim, ntask);
      const bool more = t1 < ntask;
      int cn = 0;
      if (more && lane == 0) cn = atomicAdd(&C.next, a.claim);
#pragma unroll 1
      for (int t = t0; t < t1; t++) {
        const long long td0 = FS_CLOCK();
        int dxa, dy0; bool va, vb;
        if (t < S.ntaskA) {
          const int e = S.ttab[t];
          dxa = 64 * (e >> 8) + 8 * (lane >> 2) + (lane & 3); dy0 = e & 255;
          va = dxa < S.ncx; vb = dxa + 4 < S.ncx;
        } else {
          const int job = (t - S.ntaskA) * 32 + lane;
          const int i = job / S.ngy, gy = job - i * S.ngy;
          dxa = 64 * S.ncbA + 8 * (i >> 2) + (i & 3); dy0 = gy * K;
          va = i < S.npb && dxa < S.ncx; vb = va && dxa + 4 < S.ncx;
          if (!va) dy0 = 0;
        }
        if (!va) dxa = 0;
        const int cc = dxa & 3;
        const uint8_t *wb = win + cc * G.copy_bytes + (dy0 + cc) * PITCH + (dxa >> 2) * 4;
        const long long tt0 = FS_CLOCK();
        c_dec += tt0 - td0;
        uint32_t vm = 0;
#pragma unroll
        for (int j = 0; j < K; j++) if (dy0 + j < S.ncy) vm |= (va ? 1u << j : 0u) | (vb ? 1u << (K + j) : 0u);
        // lower bound of the mv cost over the lane's eight candidates (columns dxa, dxa+4; rows dy0..dy0+3)
        const uint32_t mmin = min((uint32_t)S.mxs[dxa], (uint32_t)S.mxs[dxa + 4]) +
                              min(min((uint32_t)S.mys[dy0], (uint32_t)S.mys[dy0 + 1]), min((uint32_t)S.mys[dy0 + 2], (uint32_t)S.mys[dy0 + 3]));
        uint32_t pass = 0;
#pragma unroll 1
        for (int rep = 0; rep <= (a.flags >> 8); rep++)      // flags >> 8: extra repetitions of the task (throughput probe, B2ME_FS_REP)
          if (FS_REDUCED && t >= S.ntaskA && S.npb <= 4)         // trailing column block of <= 4 columns: first column of the pairs only
            pass |= fs_task4<PITCH, 1, 4>(S, WS[warp], STG[warp], wb, dxa, dy0, vm, mmin, (uint32_t)a.one, FsTaskCtx{t, R, g, a.lambda_f}, ncold);
          else if (FS_REDUCED && t < S.ntaskA && dy0 + 1 >= S.ncy)  // last row group, one row
            pass |= fs_task4<PITCH, 2, 1>(S, WS[warp], STG[warp], wb, dxa, dy0, vm, mmin, (uint32_t)a.one, FsTaskCtx{t, R, g, a.lambda_f}, ncold);
          else
            pass |= fs_task4<PITCH, 2, 4>(S, WS[warp], STG[warp], wb, dxa, dy0, vm, mmin, (uint32_t)a.one, FsTaskCtx{t, R, g, a.lambda_f}, ncold);
        const long long tt1 = FS_CLOCK();
        c_task += tt1 - tt0;
        if (__any_sync(0xffffffffu, pass != 0u)) {                      // survivors (rare): out of line
          nh += fs_survivors<PITCH>(S, WS[warp], win, G.copy_bytes, pgt, pass, dxa, dy0, R, g, a.lambda_f);
          c_exact += FS_CLOCK() - tt1;
        }
      }
      const long long tp0 = FS_CLOCK();
      __syncwarp();
      cn = __shfl_sync(0xffffffffu, cn, 0);
      cn = more ? (cn & 0xfffff) : ntask;
      if (cn < ntask) {                              // more tasks of this unit in hand
        if (lane == 0) { FS_FENCE(); atomicAdd(&C.done, t1 - t0); }
        t0 = cn;
        c_post += FS_CLOCK() - tp0;
        continue;
      }
      int d = 0;
      if (lane == 0) { FS_FENCE(); d = atomicAdd(&C.done, t1 - t0) + (t1 - t0); }
      d = __shfl_sync(0xffffffffu, d, 0);
      if (b) ex1 = ep; else ex0 = ep;                // nothing left to claim in this epoch
      if (d == ntask) {                              // unit complete: hand the buffer to the producer warp
        if (lane == 0) { FS_FENCE(); *reinterpret_cast<volatile int *>(&C.finished_epoch) = ep; FS_FENCE(); mbar_arrive(&evt); }
      }
      c_post += FS_CLOCK() - tp0;
      break;
      }
    }
    if (nend == 2) break;
    if (!any) { __nanosleep(200); c_idle += FS_CLOCK() - ti0; }
  }
  nh = __reduce_add_sync(0xffffffffu, nh);
  if (lane == 0 && nh) atomicAdd(&st.nhits, nh);
  if (lane == 0 && ncold && a.stats) { atomicAdd(&a.stats[14], (unsigned long long)(ncold >> 40)); atomicAdd(&a.stats[15], (unsigned long long)(ncold & ((1ll << 40) - 1))); }
  if (lane == 0) { atomicAdd(&st.cyc[3], (unsigned long long)c_idle); atomicAdd(&st.cyc[6], (unsigned long long)c_claim); atomicAdd(&st.cyc[7], (unsigned long long)c_dec); atomicAdd(&st.cyc[8], (unsigned long long)c_post); }
  if (lane == 0) {
    atomicAdd(&st.cyc[0], (unsigned long long)c_task); atomicAdd(&st.cyc[1], (unsigned long long)c_exact);
    atomicAdd(&st.cyc[4], (unsigned long long)(clock64() - t_begin));
    if (blockIdx.x == 0 && warp == 0 && a.stats) {
      unsigned long long g_end; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_end));
      atomicAdd(&a.stats[9], (unsigned long long)(clock64() - t_begin)); atomicAdd(&a.stats[10], g_end - g_begin);
    }
  }
  }
  __syncthreads();
  if (tid == 0) {
    if (st.err) *a.errflag = 1;
    if (a.stats) {
      atomicAdd(&a.stats[0], (unsigned long long)st.nhits);
      atomicAdd(&a.stats[1], (unsigned long long)st.ngroups);
      atomicAdd(&a.stats[2], (unsigned long long)st.nitems);
      for (int i = 0; i < 6; i++) atomicAdd(&a.stats[3 + i], st.cyc[i]);
      for (int i = 6; i < 9; i++) atomicAdd(&a.stats[5 + i], st.cyc[i]);     // [9], [10]: CTA 0's clock calibration
    }
  }
}

FsGeom fs_geom_host(int R) { return fs_geom(R); }

cudaError_t launch_sad_fs(const FsArgs &a, const CUtensorMap *tm, int sm_count, cudaStream_t s, int *smem_bytes_out)
{
  const FsGeom G = fs_geom(a.R);
  static int configured160[64] = {0};                   // cudaFuncSetAttribute is per DEVICE: one flag per device ordinal
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  cudaError_t e;
  if (smem_bytes_out) *smem_bytes_out = G.total;
  if (G.pitch == 96) {
    // CTA shape: workers x resident CTAs per SM (B2ME_FS_VAR = "4x3" default; "3x3", "5x3", "6x3", "7x2", "6x2": development probes)
    static int var = -1;
    if (var < 0) { const char *e = getenv("B2ME_FS_VAR"); var = (e && e[0] && e[1] == 'x' && e[2]) ? (e[0] - '0') * 10 + (e[2] - '0') : 43; }
#define FS_LAUNCH96(NW, MB)                                                                                                   \
    {                                                                                                                         \
      static int configured[64] = {0}, occs[64] = {0};                                                                        \
      if (configured[dev] < G.total) {                                                                                        \
        e = cudaFuncSetAttribute(k_sad_fs<96, NW, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, G.total);                 \
        if (e != cudaSuccess) return e;                                                                                       \
        configured[dev] = G.total; occs[dev] = 0;                                                                             \
      }                                                                                                                       \
      if (occs[dev] < 1) {                                                                                                    \
        int occ = 0;                                                                                                          \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sad_fs<96, NW, MB>, (NW + 1) * 32, G.total);                    \
        if (getenv("B2ME_FS_PROFILE")) fprintf(stderr, "[b2me] k_sad_fs<96,%d>: %d CTAs/SM (dyn smem %d)\n", NW, occ, G.total); \
        occs[dev] = occ < 1 ? 1 : occ;                                                                                        \
      }                                                                                                                       \
      const int occ = occs[dev];                                                                                              \
      const int grid = min((a.nitems + 1) / 2, sm_count * occ);                                                               \
      k_sad_fs<96, NW, MB><<<grid, (NW + 1) * 32, G.total, s>>>(tm, a);                                                       \
    }
    if (var == 72) FS_LAUNCH96(7, 2)
    else if (var == 62) FS_LAUNCH96(6, 2)
    else if (var == 33) FS_LAUNCH96(3, 3)
    else if (var == 53) FS_LAUNCH96(5, 3)
    else if (var == 63) FS_LAUNCH96(6, 3)
    else FS_LAUNCH96(4, 3)
#undef FS_LAUNCH96
  } else if (G.pitch == 160) {
    if (configured160[dev] < G.total) {
      e = cudaFuncSetAttribute(k_sad_fs<160, 12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G.total);
      if (e != cudaSuccess) return e;
      configured160[dev] = G.total;
    }
    const int grid = min((a.nitems + 1) / 2, sm_count);
    k_sad_fs<160, 12, 1><<<grid, G.threads, G.total, s>>>(tm, a);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace b2
