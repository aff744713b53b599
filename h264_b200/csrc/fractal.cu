// fractal.cu -- the fractal range/domain block search of version1 (b2fr_* entry points).
//
// Replaces   compute_domain_Sum / compute_range_Sum   V1/src/compute.c:277-684 / 686-1091
//            compute_rdSum                            V1/src/compute.c:192-215
//            compute_rms                              V1/src/compute.c:6-189
//            full_search / bound_chk                  V1/src/block_enc.c:1933-1977 / 2894-2919
//
// Formulation.  version1 matches a range block against SAME-SIZE blocks of the previous
// reconstructed frame displaced by (i,j), |i|,|j| <= Search_Range (SURVEY Q-F1: no isometries, no
// decimation in the shipped code).  Every range block of every level lies on a fixed grid
// (Q-F10) and the 1+2+2+4+8+8+16 = 41 blocks of a macroblock share the displacement set, so
//   sum r*d, sum d, sum d^2  of a block at displacement (i,j)
// are sums of the same quantities of its 4x4 sub-blocks at (i,j): one CTA per (macroblock, plane
// set) computes the sixteen 4x4 cross terms per displacement once and tree-sums them into the 41
// partitions -- all exact integers.  The least-squares fit, the QUAN_A quantisation, the range
// test and the collage error are then evaluated in double precision with the reference's operand
// order and explicit round-to-nearest intrinsics (no FMA contraction), so alpha/beta/rms are the
// same IEEE-754 values x86-64 SSE2 produces.  argmin = strict '<' in the
// reference's ring-walk order (10.8), i.e. the lexicographic minimum of (rms, visit order).
#include <cstdlib>
#include <cstring>
#include <new>
#include "b2_common.cuh"
#include "../../include/b2me.h"

using namespace b2;

struct b2fr_ctx {
  int device, W, H, R;
  uint8_t *d_org[3];            // range planes Y,U,V
  uint8_t *d_ref[4][3];         // domain plane sets C,H,M,N x Y,U,V
  int *d_s4[4][3], *d_q4[4][3]; // 4x4 sliding sums / sums of squares of the domain planes  [h][w]
  int *d_rs4[3], *d_rq4[3];     // 4x4 grid sums of the range planes [h/4][w/4]
  // eager search results per (set, comp): [nmb][41]
  int *d_xy[4][3]; double *d_so[4][3]; double *d_rms[4][3];
  int *h_xy[4][3]; double *h_so[4][3]; double *h_rms[4][3];
  int valid[4][3];              // host copy of (set, comp) results is current
  int dvalid[4][3];             // device results of (set, comp) are current
  int loaded[4][3];             // the caller has uploaded (set, comp) at least once (else: zero plane, zero tables)
  b2fr_node *d_nodes[3];        // TRANS_NODE trees of the last b2fr_encode_plane / b2fr_decode_plane per component [nmb][21]
  int nodes_valid[3];
  uint8_t *d_rec;               // reconstructed plane scratch
  int *d_tab;                   // scratch for table read-back
  cudaStream_t stream;
  int64_t launches;
  char err[512];
};

static char g_frerr[512] = "no context";
#define FR_CHECK(ctx, expr) B2_CUDA_CHECK(ctx, expr)

static inline int comp_w(const b2fr_ctx *c, int comp) { return comp ? c->W / 2 : c->W; }
static inline int comp_h(const b2fr_ctx *c, int comp) { return comp ? c->H / 2 : c->H; }
// V1: mb grid of a chroma plane = (frmWidthInMbs/2) x (frmHeightInMbs/2)  (block_enc.c:514-515, image.c:1126-1133)
static inline int comp_mbw(const b2fr_ctx *c, int comp) { return comp ? (c->W / 16) / 2 : c->W / 16; }
static inline int comp_mbh(const b2fr_ctx *c, int comp) { return comp ? (c->H / 16) / 2 : c->H / 16; }

namespace b2 {

// ---- compute_domain_Sum: S4/Q4 at every pixel offset where a 4x4 block fits, else 0 ----------
template <int STRIP>
__global__ void __launch_bounds__(256) k_frac_domain_sums(const uint8_t *__restrict__ img, int w, int h,
                                                           int *__restrict__ s4, int *__restrict__ q4)
{
  // separable: each thread owns one column x of a STRIP-row strip; horizontal 4-sums of a row are
  // formed from the row's bytes, vertical 4-sums slide down the strip.  (STRIP 16 for planes that would not fill the
  // chip with 64-row strips: a CIF plane is 10 CTAs of 64 rows, latency-bound at 37 us.)
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * STRIP;
  if (x >= w) return;
  int hs[4] = {0, 0, 0, 0}, hq[4] = {0, 0, 0, 0};
  const bool xin = x + 3 < w;
  for (int r = 0; r < STRIP + 3; r++) {
    const int y = y0 + r;
    int a = 0, b = 0;
    if (y < h && xin) {
      const uint8_t *p = img + (size_t)y * w + x;
      const int p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
      a = p0 + p1 + p2 + p3; b = p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3;
    }
    hs[r & 3] = a; hq[r & 3] = b;
    if (r >= 3) {
      const int yo = y - 3;
      if (yo < h && yo < y0 + STRIP) {
        const bool ok = xin && y < h;
        s4[(size_t)yo * w + x] = ok ? hs[0] + hs[1] + hs[2] + hs[3] : 0;
        q4[(size_t)yo * w + x] = ok ? hq[0] + hq[1] + hq[2] + hq[3] : 0;
      }
    }
  }
}

// ---- compute_range_Sum: 4x4 sums on the block grid --------------------------------------------
__global__ void __launch_bounds__(256) k_frac_range_sums(const uint8_t *__restrict__ img, int w, int h,
                                                          int *__restrict__ rs4, int *__restrict__ rq4)
{
  const int gw = w >> 2, gh = h >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= gw * gh) return;
  const int gx = i % gw, gy = i / gw;
  int s = 0, q = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uchar4 v = *reinterpret_cast<const uchar4 *>(img + (size_t)(gy * 4 + r) * w + gx * 4);
    s += v.x + v.y + v.z + v.w;
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  rs4[i] = s; rq4[i] = q;
}

// table of block size (bw x bh) at every offset, built from S4 (parity read-back of compute_domain_Sum)
__global__ void __launch_bounds__(256) k_frac_table(const int *__restrict__ s4, int w, int h, int bw, int bh, int *__restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int x = i % w, y = i / w;
  int s = 0;
  if (x + bw <= w && y + bh <= h)
    for (int yy = 0; yy < bh; yy += 4)
      for (int xx = 0; xx < bw; xx += 4) s += s4[(size_t)(y + yy) * w + x + xx];
  out[i] = s;
}

// visit order of displacement (i,j) in full_search's ring walk (block_enc.c:1944-1974)
__host__ __device__ __forceinline__ int v1_order(int i, int j)
{
  const int ai = i < 0 ? -i : i, aj = j < 0 ? -j : j, l = ai > aj ? ai : aj;
  if (l == 0) return 0;
  const int base = 1 + 4 * l * (l - 1);
  if (j == -l && i < l) return base + (i + l);
  if (i == l && j < l) return base + 2 * l + (j + l);
  if (j == l && i > -l) return base + 4 * l + (l - i);
  return base + 6 * l + (l - j);
}

// QUAN_A (V1/inc/defines_enc.h:591-601) on an int: C '%' and '/' truncate toward zero
__device__ __forceinline__ int quan_a(int x)
{
  int b = x % 10, c = x / 10;
  if (b > 2 && b < 8) b = 5;
  else if (b > 7) { b = 0; c += 1; }
  else b = 0;
  return c * 10 + b;
}
// (int)double as x86-64 cvttsd2si does it: out-of-range and NaN give INT_MIN
__device__ __forceinline__ int x86_d2i(double v)
{
  if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
  return __double2int_rz(v);
}

// compute_rms (compute.c:156-188) from the five integer sums; returns rms, writes alpha, beta.
__device__ __forceinline__ double v1_rms(int n, int ids1, int ids2, int irs1, int irs2, long long ird, double *alpha_o, double *beta_o)
{
  const double no = (double)n, dsum1 = (double)ids1, dsum2 = (double)ids2, rsum1 = (double)irs1, rsum2 = (double)irs2, rdsum = (double)ird;
  const double det = __dsub_rn(__dmul_rn(no, dsum2), __dmul_rn(dsum1, dsum1));
  double alpha = 0.0;
  if (det != 0.0) alpha = __ddiv_rn(__dsub_rn(__dmul_rn(no, rdsum), __dmul_rn(rsum1, dsum1)), det);
  int a = x86_d2i(__dmul_rn(alpha, 100.0));
  // n is a power of two for every partition (16 .. 256): x / n == x * 2^-k bit for bit, and an FP64 division is ~30 instructions
  const bool p2 = (n & (n - 1)) == 0;
  const double inv_no = 1.0 / no;                 // folded at compile time where n is (the partition loop is unrolled); exact for a power of two
  double beta = p2 ? __dmul_rn(rsum1, inv_no) : __ddiv_rn(rsum1, no);
  a = quan_a(a);
  beta = (double)quan_a(x86_d2i(beta));
  alpha = __ddiv_rn((double)a, 100.0);
  *alpha_o = alpha; *beta_o = beta;
  if (alpha < -2.35 || alpha > 4.0) return 1e30;
  if (beta < -60.0 || beta > 255.0) return 1e30;
  const double ad = __dmul_rn(alpha, dsum1);
  const double t = __dsub_rn(beta, p2 ? __dmul_rn(ad, inv_no) : __ddiv_rn(ad, no));
  const double in1 = __dadd_rn(__dsub_rn(__dmul_rn(alpha, dsum2), __dmul_rn(2.0, rdsum)), __dmul_rn(__dmul_rn(2.0, t), dsum1));
  const double in2 = __dsub_rn(__dmul_rn(t, no), __dmul_rn(2.0, rsum1));
  return __dadd_rn(__dadd_rn(rsum2, __dmul_rn(alpha, in1)), __dmul_rn(t, in2));
}

struct FrArgs {
  const uint8_t *org, *ref;      // range / domain plane of this component
  const int *s4, *q4, *rs4, *rq4;
  int w, h, mbw, R;
  int *xy; double *so; double *rms;   // [nmb][41] results
};

constexpr int FR_NT = 256;
#ifndef FR_MINBV
#define FR_MINBV 2
#endif
constexpr int FR_MINB = FR_MINBV;      // resident CTAs per SM asked of ptxas (164 registers left one CTA of 8 warps per SM: 11 % issue)

// One CTA per macroblock: all displacements x 41 partitions.
__global__ void __launch_bounds__(FR_NT, FR_MINB) k_frac_window(const FrArgs a)
{
  extern __shared__ __align__(16) uint8_t sm[];
  const int R = a.R, ww = 16 + 2 * R;
  uint8_t *rng = sm;                       // [16][16]
  uint8_t *win = sm + 256;                 // [ww][ww], pixels outside the plane are 0 (never used by a valid candidate)
  __shared__ int rs[NPART], rq[NPART];     // range sums per partition
  __shared__ double wb_rms[FR_NT / 32][NPART], wb_al[FR_NT / 32][NPART], wb_be[FR_NT / 32][NPART];
  __shared__ int wb_ord[FR_NT / 32][NPART], wb_ij[FR_NT / 32][NPART];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mb = blockIdx.x, mbx = mb % a.mbw, mby = mb / a.mbw;
  const int bx0 = mbx * 16, by0 = mby * 16;

  rng[tid] = a.org[(size_t)(by0 + (tid >> 4)) * a.w + bx0 + (tid & 15)];
  for (int i = tid; i < ww * ww; i += FR_NT) {
    const int x = bx0 - R + i % ww, y = by0 - R + i / ww;
    win[i] = (x >= 0 && y >= 0 && x < a.w && y < a.h) ? a.ref[(size_t)y * a.w + x] : 0;
  }
  if (tid < NPART) {
    const PartGeom g = part_geom(tid);
    int s = 0, q = 0;
    const int gw = a.w >> 2;
    for (int yy = 0; yy < g.h; yy += 4)
      for (int xx = 0; xx < g.w; xx += 4) {
        const int gi = ((by0 + g.oy + yy) >> 2) * gw + ((bx0 + g.ox + xx) >> 2);
        s += a.rs4[gi]; q += a.rq4[gi];
      }
    rs[tid] = s; rq[tid] = q;
  }
  for (int i = tid; i < (FR_NT / 32) * NPART; i += FR_NT) { (&wb_rms[0][0])[i] = 2e30; (&wb_ord[0][0])[i] = 0x7fffffff; }
  __syncthreads();

  const int nd = (2 * R + 1) * (2 * R + 1);
  for (int d0 = 0; d0 < nd; d0 += FR_NT) {
    const int d = d0 + tid;
    const bool live = d < nd;
    const int di = live ? d % (2 * R + 1) - R : 0, dj = live ? d / (2 * R + 1) - R : 0;
    // sixteen 4x4 cross terms at this displacement
    int rd4[16], ds4[16], dq4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int sx = (k & 3) * 4, sy = (k >> 2) * 4;
      int acc = 0;
#pragma unroll
      for (int yy = 0; yy < 4; yy++) {
        const uint8_t *wr = win + (sy + yy + dj + R) * ww + sx + di + R;
        const uint8_t *rr = rng + (sy + yy) * 16 + sx;
#pragma unroll
        for (int xx = 0; xx < 4; xx++) acc += (int)rr[xx] * (int)wr[xx];
      }
      rd4[k] = acc;
      const int gx = bx0 + sx + di, gy = by0 + sy + dj;
      const bool in = live && gx >= 0 && gy >= 0 && gx + 4 <= a.w && gy + 4 <= a.h;
      ds4[k] = in ? a.s4[(size_t)gy * a.w + gx] : 0;
      dq4[k] = in ? a.q4[(size_t)gy * a.w + gx] : 0;
    }
    const int ord = v1_order(di, dj);
#pragma unroll
    for (int p = 0; p < NPART; p++) {
      const PartGeom g = part_geom(p);
      int rd = 0, ds = 0, dq = 0;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const int kx = (k & 3) * 4, ky = (k >> 2) * 4;
        if (kx >= g.ox && kx < g.ox + g.w && ky >= g.oy && ky < g.oy + g.h) { rd += rd4[k]; ds += ds4[k]; dq += dq4[k]; }
      }
      // bound_chk: the domain block lies inside the plane (the +-R test holds by construction);
      // the (0,0) start candidate is evaluated unconditionally (block_enc.c:1941)
      const int m = bx0 + g.ox + di, n = by0 + g.oy + dj;
      const bool ok = live && (ord == 0 || (m >= 0 && n >= 0 && m + g.w <= a.w && n + g.h <= a.h));
      double al = 0.0, be = 0.0, rms = 2e30;
      int o = 0x7fffffff;
      if (ok) { rms = v1_rms(g.w * g.h, ds, dq, rs[p], rq[p], rd, &al, &be); o = ord; }
      const int ij = (di & 0xffff) | (dj << 16);
      // lexicographic (rms, visit order) minimum over the warp: the minimum rms through shuffles, the visit order among the
      // lanes that hold it through one integer reduction (orders are distinct); the winning LANE then compares with the warp's
      // running best and writes its own scale / offset / displacement -- nothing but the rms ever crosses lanes
      double rmin = rms;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) { const double r2 = __shfl_xor_sync(0xffffffffu, rmin, s); rmin = r2 < rmin ? r2 : rmin; }
      const int omin = __reduce_min_sync(0xffffffffu, rms == rmin ? o : 0x7fffffff);
      if (rms == rmin && o == omin && o != 0x7fffffff && (rms < wb_rms[warp][p] || (rms == wb_rms[warp][p] && o < wb_ord[warp][p]))) {
        wb_rms[warp][p] = rms; wb_ord[warp][p] = o; wb_al[warp][p] = al; wb_be[warp][p] = be; wb_ij[warp][p] = ij;
      }
    }
  }
  __syncthreads();
  if (tid < NPART) {
    int bw = 0;
    for (int w2 = 1; w2 < FR_NT / 32; w2++)
      if (wb_rms[w2][tid] < wb_rms[bw][tid] || (wb_rms[w2][tid] == wb_rms[bw][tid] && wb_ord[w2][tid] < wb_ord[bw][tid])) bw = w2;
    // full_search semantics: the (0,0) candidate seeds best_rms/scale/offset even when it is
    // range-rejected (1e30); a later candidate replaces it only if strictly smaller.  (0,0) has
    // visit order 0, so the lexicographic minimum is exactly that rule.  x,y are written only when
    // a non-start candidate wins (Q-F11): the caller's zero stays otherwise.
    const size_t o = (size_t)mb * NPART + tid;
    const int ij = wb_ij[bw][tid];
    a.xy[2 * o] = (int)(short)(ij & 0xffff); a.xy[2 * o + 1] = ij >> 16;
    a.so[2 * o] = wb_al[bw][tid]; a.so[2 * o + 1] = wb_be[bw][tid];
    a.rms[o] = wb_rms[bw][tid];
  }
}

// ---- F5: the partition cascade of every macroblock ----------------------------------------------------------
// encode_one_macroblock / encode_block_rect / encode_block_8 / encode_block_4 (V1/src/block_enc.c:508-1051,
// 1072-1334, 1337-1675, 1676-1930; num_regions == 1, search_mode == 0, currentVideo == 'C').  Every block is
// searched on the four plane sets (k_frac_window) and the best kept with a strict '<' (C, H, M, N order); the cascade
// compares those rms values with tol^2 * n and, for the macroblock, the squared normalised cross-correlation `chun` of
// the range block with the CO-LOCATED block of set C (:811-848).  One thread per macroblock replays the reference's
// writes to its TRANS_NODE tree (fields survive between the 16x8 / 8x16 / 8x8 attempts on the same nodes; `reference`
// of a node is reset only by the H comparison; encode_block_4 also sets partition = 1 when H wins, :1773; after the
// 16x8 / 8x16 attempts a macroblock ALWAYS goes on to 8x8, :915, while an 8x8 block stops at 8x4 / 4x8, :1641).
struct FdArgs {
  const uint8_t *org, *refC; int w, mbw, nmb;
  const int *xy[4]; const double *so[4], *rms[4];
  double tol16, tol8;
  b2fr_node *nodes;
};

// chun by one warp: the sums and squared deviations are exact in double (order-free), so they are reduced in parallel;
// the 256 correlation terms are formed in parallel (two divisions each) and then added by ONE lane in the reference's
// column-major order, the only order-dependent part.
__device__ double fd_chun_warp(const uint8_t *org, const uint8_t *ref, int w, int bx, int by, double *term /* smem [256] */)
{
  const int lane = threadIdx.x & 31;
  double sumR = 0, sumD = 0;
  for (int ii = lane; ii < 256; ii += 32) {           // ii = column-major index: j = bx + ii / 16, i = by + ii % 16
    const size_t o = (size_t)(by + (ii & 15)) * w + bx + (ii >> 4);
    sumR += (double)org[o]; sumD += (double)ref[o];
  }
  for (int m = 16; m; m >>= 1) { sumR += __shfl_xor_sync(0xffffffffu, sumR, m); sumD += __shfl_xor_sync(0xffffffffu, sumD, m); }
  const double r = __ddiv_rn(sumR, 256.0), d = __ddiv_rn(sumD, 256.0);
  double sR = 0, sD = 0;
  for (int ii = lane; ii < 256; ii += 32) {
    const size_t o = (size_t)(by + (ii & 15)) * w + bx + (ii >> 4);
    const double a = __dsub_rn((double)org[o], r), b = __dsub_rn((double)ref[o], d);
    sR += __dmul_rn(a, a); sD += __dmul_rn(b, b);     // multiples of 2^-16 below 2^24: every partial sum is exact
  }
  for (int m = 16; m; m >>= 1) { sR += __shfl_xor_sync(0xffffffffu, sR, m); sD += __shfl_xor_sync(0xffffffffu, sD, m); }
  const double qR = __dsqrt_rn(sR), qD = __dsqrt_rn(sD);
  for (int ii = lane; ii < 256; ii += 32) {
    const size_t o = (size_t)(by + (ii & 15)) * w + bx + (ii >> 4);
    const double a = __dsub_rn((double)org[o], r), b = __dsub_rn((double)ref[o], d);
    term[ii] = __dmul_rn(__ddiv_rn(a, qR), __ddiv_rn(b, qD));
  }
  __syncwarp();
  double mr = 0;
  if (lane == 0)
    for (int ii = 0; ii < 256; ii++) mr = __dadd_rn(mr, term[ii]);
  mr = __shfl_sync(0xffffffffu, mr, 0);
  __syncwarp();
  return __dmul_rn(mr, mr);
}

// the four searches of one block: C into the node, then H, M, N with strict '<'
__device__ double fd_search4(const FdArgs &a, size_t base, int p, b2fr_node &t, bool is4x4)
{
  const size_t o = base + p;
  double rms = a.rms[0][o];
  if (a.xy[0][2 * o] || a.xy[0][2 * o + 1]) { t.x = a.xy[0][2 * o]; t.y = a.xy[0][2 * o + 1]; }     // Q-F11
  t.scale = a.so[0][2 * o]; t.offset = a.so[0][2 * o + 1];
  for (int s = 1; s < 4; s++) {
    const double r = a.rms[s][o];
    if (r < rms) {
      if (s == 1 && is4x4) t.partition = 1;
      t.reference = s; rms = r;
      t.x = a.xy[s][2 * o]; t.y = a.xy[s][2 * o + 1]; t.scale = a.so[s][2 * o]; t.offset = a.so[s][2 * o + 1];
      t.block_type = 0;
    } else if (s == 1) t.reference = 0;
  }
  return rms;
}

__global__ void __launch_bounds__(128) k_frac_decide(const FdArgs a)
{
  // one WARP per macroblock: chun across the lanes, the replay of the tree by lane 0
  __shared__ double terms[4][256];
  const int mb = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (mb >= a.nmb) return;
  const double chun = fd_chun_warp(a.org, a.refC, a.w, (mb % a.mbw) * 16, (mb / a.mbw) * 16, terms[threadIdx.x >> 5]);
  if (threadIdx.x & 31) return;
  b2fr_node nd[21];
  for (int i = 0; i < 21; i++) { nd[i].block_type = nd[i].partition = nd[i].reference = nd[i].x = nd[i].y = nd[i].reserved = 0; nd[i].scale = nd[i].offset = 0.0; }
  const size_t base = (size_t)mb * NPART;
  const int bx = (mb % a.mbw) * 16, by = (mb / a.mbw) * 16;
  const double t16 = __dmul_rn(__dmul_rn(a.tol16, a.tol16), 256.0);
  const double r16 = fd_search4(a, base, 0, nd[0], false);
  if (chun <= 1.0 && chun >= 0.9 && r16 > t16) {
    const double tt = __dmul_rn(a.tol8, a.tol8);
    for (int mode = 1; mode < 3; mode++) {
      nd[0].partition = mode;
      for (int i = 0; i < 2; i++) {
        b2fr_node &c = nd[1 + 5 * i];
        c.x = 0; c.y = 0;
        if (fd_search4(a, base, (mode == 1 ? 1 : 3) + i, c, false) > __dmul_rn(tt, 128.0)) break;
      }
    }
    nd[0].partition = 3;
    for (int k = 0; k < 4; k++) {
      b2fr_node *c = &nd[1 + 5 * k];
      const int by2 = k >> 1, bx2 = k & 1;
      c->partition = 0; c->reference = 0; c->x = 0; c->y = 0;
      if (!(fd_search4(a, base, 5 + k, *c, false) > __dmul_rn(tt, 64.0))) continue;
      int mode;
      for (mode = 1; mode < 3; mode++) {
        int ok = 0;
        c->partition = mode;
        for (int i = 0; i < 2; i++) {
          b2fr_node &g = c[1 + i];
          const int p = mode == 1 ? 9 + 2 * (2 * by2 + i) + bx2 : 17 + 4 * by2 + 2 * bx2 + i;
          g.x = 0; g.y = 0;
          if (fd_search4(a, base, p, g, false) > __dmul_rn(tt, 32.0)) break;
          ok++;
        }
        if (ok == 2) mode = 4;
      }
      if (mode < 4) {
        c->partition = 3;
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++) {
            b2fr_node &g = c[1 + i * 2 + j];
            g.x = g.y = 0;
            fd_search4(a, base, 25 + 4 * (2 * by2 + i) + 2 * bx2 + j, g, true);
          }
      }
    }
  }
  for (int i = 0; i < 21; i++) a.nodes[(size_t)mb * 21 + i] = nd[i];
}

// ---- F8: the fractal prediction of every macroblock ------------------------------------------------------------
// decode_one_macroblock / decode_block_rect / decode_block_8 / decode_block_4 (V1/src/block_dec.c:20, 285, 760, 978;
// num_regions == 1): each leaf block of the TRANS_NODE tree is rec = (unsigned char) bound(0.5 + scale * d + offset -
// scale * mean_d), d = the displaced block of plane set `reference`, mean_d = its sum table entry / n (:228-232).
// One CTA per macroblock, one thread per pixel; the domain block's sum comes from the 4x4 sliding sums (exact integers,
// like the reference's tables), the doubles follow the reference's operation order.
struct FpArgs {
  const uint8_t *ref[4]; const int *s4[4]; int w, h, mbw;
  const b2fr_node *nodes; uint8_t *out;
};

__global__ void __launch_bounds__(256) k_frac_predict(const FpArgs a)
{
  const int mb = blockIdx.x, px = threadIdx.x & 15, py = threadIdx.x >> 4;
  const b2fr_node *root = a.nodes + (size_t)mb * 21;
  const b2fr_node *t = root;
  int bx = 0, by = 0, bw = 16, bh = 16, mb_level = 1;
  if (root->partition == 1) { const int i = py >> 3; t = root + 1 + 5 * i; by = 8 * i; bh = 8; mb_level = 0; }
  else if (root->partition == 2) { const int i = px >> 3; t = root + 1 + 5 * i; bx = 8 * i; bw = 8; mb_level = 0; }
  else if (root->partition != 0) {
    const int k = (py >> 3) * 2 + (px >> 3);
    const b2fr_node *c = root + 1 + 5 * k;
    bx = (k & 1) * 8; by = (k >> 1) * 8; bw = bh = 8; mb_level = 0; t = c;
    if (c->partition == 1) { const int i = (py & 7) >> 2; t = c + 1 + i; by += 4 * i; bh = 4; }
    else if (c->partition == 2) { const int i = (px & 7) >> 2; t = c + 1 + i; bx += 4 * i; bw = 4; }
    else if (c->partition != 0) { const int i = ((py & 7) >> 2) * 2 + ((px & 7) >> 2); t = c + 1 + i; bx += (i & 1) * 4; by += (i >> 1) * 4; bw = bh = 4; }
  }
  const int r = t->reference, set = (r >= 0 && r <= 3) ? r : (mb_level ? 0 : 3);
  const int x0 = (mb % a.mbw) * 16, y0 = (mb / a.mbw) * 16;
  const int ox = x0 + bx + t->x, oy = y0 + by + t->y;
  int sum = 0;                                        // (coordinates clamped: foreign trees must not read outside the plane)
  for (int j = 0; j < bh; j += 4)
    for (int i = 0; i < bw; i += 4) sum += a.s4[set][(size_t)iclamp(oy + j, 0, a.h - 1) * a.w + iclamp(ox + i, 0, a.w - 1)];
  const double avg = __ddiv_rn((double)sum, (double)(bh * bw));
  const double d = (double)a.ref[set][(size_t)iclamp(oy + py - by, 0, a.h - 1) * a.w + iclamp(ox + px - bx, 0, a.w - 1)];
  const double v = __dsub_rn(__dadd_rn(__dadd_rn(0.5, __dmul_rn(t->scale, d)), t->offset), __dmul_rn(t->scale, avg));
  a.out[(size_t)(y0 + py) * a.w + x0 + px] = (uint8_t)(v < 0.0 ? 0 : (v > 255.0 ? 255 : (int)v));
}

}  // namespace b2

// ================================ C ABI ======================================================
extern "C" const char *b2fr_last_error(b2fr_ctx *c) { return c ? c->err : g_frerr; }

extern "C" int b2fr_create(b2fr_ctx **out, int device, int width, int height, int search_range)
{
  if (!out || width < 16 || height < 16 || (width & 15) || (height & 15) || search_range < 0 || search_range > 32) {
    snprintf(g_frerr, sizeof(g_frerr), "b2fr_create: invalid argument");
    return B2ME_EINVAL;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    snprintf(g_frerr, sizeof(g_frerr), "b2fr_create: no CUDA device %d (%s)", device, cudaGetErrorString(e));
    return B2ME_ECUDA;
  }
  b2fr_ctx *c = new (std::nothrow) b2fr_ctx();
  if (!c) return B2ME_ENOMEM;
  memset(c, 0, sizeof(*c));
  c->device = device; c->W = width; c->H = height; c->R = search_range;
  *out = c;
  FR_CHECK(c, cudaSetDevice(device));
  FR_CHECK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int k = 0; k < 3; k++) {
    const size_t np = (size_t)comp_w(c, k) * comp_h(c, k);
    const int nmb = comp_mbw(c, k) * comp_mbh(c, k);
    FR_CHECK(c, cudaMalloc(&c->d_org[k], np));
    FR_CHECK(c, cudaMemset(c->d_org[k], 0, np));
    FR_CHECK(c, cudaMalloc(&c->d_rs4[k], np / 16 * sizeof(int)));
    FR_CHECK(c, cudaMalloc(&c->d_rq4[k], np / 16 * sizeof(int)));
    FR_CHECK(c, cudaMemset(c->d_rs4[k], 0, np / 16 * sizeof(int)));
    FR_CHECK(c, cudaMemset(c->d_rq4[k], 0, np / 16 * sizeof(int)));
    for (int s = 0; s < 4; s++) {
      FR_CHECK(c, cudaMalloc(&c->d_ref[s][k], np));
      FR_CHECK(c, cudaMemset(c->d_ref[s][k], 0, np));                 // calloc'ed planes (Q-F3)
      FR_CHECK(c, cudaMalloc(&c->d_s4[s][k], np * sizeof(int)));
      FR_CHECK(c, cudaMalloc(&c->d_q4[s][k], np * sizeof(int)));
      FR_CHECK(c, cudaMemset(c->d_s4[s][k], 0, np * sizeof(int)));    // tables start zero and stay zero until built
      FR_CHECK(c, cudaMemset(c->d_q4[s][k], 0, np * sizeof(int)));
      const size_t nr = (size_t)(nmb > 0 ? nmb : 1) * NPART;
      FR_CHECK(c, cudaMalloc(&c->d_xy[s][k], nr * 2 * sizeof(int)));
      FR_CHECK(c, cudaMalloc(&c->d_so[s][k], nr * 2 * sizeof(double)));
      FR_CHECK(c, cudaMalloc(&c->d_rms[s][k], nr * sizeof(double)));
      FR_CHECK(c, cudaMallocHost(&c->h_xy[s][k], nr * 2 * sizeof(int)));
      FR_CHECK(c, cudaMallocHost(&c->h_so[s][k], nr * 2 * sizeof(double)));
      FR_CHECK(c, cudaMallocHost(&c->h_rms[s][k], nr * sizeof(double)));
    }
  }
  FR_CHECK(c, cudaMalloc(&c->d_tab, (size_t)width * height * sizeof(int)));
  return B2ME_OK;
}

extern "C" void b2fr_destroy(b2fr_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  for (int k = 0; k < 3; k++) {
    cudaFree(c->d_org[k]); cudaFree(c->d_rs4[k]); cudaFree(c->d_rq4[k]);
    for (int s = 0; s < 4; s++) {
      cudaFree(c->d_ref[s][k]); cudaFree(c->d_s4[s][k]); cudaFree(c->d_q4[s][k]);
      cudaFree(c->d_xy[s][k]); cudaFree(c->d_so[s][k]); cudaFree(c->d_rms[s][k]);
      cudaFreeHost(c->h_xy[s][k]); cudaFreeHost(c->h_so[s][k]); cudaFreeHost(c->h_rms[s][k]);
    }
  }
  cudaFree(c->d_tab); cudaFree(c->d_rec);
  for (int k = 0; k < 3; k++) cudaFree(c->d_nodes[k]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

static void invalidate(b2fr_ctx *c, int set)
{
  for (int s = 0; s < 4; s++)
    if (set < 0 || s == set) for (int k = 0; k < 3; k++) c->valid[s][k] = c->dvalid[s][k] = 0;
  for (int k = 0; k < 3; k++) c->nodes_valid[k] = 0;       // trees of other pictures
}

extern "C" int b2fr_set_range(b2fr_ctx *c, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
  if (!c || !y) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const uint8_t *src[3] = {y, u, v};
  for (int k = 0; k < 3; k++) {
    if (!src[k]) continue;
    const int w = comp_w(c, k), h = comp_h(c, k);
    FR_CHECK(c, cudaMemcpyAsync(c->d_org[k], src[k], (size_t)w * h, cudaMemcpyHostToDevice, c->stream));
    const int n = (w / 4) * (h / 4);
    k_frac_range_sums<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_org[k], w, h, c->d_rs4[k], c->d_rq4[k]);
    c->launches++;
  }
  FR_CHECK(c, cudaGetLastError());
  FR_CHECK(c, cudaStreamSynchronize(c->stream));
  invalidate(c, -1);
  return B2ME_OK;
}

extern "C" int b2fr_set_domain(b2fr_ctx *c, int plane_set, const uint8_t *y, const uint8_t *u, const uint8_t *v, int build_sums)
{
  if (!c || plane_set < 0 || plane_set > 3 || !y) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const uint8_t *src[3] = {y, u, v};
  for (int k = 0; k < 3; k++) {
    if (!src[k]) continue;
    const int w = comp_w(c, k), h = comp_h(c, k);
    FR_CHECK(c, cudaMemcpyAsync(c->d_ref[plane_set][k], src[k], (size_t)w * h, cudaMemcpyHostToDevice, c->stream));
    c->loaded[plane_set][k] = 1;
    if (build_sums) {
      if ((size_t)w * h >= (size_t)1 << 20) {
        dim3 grid((w + 255) / 256, (h + 63) / 64);
        k_frac_domain_sums<64><<<grid, 256, 0, c->stream>>>(c->d_ref[plane_set][k], w, h, c->d_s4[plane_set][k], c->d_q4[plane_set][k]);
      } else {
        dim3 grid((w + 255) / 256, (h + 15) / 16);
        k_frac_domain_sums<16><<<grid, 256, 0, c->stream>>>(c->d_ref[plane_set][k], w, h, c->d_s4[plane_set][k], c->d_q4[plane_set][k]);
      }
      c->launches++;
    }
  }
  FR_CHECK(c, cudaGetLastError());
  FR_CHECK(c, cudaStreamSynchronize(c->stream));
  invalidate(c, plane_set);
  return B2ME_OK;
}

static int run_window(b2fr_ctx *c, int set, int comp)
{
  const int mbw = comp_mbw(c, comp), mbh = comp_mbh(c, comp);
  if (mbw * mbh == 0) { snprintf(c->err, sizeof(c->err), "component %d has no whole macroblock", comp); return B2ME_EINVAL; }
  FrArgs a;
  a.org = c->d_org[comp]; a.ref = c->d_ref[set][comp];
  a.s4 = c->d_s4[set][comp]; a.q4 = c->d_q4[set][comp]; a.rs4 = c->d_rs4[comp]; a.rq4 = c->d_rq4[comp];
  // A plane set the caller never loaded is all zero with all-zero tables (the shipped program's H/M/N sets, SURVEY Q-F3):
  // every candidate of a block then has the same alpha / beta / rms, the strict '<' of full_search never leaves the start
  // candidate, and searching with range 0 gives the identical result at 1/225 of the work.
  const int R = c->loaded[set][comp] ? c->R : 0;
  a.w = comp_w(c, comp); a.h = comp_h(c, comp); a.mbw = mbw; a.R = R;
  a.xy = c->d_xy[set][comp]; a.so = c->d_so[set][comp]; a.rms = c->d_rms[set][comp];
  const int ww = 16 + 2 * R;
  const int smem = 256 + ww * ww;
  k_frac_window<<<mbw * mbh, FR_NT, smem, c->stream>>>(a);
  c->launches++;
  FR_CHECK(c, cudaGetLastError());
  return B2ME_OK;
}

// results of (set, comp) on the device only (what the cascade needs); ensure() adds the host copy the look-up calls read
static int ensure_dev(b2fr_ctx *c, int set, int comp)
{
  if (c->dvalid[set][comp]) return B2ME_OK;
  int r = run_window(c, set, comp);
  if (r) return r;
  c->dvalid[set][comp] = 1;
  return B2ME_OK;
}

static int ensure(b2fr_ctx *c, int set, int comp)
{
  if (c->valid[set][comp]) return B2ME_OK;
  int r = ensure_dev(c, set, comp);
  if (r) return r;
  const size_t nr = (size_t)comp_mbw(c, comp) * comp_mbh(c, comp) * NPART;
  FR_CHECK(c, cudaMemcpyAsync(c->h_xy[set][comp], c->d_xy[set][comp], nr * 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FR_CHECK(c, cudaMemcpyAsync(c->h_so[set][comp], c->d_so[set][comp], nr * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  FR_CHECK(c, cudaMemcpyAsync(c->h_rms[set][comp], c->d_rms[set][comp], nr * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  FR_CHECK(c, cudaStreamSynchronize(c->stream));
  c->valid[set][comp] = 1;
  return B2ME_OK;
}

extern "C" int b2fr_search_plane(b2fr_ctx *c, int plane_set, int con, int32_t *xy, double *scale_offset, double *rms)
{
  if (!c || plane_set < 0 || plane_set > 3 || con < 1 || con > 3 || !xy || !scale_offset || !rms) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const int comp = con - 1;
  int r = ensure(c, plane_set, comp);
  if (r) return r;
  const size_t nr = (size_t)comp_mbw(c, comp) * comp_mbh(c, comp) * NPART;
  memcpy(xy, c->h_xy[plane_set][comp], nr * 2 * sizeof(int));
  memcpy(scale_offset, c->h_so[plane_set][comp], nr * 2 * sizeof(double));
  memcpy(rms, c->h_rms[plane_set][comp], nr * sizeof(double));
  return B2ME_OK;
}

extern "C" int b2fr_full_search(b2fr_ctx *c, int plane_set, int block_x, int block_y, int block_size_x, int block_size_y,
                                int con, int32_t xy[2], double scale_offset[2], double *rms)
{
  if (!c || plane_set < 0 || plane_set > 3 || con < 1 || con > 3 || !xy || !scale_offset || !rms) return B2ME_EINVAL;
  const int comp = con - 1;
  const int w = comp_w(c, comp), h = comp_h(c, comp);
  if (block_x < 0 || block_y < 0 || block_x + block_size_x > w || block_y + block_size_y > h) return B2ME_EINVAL;
  int part = -1;
  for (int p = 0; p < NPART; p++) {
    const PartGeom g = part_geom(p);
    if (g.w == block_size_x && g.h == block_size_y && g.ox == (block_x & 15) && g.oy == (block_y & 15)) { part = p; break; }
  }
  const int mbx = block_x >> 4, mby = block_y >> 4;
  if (part < 0 || mbx >= comp_mbw(c, comp) || mby >= comp_mbh(c, comp)) {
    snprintf(c->err, sizeof(c->err), "block (%d,%d) %dx%d is not on the range grid", block_x, block_y, block_size_x, block_size_y);
    return B2ME_EINVAL;
  }
  FR_CHECK(c, cudaSetDevice(c->device));
  int r = ensure(c, plane_set, comp);
  if (r) return r;
  const size_t o = (size_t)(mby * comp_mbw(c, comp) + mbx) * NPART + part;
  const int *hx = c->h_xy[plane_set][comp];
  // Q-F11: x,y are written only when a displaced candidate won
  if (hx[2 * o] != 0 || hx[2 * o + 1] != 0) { xy[0] = hx[2 * o]; xy[1] = hx[2 * o + 1]; }
  scale_offset[0] = c->h_so[plane_set][comp][2 * o]; scale_offset[1] = c->h_so[plane_set][comp][2 * o + 1];
  *rms = c->h_rms[plane_set][comp][o];
  return B2ME_OK;
}

extern "C" int b2fr_encode_plane(b2fr_ctx *c, int con, const double tol[3], b2fr_node *nodes)
{
  if (!c || con < 1 || con > 3 || !tol || !nodes) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const int comp = con - 1;
  for (int s = 0; s < 4; s++) { int r = ensure_dev(c, s, comp); if (r) return r; }
  const int mbw = comp_mbw(c, comp), nmb = mbw * comp_mbh(c, comp);
  if (!c->d_nodes[comp]) FR_CHECK(c, cudaMalloc(&c->d_nodes[comp], sizeof(b2fr_node) * 21 * (size_t)nmb));
  b2fr_node *d = c->d_nodes[comp];
  FdArgs a;
  a.org = c->d_org[comp]; a.refC = c->d_ref[0][comp]; a.w = comp_w(c, comp); a.mbw = mbw; a.nmb = nmb;
  for (int s = 0; s < 4; s++) { a.xy[s] = c->d_xy[s][comp]; a.so[s] = c->d_so[s][comp]; a.rms[s] = c->d_rms[s][comp]; }
  a.tol16 = tol[0]; a.tol8 = tol[1]; a.nodes = d;
  k_frac_decide<<<(nmb + 3) / 4, 128, 0, c->stream>>>(a);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(nodes, d, sizeof(b2fr_node) * 21 * (size_t)nmb, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "b2fr_encode_plane: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  c->nodes_valid[comp] = 1;
  return B2ME_OK;
}

extern "C" int b2fr_decode_plane(b2fr_ctx *c, int con, const b2fr_node *nodes, uint8_t *rec)
{
  if (!c || con < 1 || con > 3 || !rec) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const int comp = con - 1, w = comp_w(c, comp), h = comp_h(c, comp);
  const int mbw = comp_mbw(c, comp), nmb = mbw * comp_mbh(c, comp);
  if (nmb == 0) { snprintf(c->err, sizeof(c->err), "component %d has no whole macroblock", comp); return B2ME_EINVAL; }
  if (!c->d_nodes[comp]) FR_CHECK(c, cudaMalloc(&c->d_nodes[comp], sizeof(b2fr_node) * 21 * (size_t)nmb));
  if (nodes) {
    // the displaced blocks must lie inside the plane (the search's bound_chk guarantees it for its own trees)
    for (int mb = 0; mb < nmb; mb++)
      for (int k = 0; k < 21; k++) {
        const b2fr_node &t = nodes[(size_t)mb * 21 + k];
        const int x0 = (mb % mbw) * 16, y0 = (mb / mbw) * 16;
        if (x0 + t.x < -16 || y0 + t.y < -16 || x0 + t.x > w || y0 + t.y > h || abs(t.x) > 64 || abs(t.y) > 64) {
          snprintf(c->err, sizeof(c->err), "b2fr_decode_plane: node %d of macroblock %d displaces its block out of the plane", k, mb);
          return B2ME_EINVAL;
        }
      }
    FR_CHECK(c, cudaMemcpyAsync(c->d_nodes[comp], nodes, sizeof(b2fr_node) * 21 * (size_t)nmb, cudaMemcpyHostToDevice, c->stream));
    c->nodes_valid[comp] = 1;
  } else if (!c->nodes_valid[comp]) {
    snprintf(c->err, sizeof(c->err), "b2fr_decode_plane: no trees on the device for component %d (call b2fr_encode_plane or pass nodes)", con);
    return B2ME_EINVAL;
  }
  if (!c->d_rec) FR_CHECK(c, cudaMalloc(&c->d_rec, (size_t)c->W * c->H));
  FR_CHECK(c, cudaMemsetAsync(c->d_rec, 0, (size_t)w * h, c->stream));
  FpArgs a;
  for (int s = 0; s < 4; s++) { a.ref[s] = c->d_ref[s][comp]; a.s4[s] = c->d_s4[s][comp]; }
  a.w = w; a.h = h; a.mbw = mbw; a.nodes = c->d_nodes[comp]; a.out = c->d_rec;
  k_frac_predict<<<nmb, 256, 0, c->stream>>>(a);
  c->launches++;
  FR_CHECK(c, cudaGetLastError());
  FR_CHECK(c, cudaMemcpyAsync(rec, c->d_rec, (size_t)w * h, cudaMemcpyDeviceToHost, c->stream));
  FR_CHECK(c, cudaStreamSynchronize(c->stream));
  return B2ME_OK;
}

extern "C" int b2fr_get_domain_table(b2fr_ctx *c, int plane_set, int con, int bw, int bh, int squares, int32_t *out)
{
  if (!c || plane_set < 0 || plane_set > 3 || con < 1 || con > 3 || !out || (bw & 3) || (bh & 3) || bw < 4 || bh < 4 || bw > 16 || bh > 16) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const int comp = con - 1, w = comp_w(c, comp), h = comp_h(c, comp);
  k_frac_table<<<(w * h + 255) / 256, 256, 0, c->stream>>>(squares ? c->d_q4[plane_set][comp] : c->d_s4[plane_set][comp], w, h, bw, bh, c->d_tab);
  c->launches++;
  FR_CHECK(c, cudaGetLastError());
  FR_CHECK(c, cudaMemcpyAsync(out, c->d_tab, (size_t)w * h * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FR_CHECK(c, cudaStreamSynchronize(c->stream));
  return B2ME_OK;
}

extern "C" int b2fr_get_range_table(b2fr_ctx *c, int con, int squares, int32_t *out)
{
  if (!c || con < 1 || con > 3 || !out) return B2ME_EINVAL;
  FR_CHECK(c, cudaSetDevice(c->device));
  const int comp = con - 1;
  const size_t n = (size_t)(comp_w(c, comp) / 4) * (comp_h(c, comp) / 4);
  FR_CHECK(c, cudaMemcpy(out, squares ? c->d_rq4[comp] : c->d_rs4[comp], n * sizeof(int), cudaMemcpyDeviceToHost));
  return B2ME_OK;
}

extern "C" int64_t b2fr_launch_count(b2fr_ctx *c) { return c ? c->launches : 0; }
