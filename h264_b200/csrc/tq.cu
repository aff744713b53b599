// tq.cu -- residual transform + quantisation + reconstruction of 4x4 / 8x8 luma blocks (b2tq_*).
//
// Replaces   residual_transform_quant_luma_4x4   JM/lencod/src/block.c:660-724   (check_zero :626)
//            residual_transform_quant_luma_8x8   JM/lencod/src/transform8x8.c:522-602
//            forward4x4 / inverse4x4 / forward8x8 / inverse8x8   JM/lcommon/src/transform.c:20,70,353,450
//            quant_4x4_normal / quant_8x8_normal JM/lencod/src/quant4x4_normal.c:39, quant8x8_normal.c:43
//            sample_reconstruct                  JM/lcommon/src/blk_prediction.c:48
//            dct_luma (version1, mode 1)         V1/src/block.c:836-1045
//
// One thread per block; blocks are independent, the kernel is HBM-bound (48 B in / 73 B out per 4x4
// block) and all loads/stores are 16-byte vectors of consecutive blocks.  The run/level lists are
// the reference's ACLevel/ACRun arrays (zero-terminated, zero-padded here).
#include <cstring>
#include "b2_common.cuh"
#include "../../include/b2me.h"

namespace b2 {


// scan tables: {i (horizontal), j (vertical)}  JM/lencod/src/block.c:169-184, transform8x8.c:44-72
__device__ constexpr uint8_t ZZ4[16][2] = {{0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{1,3},{2,2},{3,1},{3,2},{2,3},{3,3}};
__device__ constexpr uint8_t FS4[16][2] = {{0,0},{0,1},{1,0},{0,2},{0,3},{1,1},{1,2},{1,3},{2,0},{2,1},{2,2},{2,3},{3,0},{3,1},{3,2},{3,3}};
__device__ constexpr uint8_t ZZ8[64][2] = {
  {0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{0,4},{1,3},{2,2},{3,1},{4,0},{5,0},
  {4,1},{3,2},{2,3},{1,4},{0,5},{0,6},{1,5},{2,4},{3,3},{4,2},{5,1},{6,0},{7,0},{6,1},{5,2},{4,3},
  {3,4},{2,5},{1,6},{0,7},{1,7},{2,6},{3,5},{4,4},{5,3},{6,2},{7,1},{7,2},{6,3},{5,4},{4,5},{3,6},
  {2,7},{3,7},{4,6},{5,5},{6,4},{7,3},{7,4},{6,5},{5,6},{4,7},{5,7},{6,6},{7,5},{7,6},{6,7},{7,7}};
__device__ constexpr uint8_t FS8[64][2] = {
  {0,0},{0,1},{0,2},{1,0},{1,1},{0,3},{0,4},{1,2},{2,0},{1,3},{0,5},{0,6},{0,7},{1,4},{2,1},{3,0},
  {2,2},{1,5},{1,6},{1,7},{2,3},{3,1},{4,0},{3,2},{2,4},{2,5},{2,6},{2,7},{3,3},{4,1},{5,0},{4,2},
  {3,4},{3,5},{3,6},{3,7},{4,3},{5,1},{6,0},{5,2},{4,4},{4,5},{4,6},{4,7},{5,3},{6,1},{6,2},{5,4},
  {5,5},{5,6},{5,7},{6,3},{7,0},{7,1},{6,4},{6,5},{6,6},{6,7},{7,2},{7,3},{7,4},{7,5},{7,6},{7,7}};
// coefficient cost by run (block.c:72-76, transform8x8.c:83-92); disthres 1: constant 9
__device__ __forceinline__ int cost4(int run, int disthres) { return disthres ? 9 : (run < 1 ? 3 : run < 3 ? 2 : run < 6 ? 1 : 0); }
__device__ __forceinline__ int cost8(int run, int disthres) { return disthres ? 9 : (run < 4 ? 3 : run < 12 ? 2 : run < 24 ? 1 : 0); }

__device__ __forceinline__ void fwd4(int &a, int &b, int &c, int &d)
{
  const int t0 = a + d, t1 = b + c, t2 = b - c, t3 = a - d;
  a = t0 + t1; b = (t3 << 1) + t2; c = t0 - t1; d = t3 - (t2 << 1);
}
__device__ __forceinline__ void inv4(int &a, int &b, int &c, int &d)
{
  const int p0 = a + c, p1 = a - c, p2 = (b >> 1) - d, p3 = b + (d >> 1);
  a = p0 + p3; b = p1 + p2; c = p1 - p2; d = p0 - p3;
}
__device__ __forceinline__ void fwd8(int *v, int s)   // in place on v[0], v[s], ..., v[7s]
{
  const int p0 = v[0], p1 = v[s], p2 = v[2 * s], p3 = v[3 * s], p4 = v[4 * s], p5 = v[5 * s], p6 = v[6 * s], p7 = v[7 * s];
  int a0 = p0 + p7, a1 = p1 + p6, a2 = p2 + p5, a3 = p3 + p4;
  const int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
  a0 = p0 - p7; a1 = p1 - p6; a2 = p2 - p5; a3 = p3 - p4;
  const int b4 = a1 + a2 + ((a0 >> 1) + a0), b5 = a0 - a3 - ((a2 >> 1) + a2);
  const int b6 = a0 + a3 - ((a1 >> 1) + a1), b7 = a1 - a2 + ((a3 >> 1) + a3);
  v[0] = b0 + b1; v[s] = b4 + (b7 >> 2); v[2 * s] = b2 + (b3 >> 1); v[3 * s] = b5 + (b6 >> 2);
  v[4 * s] = b0 - b1; v[5 * s] = b6 - (b5 >> 2); v[6 * s] = (b2 >> 1) - b3; v[7 * s] = (b4 >> 2) - b7;
}
__device__ __forceinline__ void inv8(int *v, int s)
{
  const int p0 = v[0], p1 = v[s], p2 = v[2 * s], p3 = v[3 * s], p4 = v[4 * s], p5 = v[5 * s], p6 = v[6 * s], p7 = v[7 * s];
  int a0 = p0 + p4, a1 = p0 - p4, a2 = p6 - (p2 >> 1), a3 = p2 + (p6 >> 1);
  const int b0 = a0 + a3, b2 = a1 - a2, b4 = a1 + a2, b6 = a0 - a3;
  a0 = -p3 + p5 - p7 - (p7 >> 1); a1 = p1 + p7 - p3 - (p3 >> 1); a2 = -p1 + p7 + p5 + (p5 >> 1); a3 = p3 + p5 + p1 + (p1 >> 1);
  const int b1 = a0 + (a3 >> 2), b3 = a1 + (a2 >> 2), b5 = a2 - (a1 >> 2), b7 = a3 - (a0 >> 2);
  v[0] = b0 + b7; v[s] = b2 - b5; v[2 * s] = b4 + b3; v[3 * s] = b6 + b1;
  v[4 * s] = b6 - b1; v[5 * s] = b4 - b3; v[6 * s] = b2 + b5; v[7 * s] = b0 - b7;
}
__device__ __forceinline__ int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

template <bool FIELD>
__global__ void __launch_bounds__(128) k_tq4x4(const __grid_constant__ b2tq_params c_tq, int nblk, const uint4 *__restrict__ orig, const uint4 *__restrict__ pred,
                                               uint4 *__restrict__ level, uint4 *__restrict__ run, uint4 *__restrict__ recon,
                                               int *__restrict__ cost, uint8_t *__restrict__ nonzero)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nblk) return;
  const uint4 o4 = orig[k], p4 = pred[k];
  const uint32_t ow[4] = {o4.x, o4.y, o4.z, o4.w}, pw[4] = {p4.x, p4.y, p4.z, p4.w};
  int x[16], pr[16];
  int any = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    pr[i] = (pw[i >> 2] >> (8 * (i & 3))) & 255;
    x[i] = (int)((ow[i >> 2] >> (8 * (i & 3))) & 255) - pr[i];
    any |= x[i];
  }
  const int mode = c_tq.mode, qp_per = c_tq.qp / 6, q_bits = 15 + qp_per;
  __align__(16) short lev[16]; __align__(16) unsigned char rn[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { lev[i] = 0; rn[i] = 0; }
  int nz = 0, cst = 0, n = 0;
  if (any != 0 || mode == 1) {          // JM skips all-zero residual blocks (check_zero); version1 never does
#pragma unroll
    for (int r = 0; r < 4; r++) fwd4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
#pragma unroll
    for (int c = 0; c < 4; c++) fwd4(x[c], x[4 + c], x[8 + c], x[12 + c]);
    int runc = 0;
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int i = FIELD ? FS4[s][0] : ZZ4[s][0], j = FIELD ? FS4[s][1] : ZZ4[s][1], idx = j * 4 + i;
      const int m7 = x[idx];
      int lv = 0;
      if (m7 != 0 || mode == 1) {
        const int am = m7 < 0 ? -m7 : m7;
        lv = (am * c_tq.scale[idx] + c_tq.offset[idx]) >> q_bits;
      }
      if (lv != 0) {
        if (c_tq.cavlc && mode == 0 && lv > 2063) lv = 2063;
        cst += (lv > 1) ? 999999 : cost4(runc, c_tq.disthres);
        const int sl = m7 < 0 ? -lv : lv;
        // JM: ((level*InvScaleComp) << qp_per) + 8 >> 4 with InvScaleComp = dequant<<4  ==  V1: level*dequant << qp_per
        x[idx] = mode == 0 ? ((((sl * c_tq.invscale[idx]) << qp_per) + 8) >> 4) : (m7 < 0 ? -((lv * c_tq.invscale[idx]) << qp_per) : ((lv * c_tq.invscale[idx]) << qp_per));
        lev[n] = (short)sl; rn[n] = (unsigned char)runc; n++;
        runc = 0; nz = 1;
      } else { x[idx] = 0; runc++; }
    }
  }
  uint32_t rw[4] = {p4.x, p4.y, p4.z, p4.w};
  if (nz || mode == 1) {
#pragma unroll
    for (int r = 0; r < 4; r++) inv4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
#pragma unroll
    for (int c = 0; c < 4; c++) inv4(x[c], x[4 + c], x[8 + c], x[12 + c]);
#pragma unroll
    for (int w = 0; w < 4; w++) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = 4 * w + b;
        // JM: clip(((r + 32) >> 6) + pred)  ==  V1: clip((r + (pred << 6) + 32) >> 6)
        v |= (uint32_t)clip255(((x[i] + 32) >> 6) + pr[i]) << (8 * b);
      }
      rw[w] = v;
    }
  }
  recon[k] = make_uint4(rw[0], rw[1], rw[2], rw[3]);
  const uint4 *lv4 = reinterpret_cast<const uint4 *>(lev);
  level[2 * k] = lv4[0]; level[2 * k + 1] = lv4[1];
  run[k] = *reinterpret_cast<const uint4 *>(rn);
  cost[k] = cst; nonzero[k] = (uint8_t)nz;
}

// 8x8 transpose across the eight lanes of a block's group (lane t holds row t -> lane t holds column t): three exchange rounds,
// every register index is a compile-time constant.
__device__ __forceinline__ void transpose8(int (&x)[8], int t)
{
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {
    const bool up = (t & s) != 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i & s) continue;
      const int a = x[i], b = x[i | s];
      const int recv = __shfl_xor_sync(0xffffffffu, up ? a : b, s);
      if (up) x[i] = recv; else x[i | s] = recv;
    }
  }
}

// residual_transform_quant_luma_8x8: EIGHT threads per block (round 1 ran one thread per block on a 64-entry register array: 186
// registers, 18.6 % issue, 3 x the time of k_tq4x4 per byte).  Thread t of a group loads row t (8 bytes of orig and pred), runs the
// horizontal pass in registers, the group transposes through shuffles, the vertical pass and the quantiser run on column t.  The
// run / level list comes from a 64-bit mask of the nonzero levels in scan order (OR over the group): a level's slot in the list
// is the number of set bits below its scan position, its run the distance to the set bit before it; the lists are staged in
// shared memory and leave as 16-byte / 8-byte stores.  inverse8x8 is horizontal first (JM/lcommon/src/transform.c:450): back to
// rows, to columns, and to rows again for the reconstruction.  16 blocks per 128-thread CTA.
template <bool FIELD>
__global__ void __launch_bounds__(128) k_tq8x8(const __grid_constant__ b2tq_params c_tq, int nblk, const uint2 *__restrict__ orig, const uint2 *__restrict__ pred,
                                               uint4 *__restrict__ level, uint2 *__restrict__ run, uint2 *__restrict__ recon,
                                               int *__restrict__ cost, uint8_t *__restrict__ nonzero)
{
  __shared__ int s_scale[64], s_offset[64], s_inv[64];
  __shared__ uint8_t s_pos[64];                        // raster index j*8+i -> scan position
  __shared__ __align__(16) short s_lev[16][64];
  __shared__ __align__(8) uint8_t s_run[16][64];
  const int tid = threadIdx.x, t = tid & 7, g = tid >> 3;
  if (tid < 64) {
    s_scale[tid] = c_tq.scale[tid]; s_offset[tid] = c_tq.offset[tid]; s_inv[tid] = c_tq.invscale[tid];
    const int i = FIELD ? FS8[tid][0] : ZZ8[tid][0], j = FIELD ? FS8[tid][1] : ZZ8[tid][1];
    s_pos[j * 8 + i] = (uint8_t)tid;
  }
  reinterpret_cast<uint4 *>(s_lev[g])[t] = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint2 *>(s_run[g])[t] = make_uint2(0u, 0u);
  __syncthreads();
  const int k = blockIdx.x * 16 + g;
  const bool live = k < nblk;
  const int kk = live ? k : nblk - 1;                  // idle groups shadow the last block (the shuffles want every lane)
  const uint2 o2 = orig[(size_t)kk * 8 + t], p2 = pred[(size_t)kk * 8 + t];
  int x[8];
  int any = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t ow = i < 4 ? o2.x : o2.y, pw = i < 4 ? p2.x : p2.y;
    x[i] = (int)((ow >> (8 * (i & 3))) & 255) - (int)((pw >> (8 * (i & 3))) & 255);
    any |= x[i];
  }
  const uint32_t gmask = 0xffu << (8 * ((tid >> 3) & 3));      // this group's lanes of the warp
  any = (__ballot_sync(0xffffffffu, any != 0) & gmask) != 0u;
  const int qp_per = c_tq.qp / 6, q_bits = 16 + qp_per;
  unsigned long long M = 0ull;
  int cst = 0;
  if (__any_sync(0xffffffffu, any)) {                  // warp-uniform: some block of the warp has a residual
    fwd8(x, 1);                                        // row t
    transpose8(x, t);
    fwd8(x, 1);                                        // column t: x[j] = coefficient (i = t, j)
    int lv[8];
    uint32_t mlo = 0u, mhi = 0u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int idx = j * 8 + t, m7 = x[j], am = m7 < 0 ? -m7 : m7;
      int l = any ? (am * s_scale[idx] + s_offset[idx]) >> q_bits : 0;      // a block without residual is skipped (mode 0)
      lv[j] = m7 < 0 ? -l : l;
      const int p = s_pos[idx];
      if (l) { if (p < 32) mlo |= 1u << p; else mhi |= 1u << (p - 32); }
    }
#pragma unroll
    for (int s2 = 1; s2 < 8; s2 <<= 1) { mlo |= __shfl_xor_sync(0xffffffffu, mlo, s2); mhi |= __shfl_xor_sync(0xffffffffu, mhi, s2); }
    M = ((unsigned long long)mhi << 32) | mlo;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int idx = j * 8 + t;
      if (lv[j]) {
        const int p = s_pos[idx];
        const unsigned long long below = M & ((1ull << p) - 1ull);
        const int rank = __popcll(below), prev = below ? 63 - __clzll((long long)below) : -1, r = p - prev - 1;
        const int al = lv[j] < 0 ? -lv[j] : lv[j];
        cst += al > 1 ? 999999 : cost8(r, c_tq.disthres);
        s_lev[g][rank] = (short)lv[j]; s_run[g][rank] = (uint8_t)r;
        x[j] = (((lv[j] * s_inv[idx]) << qp_per) + 32) >> 6;
      } else x[j] = 0;
    }
#pragma unroll
    for (int s2 = 1; s2 < 8; s2 <<= 1) cst += __shfl_xor_sync(0xffffffffu, cst, s2);
  }
  __syncwarp();
  if (live) {
    level[(size_t)k * 8 + t] = reinterpret_cast<const uint4 *>(s_lev[g])[t];
    run[(size_t)k * 8 + t] = reinterpret_cast<const uint2 *>(s_run[g])[t];
  }
  const bool nz = M != 0ull;
  uint2 rec = p2;
  if (__any_sync(0xffffffffu, nz)) {
    transpose8(x, t);                                  // row t
    inv8(x, 1);
    transpose8(x, t);                                  // column t
    inv8(x, 1);
    transpose8(x, t);                                  // row t again
    if (nz) {
      uint32_t w[2] = {0u, 0u};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint32_t pw = i < 4 ? p2.x : p2.y;
        w[i >> 2] |= (uint32_t)clip255(((x[i] + 32) >> 6) + (int)((pw >> (8 * (i & 3))) & 255)) << (8 * (i & 3));
      }
      rec = make_uint2(w[0], w[1]);
    }
  }
  if (live) {
    recon[(size_t)k * 8 + t] = rec;
    if (t == 0) { cost[k] = cst; nonzero[k] = (uint8_t)nz; }
  }
}

// ---- Intra16x16 luma: residual_transform_quant_luma_16x16 (JM/lencod/src/block.c:207-345) -----------------------------------------
// 16 threads per macroblock (one per 4x4 block, raster), two macroblocks per warp.  forward4x4 per thread; the sixteen DC
// coefficients form a 4x4 matrix across the threads: hadamard4x4 / ihadamard4x4 (JM/lcommon/src/transform.c:121-214) take each
// thread's row and column through shuffles; quant_dc4x4_normal (quant4x4_normal.c:200-270) and its run/level list come from a
// ballot over the scan order; quant_ac4x4_normal (:117-190, scan positions 1..15), inverse4x4 and sample_reconstruct stay in the
// thread.  orig / pred / recon: [nmb][256] raster 16x16.
template <bool FIELD>
__global__ void __launch_bounds__(128) k_tq16x16(const __grid_constant__ b2tq_params c_tq, int nmb, const uint8_t *__restrict__ orig, const uint8_t *__restrict__ pred,
                                                 short *__restrict__ dc_level, uint8_t *__restrict__ dc_run, short *__restrict__ ac_level, uint8_t *__restrict__ ac_run,
                                                 uint8_t *__restrict__ recon, uint8_t *__restrict__ ac_coef)
{
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int mb = gid >> 4, b = gid & 15, lane = threadIdx.x & 31, hb = lane & 16;      // hb: first lane of this macroblock's half warp
  const bool live = mb < nmb;
  const int jj = b >> 2, ii = b & 3;
  const int qp_per = c_tq.qp / 6, q_bits = 15 + qp_per;
  int x[16], pr[16];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const size_t o = (size_t)mb * 256 + (4 * jj + r) * 16 + 4 * ii;
    const uint32_t ow = live ? *reinterpret_cast<const uint32_t *>(orig + o) : 0u, pw = live ? *reinterpret_cast<const uint32_t *>(pred + o) : 0u;
#pragma unroll
    for (int c = 0; c < 4; c++) { pr[4 * r + c] = (pw >> (8 * c)) & 255; x[4 * r + c] = (int)((ow >> (8 * c)) & 255) - pr[4 * r + c]; }
  }
#pragma unroll
  for (int r = 0; r < 4; r++) fwd4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
#pragma unroll
  for (int c = 0; c < 4; c++) fwd4(x[c], x[4 + c], x[8 + c], x[12 + c]);
  // ---- hadamard4x4 of the DC matrix: thread (jj, ii) ends with tblock[jj][ii] ----
  int dcv;
  {
    int p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = __shfl_sync(0xffffffffu, x[0], hb + 4 * jj + k);
    const int t0 = p[0] + p[3], t1 = p[1] + p[2], t2 = p[1] - p[2], t3 = p[0] - p[3];
    const int h = ii == 0 ? t0 + t1 : ii == 1 ? t3 + t2 : ii == 2 ? t0 - t1 : t3 - t2;
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = __shfl_sync(0xffffffffu, h, hb + 4 * k + ii);
    const int u0 = p[0] + p[3], u1 = p[1] + p[2], u2 = p[1] - p[2], u3 = p[0] - p[3];
    dcv = (jj == 0 ? u0 + u1 : jj == 1 ? u2 + u3 : jj == 2 ? u0 - u1 : u3 - u2) >> 1;
  }
  // ---- quant_dc4x4_normal: every thread quantises its own coefficient; the lists follow the scan order ----
  int dlev = 0;
  if (dcv != 0) {
    int lv = ((dcv < 0 ? -dcv : dcv) * c_tq.scale[0] + (c_tq.offset[0] << 1)) >> (q_bits + 1);
    if (lv != 0) { if (c_tq.cavlc && lv > 2063) lv = 2063; dlev = dcv < 0 ? -lv : lv; }
  }
  {
    const int si = FIELD ? FS4[b][0] : ZZ4[b][0], sj = FIELD ? FS4[b][1] : ZZ4[b][1];          // thread b doubles as scan position b
    const int slev = __shfl_sync(0xffffffffu, dlev, hb + sj * 4 + si);
    const uint32_t mask = (__ballot_sync(0xffffffffu, slev != 0) >> hb) & 0xffffu;
    if (live) { dc_level[(size_t)mb * 16 + b] = 0; dc_run[(size_t)mb * 16 + b] = 0; }
    __syncwarp();
    if (live && slev != 0) {
      const uint32_t below = mask & ((1u << b) - 1u);
      const int idx = __popc(below), run = below ? b - 1 - (31 - __clz(below)) : b;
      dc_level[(size_t)mb * 16 + idx] = (short)slev; dc_run[(size_t)mb * 16 + idx] = (uint8_t)run;
    }
    // ---- inverse DC transform + DC dequantisation (only when a DC level is nonzero) ----
    int dcr = 0;
    if (mask) {                                     // uniform per half warp; the shuffles below run for the whole warp
    }
    {
      int p[4];
#pragma unroll
      for (int k = 0; k < 4; k++) p[k] = __shfl_sync(0xffffffffu, dlev, hb + 4 * jj + k);
      const int a0 = p[0] + p[2], a1 = p[0] - p[2], a2 = p[1] - p[3], a3 = p[1] + p[3];
      const int h = ii == 0 ? a0 + a3 : ii == 1 ? a1 + a2 : ii == 2 ? a1 - a2 : a0 - a3;
#pragma unroll
      for (int k = 0; k < 4; k++) p[k] = __shfl_sync(0xffffffffu, h, hb + 4 * k + ii);
      const int c0 = p[0] + p[2], c1 = p[0] - p[2], c2 = p[1] - p[3], c3 = p[1] + p[3];
      const int v = jj == 0 ? c0 + c3 : jj == 1 ? c1 + c2 : jj == 2 ? c1 - c2 : c0 - c3;
      dcr = mask ? (((v * c_tq.invscale[0]) << qp_per) + 32) >> 6 : 0;
    }
    x[0] = dcr;
  }
  // ---- quant_ac4x4_normal of this thread's block ----
  __align__(16) short lev[16]; __align__(16) unsigned char rn[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { lev[i] = 0; rn[i] = 0; }
  int nz = 0, n = 0, runc = 0;
#pragma unroll
  for (int s = 1; s < 16; s++) {
    const int i = FIELD ? FS4[s][0] : ZZ4[s][0], j = FIELD ? FS4[s][1] : ZZ4[s][1], idx = j * 4 + i;
    const int m7 = x[idx];
    int lv = 0;
    if (m7 != 0) lv = ((m7 < 0 ? -m7 : m7) * c_tq.scale[idx] + c_tq.offset[idx]) >> q_bits;
    if (lv != 0) {
      if (c_tq.cavlc && lv > 2063) lv = 2063;
      const int sl = m7 < 0 ? -lv : lv;
      x[idx] = (((sl * c_tq.invscale[idx]) << qp_per) + 8) >> 4;
      lev[n] = (short)sl; rn[n] = (unsigned char)runc; n++;
      runc = 0; nz = 1;
    } else { x[idx] = 0; runc++; }
  }
  if (x[0] != 0 || nz) {
#pragma unroll
    for (int r = 0; r < 4; r++) inv4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
#pragma unroll
    for (int c = 0; c < 4; c++) inv4(x[c], x[4 + c], x[8 + c], x[12 + c]);
  }
  const uint32_t anyac = (__ballot_sync(0xffffffffu, nz != 0) >> hb) & 0xffffu;
  if (!live) return;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) v |= (uint32_t)clip255(((x[4 * r + c] + 32) >> 6) + pr[4 * r + c]) << (8 * c);
    *reinterpret_cast<uint32_t *>(recon + (size_t)mb * 256 + (4 * jj + r) * 16 + 4 * ii) = v;
  }
  const uint4 *lv4 = reinterpret_cast<const uint4 *>(lev);
  uint4 *lo = reinterpret_cast<uint4 *>(ac_level + ((size_t)mb * 16 + b) * 16);
  lo[0] = lv4[0]; lo[1] = lv4[1];
  *reinterpret_cast<uint4 *>(ac_run + ((size_t)mb * 16 + b) * 16) = *reinterpret_cast<const uint4 *>(rn);
  if (b == 0) ac_coef[mb] = anyac ? 15 : 0;
}

// ---- chroma of 4:2:0 macroblocks, one plane: residual_transform_quant_chroma_4x4 (JM/lencod/src/block.c:953-1200) -------------------
// Four threads per 8x8 chroma block (one per 4x4 block, raster), eight blocks per warp.  forward4x4 per thread; the four DC
// coefficients go through hadamard2x2 (JM/lcommon/src/transform.c:302-315), quant_dc2x2_normal (quantChroma_normal.c:37-96: natural
// order, offset << 1, q_bits + 1, levels de-quantised in place), ihadamard2x2 (:317-331) and >> 5 with shuffles; quant_ac4x4_normal
// (quant4x4_normal.c:117-190) per thread with its coefficient cost, summed over the four threads for the _CHROMA_COEFF_COST_ rule
// (block.c:1137-1168: fewer / smaller AC levels than that are all dropped); inverse4x4 and sample_reconstruct stay in the thread.
template <bool FIELD>
__global__ void __launch_bounds__(128) k_tq_chroma(const __grid_constant__ b2tq_params c_tq, int nmb, const uint8_t *__restrict__ orig, const uint8_t *__restrict__ pred,
                                                   short *__restrict__ dc_level, uint8_t *__restrict__ dc_run, short *__restrict__ ac_level, uint8_t *__restrict__ ac_run,
                                                   uint8_t *__restrict__ recon, uint8_t *__restrict__ cr_cbp)
{
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int mb = gid >> 2, b = gid & 3, lane = threadIdx.x & 31, q0 = lane & ~3;        // q0: first lane of this block's quad
  const bool live = mb < nmb;
  const int jj = b >> 1, ii = b & 1;
  const int qp_per = c_tq.qp / 6, q_bits = 15 + qp_per;
  int x[16], pr[16];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const size_t o = (size_t)mb * 64 + (4 * jj + r) * 8 + 4 * ii;
    const uint32_t ow = live ? *reinterpret_cast<const uint32_t *>(orig + o) : 0u, pw = live ? *reinterpret_cast<const uint32_t *>(pred + o) : 0u;
#pragma unroll
    for (int c = 0; c < 4; c++) { pr[4 * r + c] = (pw >> (8 * c)) & 255; x[4 * r + c] = (int)((ow >> (8 * c)) & 255) - pr[4 * r + c]; }
  }
#pragma unroll
  for (int r = 0; r < 4; r++) fwd4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
#pragma unroll
  for (int c = 0; c < 4; c++) fwd4(x[c], x[4 + c], x[8 + c], x[12 + c]);
  // ---- hadamard2x2: thread b ends with coefficient b of the 2x2 DC block ----
  int d[4];
#pragma unroll
  for (int k = 0; k < 4; k++) d[k] = __shfl_sync(0xffffffffu, x[0], q0 + k);
  int m = b == 0 ? d[0] + d[1] + d[2] + d[3] : b == 1 ? d[0] - d[1] + d[2] - d[3] : b == 2 ? d[0] + d[1] - d[2] - d[3] : d[0] - d[1] - d[2] + d[3];
  // ---- quant_dc2x2_normal (natural order); the level list through a ballot ----
  int dlev = 0;
  if (m != 0) {
    int lv = ((m < 0 ? -m : m) * c_tq.scale[0] + (c_tq.offset[0] << 1)) >> (q_bits + 1);
    if (lv != 0) { if (c_tq.cavlc && lv > 2063) lv = 2063; dlev = m < 0 ? -lv : lv; }
  }
  m = (dlev * c_tq.invscale[0]) << qp_per;
  const uint32_t dmask = (__ballot_sync(0xffffffffu, dlev != 0) >> q0) & 0xfu;
  if (live) { dc_level[(size_t)mb * 4 + b] = 0; dc_run[(size_t)mb * 4 + b] = 0; }
  __syncwarp();
  if (live && dlev != 0) {
    const uint32_t below = dmask & ((1u << b) - 1u);
    const int idx = __popc(below), run = below ? b - 1 - (31 - __clz(below)) : b;
    dc_level[(size_t)mb * 4 + idx] = (short)dlev; dc_run[(size_t)mb * 4 + idx] = (uint8_t)run;
  }
  // ---- ihadamard2x2, >> 5 ----
#pragma unroll
  for (int k = 0; k < 4; k++) d[k] = __shfl_sync(0xffffffffu, m, q0 + k);
  x[0] = (b == 0 ? d[0] + d[1] + d[2] + d[3] : b == 1 ? d[0] - d[1] + d[2] - d[3] : b == 2 ? d[0] + d[1] - d[2] - d[3] : d[0] - d[1] - d[2] + d[3]) >> 5;
  // ---- quant_ac4x4_normal of this thread's block ----
  __align__(16) short lev[16]; __align__(16) unsigned char rn[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { lev[i] = 0; rn[i] = 0; }
  int nz = 0, n = 0, runc = 0, cost = 0;
#pragma unroll
  for (int s = 1; s < 16; s++) {
    const int i = FIELD ? FS4[s][0] : ZZ4[s][0], j = FIELD ? FS4[s][1] : ZZ4[s][1], idx = j * 4 + i;
    const int m7 = x[idx];
    int lv = 0;
    if (m7 != 0) lv = ((m7 < 0 ? -m7 : m7) * c_tq.scale[idx] + c_tq.offset[idx]) >> q_bits;
    if (lv != 0) {
      if (c_tq.cavlc && lv > 2063) lv = 2063;
      cost += lv > 1 ? 999999 : cost4(runc, c_tq.disthres);      // COEFF_COST4x4 (block.c:72-76)
      const int sl = m7 < 0 ? -lv : lv;
      x[idx] = (((sl * c_tq.invscale[idx]) << qp_per) + 8) >> 4;
      lev[n] = (short)sl; rn[n] = (unsigned char)runc; n++;
      runc = 0; nz = 1;
    } else { x[idx] = 0; runc++; }
  }
  // ---- _CHROMA_COEFF_COST_: the four blocks' cost together ----
  int tot = cost > 999999 ? 999999 : cost;
  tot += __shfl_xor_sync(0xffffffffu, tot, 1);
  tot += __shfl_xor_sync(0xffffffffu, tot, 2);
  const uint32_t anynz = (__ballot_sync(0xffffffffu, nz != 0) >> q0) & 0xfu;
  const bool drop = anynz && tot < 4;
  if (drop && nz) {
    nz = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) x[i] = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) lev[i] = 0;
  }
  if (x[0] != 0 || nz) {
#pragma unroll
    for (int r = 0; r < 4; r++) inv4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
#pragma unroll
    for (int c = 0; c < 4; c++) inv4(x[c], x[4 + c], x[8 + c], x[12 + c]);
  }
  if (!live) return;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) v |= (uint32_t)clip255(((x[4 * r + c] + 32) >> 6) + pr[4 * r + c]) << (8 * c);
    *reinterpret_cast<uint32_t *>(recon + (size_t)mb * 64 + (4 * jj + r) * 8 + 4 * ii) = v;
  }
  const uint4 *lv4 = reinterpret_cast<const uint4 *>(lev);
  uint4 *lo = reinterpret_cast<uint4 *>(ac_level + ((size_t)mb * 4 + b) * 16);
  lo[0] = lv4[0]; lo[1] = lv4[1];
  *reinterpret_cast<uint4 *>(ac_run + ((size_t)mb * 4 + b) * 16) = *reinterpret_cast<const uint4 *>(rn);
  if (b == 0) cr_cbp[mb] = (uint8_t)((anynz && !drop) ? 2 : (dmask ? 1 : 0));
}

}  // namespace b2

using namespace b2;

static char g_tqerr[256] = "";
extern "C" const char *b2tq_last_error(void) { return g_tqerr; }
#define TQ_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { snprintf(g_tqerr, sizeof(g_tqerr), "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); return B2ME_ECUDA; } } while (0)

static int check_tq(const b2tq_params *p, int n8)
{
  if (!p || p->qp < 0 || p->qp > 51 || (p->mode != 0 && p->mode != 1) || (n8 && p->mode != 0)) {
    snprintf(g_tqerr, sizeof(g_tqerr), "b2tq: invalid parameter block");
    return B2ME_EINVAL;
  }
  return B2ME_OK;
}

extern "C" int b2tq_4x4_dev(const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level, uint8_t *run,
                            uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero, void *stream)
{
  int r = check_tq(p, 0);
  if (r) return r;
  if (nblk < 0 || !orig || !pred || !level || !run || !recon || !coeff_cost || !nonzero) return B2ME_EINVAL;
  if (nblk == 0) return B2ME_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (nblk + 127) / 128;
  if (p->field_scan) k_tq4x4<true><<<grid, 128, 0, s>>>(*p, nblk, (const uint4 *)orig, (const uint4 *)pred, (uint4 *)level, (uint4 *)run, (uint4 *)recon, coeff_cost, nonzero);
  else k_tq4x4<false><<<grid, 128, 0, s>>>(*p, nblk, (const uint4 *)orig, (const uint4 *)pred, (uint4 *)level, (uint4 *)run, (uint4 *)recon, coeff_cost, nonzero);
  TQ_CHECK(cudaGetLastError());
  return B2ME_OK;
}

extern "C" int b2tq_8x8_dev(const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level, uint8_t *run,
                            uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero, void *stream)
{
  int r = check_tq(p, 1);
  if (r) return r;
  if (nblk < 0 || !orig || !pred || !level || !run || !recon || !coeff_cost || !nonzero) return B2ME_EINVAL;
  if (nblk == 0) return B2ME_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (nblk + 15) / 16;          // eight threads per block, 16 blocks per CTA
  if (p->field_scan) k_tq8x8<true><<<grid, 128, 0, s>>>(*p, nblk, (const uint2 *)orig, (const uint2 *)pred, (uint4 *)level, (uint2 *)run, (uint2 *)recon, coeff_cost, nonzero);
  else k_tq8x8<false><<<grid, 128, 0, s>>>(*p, nblk, (const uint2 *)orig, (const uint2 *)pred, (uint4 *)level, (uint2 *)run, (uint2 *)recon, coeff_cost, nonzero);
  TQ_CHECK(cudaGetLastError());
  return B2ME_OK;
}

// host-pointer variants: stage through device buffers owned by the call
constexpr size_t TQ_SMALL = (size_t)1 << 20;
static int tq_host(int n8, int device, const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level,
                   uint8_t *run, uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero)
{
  int r = check_tq(p, n8);
  if (r) return r;
  if (nblk < 0 || !orig || !pred || !level || !run || !recon || !coeff_cost || !nonzero) return B2ME_EINVAL;
  if (nblk == 0) return B2ME_OK;
  TQ_CHECK(cudaSetDevice(device));
  const size_t px = n8 ? 64 : 16, N = (size_t)nblk;
  uint8_t *d = nullptr;
  // layout: orig | pred | recon | run | level | cost | nonzero   (each 16-byte aligned)
  const size_t o_pred = N * px, o_rec = 2 * N * px, o_run = 3 * N * px, o_lev = 4 * N * px, o_cost = 6 * N * px, o_nz = o_cost + ((N * 4 + 15) & ~(size_t)15);
  const size_t total = o_nz + N + 16;
  if (total <= TQ_SMALL) {
    // small batches (the per-block calls of a drop-in encoder): a per-thread device scratch and pinned staging buffer, one
    // copy up, one launch, one copy down -- no allocation per call
    static thread_local struct { int device; uint8_t *d, *h; cudaStream_t s; } T = {-1, nullptr, nullptr, nullptr};
    if (T.device != device) {
      if (T.d) { cudaFree(T.d); cudaFreeHost(T.h); cudaStreamDestroy(T.s); T.d = nullptr; }
      TQ_CHECK(cudaMalloc(&T.d, TQ_SMALL)); TQ_CHECK(cudaMallocHost(&T.h, TQ_SMALL));
      TQ_CHECK(cudaStreamCreateWithFlags(&T.s, cudaStreamNonBlocking));
      T.device = device;
    }
    memcpy(T.h, orig, N * px); memcpy(T.h + o_pred, pred, N * px);
    TQ_CHECK(cudaMemcpyAsync(T.d, T.h, 2 * N * px, cudaMemcpyHostToDevice, T.s));
    r = n8 ? b2tq_8x8_dev(p, nblk, T.d, T.d + o_pred, (int16_t *)(T.d + o_lev), T.d + o_run, T.d + o_rec, (int32_t *)(T.d + o_cost), T.d + o_nz, T.s)
           : b2tq_4x4_dev(p, nblk, T.d, T.d + o_pred, (int16_t *)(T.d + o_lev), T.d + o_run, T.d + o_rec, (int32_t *)(T.d + o_cost), T.d + o_nz, T.s);
    if (r) return r;
    TQ_CHECK(cudaMemcpyAsync(T.h + o_rec, T.d + o_rec, total - o_rec, cudaMemcpyDeviceToHost, T.s));
    TQ_CHECK(cudaStreamSynchronize(T.s));
    memcpy(recon, T.h + o_rec, N * px); memcpy(run, T.h + o_run, N * px); memcpy(level, T.h + o_lev, N * px * 2);
    memcpy(coeff_cost, T.h + o_cost, N * 4); memcpy(nonzero, T.h + o_nz, N);
    return B2ME_OK;
  }
  TQ_CHECK(cudaMalloc(&d, total));
  cudaStream_t s = 0;
  cudaError_t e = cudaMemcpyAsync(d, orig, N * px, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_pred, pred, N * px, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) { cudaFree(d); TQ_CHECK(e); }
  r = n8 ? b2tq_8x8_dev(p, nblk, d, d + o_pred, (int16_t *)(d + o_lev), d + o_run, d + o_rec, (int32_t *)(d + o_cost), d + o_nz, s)
         : b2tq_4x4_dev(p, nblk, d, d + o_pred, (int16_t *)(d + o_lev), d + o_run, d + o_rec, (int32_t *)(d + o_cost), d + o_nz, s);
  if (r == B2ME_OK) {
    e = cudaMemcpyAsync(recon, d + o_rec, N * px, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(run, d + o_run, N * px, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(level, d + o_lev, N * px * 2, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(coeff_cost, d + o_cost, N * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(nonzero, d + o_nz, N, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  }
  cudaFree(d);
  return r;
}
extern "C" int b2tq_4x4(int device, const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level, uint8_t *run,
                        uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero)
{ return tq_host(0, device, p, nblk, orig, pred, level, run, recon, coeff_cost, nonzero); }
extern "C" int b2tq_8x8(int device, const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred, int16_t *level, uint8_t *run,
                        uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero)
{ return tq_host(1, device, p, nblk, orig, pred, level, run, recon, coeff_cost, nonzero); }

// Intra16x16 luma macroblocks (see include/b2me.h)
extern "C" int b2tq_16x16_dev(const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                              int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *ac_coef, void *stream)
{
  int r = check_tq(p, 0);
  if (r) return r;
  if (p->mode != 0) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq_16x16: version1 has no Intra16x16 DC path (mode must be 0)"); return B2ME_EUNSUPPORTED; }
  if (nmb < 0 || !orig || !pred || !dc_level || !dc_run || !ac_level || !ac_run || !recon || !ac_coef) return B2ME_EINVAL;
  if (nmb == 0) return B2ME_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (nmb * 16 + 127) / 128;
  if (p->field_scan) k_tq16x16<true><<<grid, 128, 0, s>>>(*p, nmb, orig, pred, dc_level, dc_run, ac_level, ac_run, recon, ac_coef);
  else k_tq16x16<false><<<grid, 128, 0, s>>>(*p, nmb, orig, pred, dc_level, dc_run, ac_level, ac_run, recon, ac_coef);
  TQ_CHECK(cudaGetLastError());
  return B2ME_OK;
}
extern "C" int b2tq_16x16(int device, const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                          int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *ac_coef)
{
  if (nmb < 0 || !orig || !pred || !dc_level || !dc_run || !ac_level || !ac_run || !recon || !ac_coef) return B2ME_EINVAL;
  if (nmb == 0) return B2ME_OK;
  TQ_CHECK(cudaSetDevice(device));
  const size_t N = (size_t)nmb;
  // layout: orig 256 | pred 256 | recon 256 | ac_run 256 | ac_level 512 | dc_level 32 | dc_run 16 | ac_coef 1 (+ pad) per macroblock
  uint8_t *d = nullptr;
  const size_t o_pred = N * 256, o_rec = N * 512, o_arun = N * 768, o_alev = N * 1024, o_dlev = N * 1536, o_drun = N * 1568, o_ac = N * 1584, total = N * 1585 + 16;
  TQ_CHECK(cudaMalloc(&d, total));
  cudaError_t e = cudaMemcpy(d, orig, N * 256, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + o_pred, pred, N * 256, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); TQ_CHECK(e); }
  int r = b2tq_16x16_dev(p, nmb, d, d + o_pred, (int16_t *)(d + o_dlev), d + o_drun, (int16_t *)(d + o_alev), d + o_arun, d + o_rec, d + o_ac, 0);
  if (r == B2ME_OK) {
    e = cudaMemcpy(recon, d + o_rec, N * 256, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ac_run, d + o_arun, N * 256, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ac_level, d + o_alev, N * 512, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dc_level, d + o_dlev, N * 32, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dc_run, d + o_drun, N * 16, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ac_coef, d + o_ac, N, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq_16x16: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  }
  cudaFree(d);
  return r;
}

// intra: 0 inter block, 1 intra block of a P/B slice, 2 intra block of an I slice (the reference's
// default offset lists give 682 only to the last: q_offsets.c:426-470, CalculateOffset4x4Param :487-561).
// Default parameter block: flat quantiser matrices (quant_coef / dequant_coef, q_matrix.c:20-36,
// 38-167) and the default rounding offsets 682 (intra) / 342 (inter) << (q_bits - 11)
// (q_offsets.c:60-87, 163-188; OffsetBits = 11).  mode 1 = version1: offset (1 << q_bits) / 3.
extern "C" int b2tq_default_params(b2tq_params *p, int is8x8, int qp, int intra, int mode)
{
  static const int qc[6][3] = {{13107, 5243, 8066}, {11916, 4660, 7490}, {10082, 4194, 6554}, {9362, 3647, 5825}, {8192, 3355, 5243}, {7282, 2893, 4559}};
  static const int dq[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
  static const int qc8[6][6] = {{13107, 11428, 20972, 12222, 16777, 15481}, {11916, 10826, 19174, 11058, 14980, 14290},
                                {10082, 8943, 15978, 9675, 12710, 11985}, {9362, 8228, 14913, 8931, 11984, 11259},
                                {8192, 7346, 13159, 7740, 10486, 9777}, {7282, 6428, 11570, 6830, 9118, 8640}};
  static const int dq8[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
                                {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
  if (!p || qp < 0 || qp > 51 || intra < 0 || intra > 2 || (mode != 0 && mode != 1) || (is8x8 && mode == 1)) return B2ME_EINVAL;
  memset(p, 0, sizeof(*p));
  p->qp = qp; p->mode = mode; p->cavlc = 1;
  const int rem = qp % 6, per = qp / 6;
  if (!is8x8) {
    const int q_bits = 15 + per;
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 4; i++) {
        const int cls = ((i & 1) && (j & 1)) ? 1 : (!(i & 1) && !(j & 1)) ? 0 : 2;
        p->scale[j * 4 + i] = qc[rem][cls];
        p->invscale[j * 4 + i] = mode == 0 ? dq[rem][cls] << 4 : dq[rem][cls];
        p->offset[j * 4 + i] = mode == 0 ? (intra == 2 ? 682 : 342) << (q_bits - 11) : (1 << q_bits) / 3;
      }
  } else {
    const int q_bits = 16 + per;
    for (int j = 0; j < 8; j++)
      for (int i = 0; i < 8; i++) {
        // position classes of the 8x8 matrices (q_matrix.c:38-167)
        const int a = i & 3, b = j & 3;
        int cls;
        if (a == 0 && b == 0) cls = 0;
        else if ((a & 1) && (b & 1)) cls = 1;
        else if (a == 2 && b == 2) cls = 2;
        else if ((a == 0 && (b & 1)) || ((a & 1) && b == 0)) cls = 3;
        else if ((a == 0 && b == 2) || (a == 2 && b == 0)) cls = 4;
        else cls = 5;
        p->scale[j * 8 + i] = qc8[rem][cls];
        p->invscale[j * 8 + i] = dq8[rem][cls] << 4;
        p->offset[j * 8 + i] = (intra == 2 ? 682 : 342) << (q_bits - 11);
      }
  }
  return B2ME_OK;
}

// chroma (4:2:0) blocks of one plane (see include/b2me.h)
extern "C" int b2tq_chroma_dev(const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                               int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *cr_cbp, void *stream)
{
  if (!p || nmb < 0 || !orig || !pred || !dc_level || !dc_run || !ac_level || !ac_run || !recon || !cr_cbp) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq_chroma: bad arguments"); return B2ME_EINVAL; }
  if (p->mode != 0) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq_chroma: version1 has no chroma DC path (mode must be 0)"); return B2ME_EUNSUPPORTED; }
  { int r = check_tq(p, 0); if (r) return r; }
  if (nmb == 0) return B2ME_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (nmb * 4 + 127) / 128;
  if (p->field_scan) k_tq_chroma<true><<<grid, 128, 0, s>>>(*p, nmb, orig, pred, dc_level, dc_run, ac_level, ac_run, recon, cr_cbp);
  else k_tq_chroma<false><<<grid, 128, 0, s>>>(*p, nmb, orig, pred, dc_level, dc_run, ac_level, ac_run, recon, cr_cbp);
  TQ_CHECK(cudaGetLastError());
  return B2ME_OK;
}
extern "C" int b2tq_chroma(int device, const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                           int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *cr_cbp)
{
  if (!p || nmb < 0 || !orig || !pred || !dc_level || !dc_run || !ac_level || !ac_run || !recon || !cr_cbp) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq_chroma: bad arguments"); return B2ME_EINVAL; }
  if (nmb == 0) return B2ME_OK;
  TQ_CHECK(cudaSetDevice(device));
  const size_t n = (size_t)nmb;
  const size_t o_pred = n * 64, o_rec = o_pred + n * 64, o_dlev = o_rec + n * 64, o_drun = o_dlev + n * 8, o_alev = (o_drun + n * 4 + 15) & ~(size_t)15,
               o_arun = o_alev + n * 128, o_cbp = o_arun + n * 64, total = o_cbp + n;
  uint8_t *d = nullptr;
  TQ_CHECK(cudaMalloc(&d, total));
  cudaMemcpy(d, orig, n * 64, cudaMemcpyHostToDevice); cudaMemcpy(d + o_pred, pred, n * 64, cudaMemcpyHostToDevice);
  int r = b2tq_chroma_dev(p, nmb, d, d + o_pred, (int16_t *)(d + o_dlev), d + o_drun, (int16_t *)(d + o_alev), d + o_arun, d + o_rec, d + o_cbp, 0);
  if (!r) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(recon, d + o_rec, n * 64, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dc_level, d + o_dlev, n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dc_run, d + o_drun, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ac_level, d + o_alev, n * 128, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ac_run, d + o_arun, n * 64, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(cr_cbp, d + o_cbp, n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { snprintf(g_tqerr, sizeof(g_tqerr), "b2tq_chroma: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  }
  cudaFree(d);
  return r;
}
