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

template <bool FIELD>
__global__ void __launch_bounds__(64) k_tq8x8(const __grid_constant__ b2tq_params c_tq, int nblk, const uint4 *__restrict__ orig, const uint4 *__restrict__ pred,
                                              short *__restrict__ level, uint8_t *__restrict__ run, uint4 *__restrict__ recon,
                                              int *__restrict__ cost, uint8_t *__restrict__ nonzero)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nblk) return;
  int x[64];
  uint32_t pw[16];
  int any = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint4 o4 = orig[4 * k + q], p4 = pred[4 * k + q];
    const uint32_t ow[4] = {o4.x, o4.y, o4.z, o4.w};
    pw[4 * q] = p4.x; pw[4 * q + 1] = p4.y; pw[4 * q + 2] = p4.z; pw[4 * q + 3] = p4.w;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int idx = 16 * q + i;
      x[idx] = (int)((ow[i >> 2] >> (8 * (i & 3))) & 255) - (int)((pw[4 * q + (i >> 2)] >> (8 * (i & 3))) & 255);
      any |= x[idx];
    }
  }
  const int qp_per = c_tq.qp / 6, q_bits = 16 + qp_per;
  // run/level lists are compacted into thread-local arrays (L1) and leave as 16-byte stores: scattered 2- and 1-byte
  // stores to the thread's own 128 + 64 output bytes made this kernel 4.5x slower than k_tq4x4 per byte
  __align__(16) short lv_out[64];
  __align__(16) uint8_t rn_out[64];
  int nz = 0, cst = 0, n = 0;
  if (any != 0) {
#pragma unroll
    for (int r = 0; r < 8; r++) fwd8(x + 8 * r, 1);
#pragma unroll
    for (int c = 0; c < 8; c++) fwd8(x + c, 8);
    int runc = 0;
#pragma unroll
    for (int s = 0; s < 64; s++) {
      const int i = FIELD ? FS8[s][0] : ZZ8[s][0], j = FIELD ? FS8[s][1] : ZZ8[s][1], idx = j * 8 + i;
      const int m7 = x[idx];
      int lv = 0;
      if (m7 != 0) {
        const int am = m7 < 0 ? -m7 : m7;
        lv = (am * c_tq.scale[idx] + c_tq.offset[idx]) >> q_bits;
      }
      if (lv != 0) {
        cst += (lv > 1) ? 999999 : cost8(runc, c_tq.disthres);
        const int sl = m7 < 0 ? -lv : lv;
        x[idx] = (((sl * c_tq.invscale[idx]) << qp_per) + 32) >> 6;
        lv_out[n] = (short)sl; rn_out[n] = (uint8_t)runc; n++;
        runc = 0; nz = 1;
      } else { x[idx] = 0; runc++; }
    }
  }
  for (int i = n; i < 64; i++) { lv_out[i] = 0; rn_out[i] = 0; }
  {
    uint4 *lo = reinterpret_cast<uint4 *>(level + (size_t)k * 64), *ro = reinterpret_cast<uint4 *>(run + (size_t)k * 64);
#pragma unroll
    for (int i = 0; i < 8; i++) lo[i] = reinterpret_cast<const uint4 *>(lv_out)[i];
#pragma unroll
    for (int i = 0; i < 4; i++) ro[i] = reinterpret_cast<const uint4 *>(rn_out)[i];
  }
  if (nz) {
#pragma unroll
    for (int r = 0; r < 8; r++) inv8(x + 8 * r, 1);
#pragma unroll
    for (int c = 0; c < 8; c++) inv8(x + c, 8);
#pragma unroll
    for (int w = 0; w < 16; w++) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = 4 * w + b;
        v |= (uint32_t)clip255(((x[i] + 32) >> 6) + (int)((pw[w] >> (8 * b)) & 255)) << (8 * b);
      }
      pw[w] = v;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) recon[4 * k + q] = make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
  cost[k] = cst; nonzero[k] = (uint8_t)nz;
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
  const int grid = (nblk + 63) / 64;
  if (p->field_scan) k_tq8x8<true><<<grid, 64, 0, s>>>(*p, nblk, (const uint4 *)orig, (const uint4 *)pred, level, run, (uint4 *)recon, coeff_cost, nonzero);
  else k_tq8x8<false><<<grid, 64, 0, s>>>(*p, nblk, (const uint4 *)orig, (const uint4 *)pred, level, run, (uint4 *)recon, coeff_cost, nonzero);
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
