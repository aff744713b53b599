// deblock.cu -- in-loop deblocking filter of a frame picture on the device.
//
// Replaces DeblockFrame (JM/lencod/src/loopFilter.c:63-111; DeblockMb :196-377) with the non-MBAFF functions of
// JM/lencod/src/loop_filter_normal.c: GetStrengthVer / GetStrengthHor (:52-283), EdgeLoopLumaVer / Hor (:285-585),
// EdgeLoopChromaVer / Hor (:590-758); alpha / beta / clip tables: JM/lencod/inc/loop_filter.h:34-46 (H.264 tables 8-16, 8-17).
//
// The filter is sequential in macroblock order: macroblock (x, y) modifies three sample columns of (x - 1, y) and three rows of
// (x, y - 1), and (x + 1, y - 1) modifies three columns of (x, y - 1) -- including the 3 x 3 corner that (x, y)'s top edge touches
// afterwards.  So (x, y) may start once (x - 1, y) and (x + 1, y - 1) are done: the 2:1 wavefront the reference itself describes
// under JM_PARALLEL_DEBLOCK (loopFilter.c:92-109).  One CTA (one warp) per macroblock ROW walks its row left to right and waits on
// the progress counter of the row above; all rows are resident at once (a row CTA is one warp and 1.3 KB of shared memory), so
// the wait cannot deadlock.  A macroblock is staged in shared memory with the four columns / rows of its left / top neighbours
// (one pass over global memory in, one out), its 32 boundary strengths are computed by the 32 lanes, and the eight edge steps run
// on the staged tile: lanes 0-15 the 16 luma lines of an edge, lanes 16-31 the 2 x 8 chroma lines.
// Sample loads bypass L1 (ld.global.cg): a line fetched for one macroblock also holds samples that the row above rewrites later.
#include <cstdio>
#include "b2_common.cuh"
#include "../../include/b2me.h"

namespace b2 {

__constant__ uint8_t c_dbk_alpha[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
__constant__ uint8_t c_dbk_beta[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
__constant__ uint8_t c_dbk_clip[52][5] = {
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,1,1,1},{0,0,1,1,1},{0,1,1,1,1},
  {0,1,1,1,1},{0,1,1,1,1},{0,1,1,1,1},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,2,3,3},{0,1,2,3,3},{0,2,2,3,3},{0,2,2,4,4},{0,2,3,4,4},
  {0,2,3,4,4},{0,3,3,5,5},{0,3,4,6,6},{0,3,4,6,6},{0,4,5,7,7},{0,4,5,8,8},{0,4,6,9,9},{0,5,7,10,10},{0,6,8,11,11},{0,6,8,13,13},{0,7,10,14,14},{0,8,11,16,16},
  {0,9,12,18,18},{0,10,13,20,20},{0,11,15,23,23},{0,13,17,25,25}};

struct DbkArgs {
  int W, H, mbw, mbh;
  uint8_t *y; int yp; uint8_t *u, *v; int cp;
  const b2dbk_mb *mbs; const b2dbk_blk *blks; int *progress;
};

__device__ __forceinline__ int dbk_mvne(const b2dbk_blk &a, int la, const b2dbk_blk &b, int lb)
{ return (abs(a.mv[la][0] - b.mv[lb][0]) >= 4) | (abs(a.mv[la][1] - b.mv[lb][1]) >= 4); }

// one line of samples across an edge in the staged tile: q[-k * st] = p(k - 1), q[k * st] = q(k)
__device__ __forceinline__ void dbk_luma_line(uint8_t *q, int st, int bs, int alpha, int beta, int c0)
{
  const int p0 = q[-st], p1 = q[-2 * st], p2 = q[-3 * st], q0 = q[0], q1 = q[st], q2 = q[2 * st];
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) {
    const int small_gap = abs(q0 - p0) < ((alpha >> 2) + 2);
    const int ap = (abs(p0 - p2) < beta) & small_gap, aq = (abs(q0 - q2) < beta) & small_gap, s = p0 + q0;
    if (ap) {
      const int p3 = q[-4 * st];
      q[-st] = (uint8_t)((q1 + ((p1 + s) << 1) + p2 + 4) >> 3); q[-2 * st] = (uint8_t)((p2 + p1 + s + 2) >> 2); q[-3 * st] = (uint8_t)((((p3 + p2) << 1) + p2 + p1 + s + 4) >> 3);
    } else q[-st] = (uint8_t)(((p1 << 1) + p0 + q1 + 2) >> 2);
    if (aq) {
      const int q3 = q[3 * st];
      q[0] = (uint8_t)((p1 + ((q1 + s) << 1) + q2 + 4) >> 3); q[st] = (uint8_t)((q2 + q0 + p0 + q1 + 2) >> 2); q[2 * st] = (uint8_t)((((q3 + q2) << 1) + q2 + q1 + s + 4) >> 3);
    } else q[0] = (uint8_t)(((q1 << 1) + q0 + p1 + 2) >> 2);
  } else {
    const int avg = (p0 + q0 + 1) >> 1, ap = abs(p0 - p2) < beta, aq = abs(q0 - q2) < beta, tc = c0 + ap + aq;
    const int dif = iclamp((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    if (ap) q[-2 * st] = (uint8_t)(p1 + iclamp((p2 + avg - (p1 << 1)) >> 1, -c0, c0));
    if (dif) { q[-st] = (uint8_t)iclamp(p0 + dif, 0, 255); q[0] = (uint8_t)iclamp(q0 - dif, 0, 255); }
    if (aq) q[st] = (uint8_t)(q1 + iclamp((q2 + avg - (q1 << 1)) >> 1, -c0, c0));
  }
}
__device__ __forceinline__ void dbk_chroma_line(uint8_t *q, int st, int bs, int alpha, int beta, int c0)
{
  const int p0 = q[-st], p1 = q[-2 * st], q0 = q[0], q1 = q[st];
  if (abs(q0 - p0) >= alpha || abs(q0 - q1) >= beta || abs(p0 - p1) >= beta) return;
  if (bs == 4) { q[-st] = (uint8_t)(((p1 << 1) + p0 + q1 + 2) >> 2); q[0] = (uint8_t)(((q1 << 1) + q0 + p1 + 2) >> 2); }
  else {
    const int tc = c0 + 1, dif = iclamp((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    if (dif) { q[-st] = (uint8_t)iclamp(p0 + dif, 0, 255); q[0] = (uint8_t)iclamp(q0 - dif, 0, 255); }
  }
}

__device__ __forceinline__ uint32_t ld_cg32(const uint8_t *p) { return __ldcg(reinterpret_cast<const uint32_t *>(p)); }

constexpr int DBK_YP = 32, DBK_CP = 16;      // tile pitches

__global__ void __launch_bounds__(32) k_deblock(const DbkArgs a)
{
  __shared__ __align__(16) uint8_t ty[20 * DBK_YP];          // luma rows -4..15, columns -4..15
  __shared__ __align__(16) uint8_t tc[2][12 * DBK_CP];       // chroma rows -4..7, columns -4..7, per plane
  __shared__ uint8_t sbs[2][4][4];
  const int row = blockIdx.x, lane = threadIdx.x;
  const int bw = a.W >> 2;
  for (int mbx = 0; mbx < a.mbw; mbx++) {
    if (row > 0) {                                           // (mbx + 1, row - 1) must be done
      const int need = min(mbx + 2, a.mbw);
      if (lane == 0) {
        const volatile int *pr = a.progress + row - 1;
        while (*pr < need) __nanosleep(40);
      }
      __syncwarp();
      __threadfence();
    }
    const int mb = row * a.mbw + mbx;
    const b2dbk_mb mq = a.mbs[mb];
    if (!mq.disable) {
      const b2dbk_mb ml = mbx ? a.mbs[mb - 1] : mq, mt = row ? a.mbs[mb - a.mbw] : mq;
      // ---- stage the macroblock with four columns / rows of its left / top neighbours ----
      if (lane < 20) {
        const int gy = row * 16 - 4 + lane;
        if (gy >= 0) {
          const uint8_t *src = a.y + (size_t)gy * a.yp + mbx * 16 - 4;
          uint32_t *dst = reinterpret_cast<uint32_t *>(ty + lane * DBK_YP);
#pragma unroll
          for (int k = 0; k < 5; k++) if (k || mbx) dst[k] = ld_cg32(src + 4 * k);
        }
      }
      if (lane < 24) {
        const int pl = lane / 12, r = lane - 12 * pl, gy = row * 8 - 4 + r;
        if (gy >= 0) {
          const uint8_t *src = (pl ? a.v : a.u) + (size_t)gy * a.cp + mbx * 8 - 4;
          uint32_t *dst = reinterpret_cast<uint32_t *>(tc[pl] + r * DBK_CP);
#pragma unroll
          for (int k = 0; k < 3; k++) if (k || mbx) dst[k] = ld_cg32(src + 4 * k);
        }
      }
      // ---- boundary strengths: lane = (direction, edge, segment) ----
      {
        const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
        const int qx = dir ? k : e, qy = dir ? e : k, px = dir ? qx : (qx + 3) & 3, py = dir ? (qy + 3) & 3 : qy;
        int s = 0;
        if (!(e == 0 && (dir ? row : mbx) == 0)) {
          const b2dbk_mb &mp = e ? mq : (dir ? mt : ml);
          if (mp.intra || mq.intra) s = e ? 3 : 4;
          else if (((mp.cbp_blk >> (py * 4 + px)) & 1) || ((mq.cbp_blk >> (qy * 4 + qx)) & 1)) s = 2;
          else {
            const b2dbk_blk *pq = a.blks + (size_t)(row * 4 + qy) * bw + mbx * 4 + qx;
            const b2dbk_blk bq = *pq, bp = *(dir ? pq - bw : pq - 1);
            const int p0 = bp.ref[0], p1 = bp.ref[1], q0 = bq.ref[0], q1 = bq.ref[1];
            if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) s = 1;
            else if (p0 != p1) s = p0 == q0 ? (dbk_mvne(bp, 0, bq, 0) | dbk_mvne(bp, 1, bq, 1)) : (dbk_mvne(bp, 0, bq, 1) | dbk_mvne(bp, 1, bq, 0));
            else s = (dbk_mvne(bp, 0, bq, 0) | dbk_mvne(bp, 1, bq, 1)) && (dbk_mvne(bp, 0, bq, 1) | dbk_mvne(bp, 1, bq, 0));
          }
        }
        sbs[dir][e][k] = (uint8_t)s;
      }
      __syncwarp();
      // ---- the eight edge steps ----
#pragma unroll 1
      for (int dir = 0; dir < 2; dir++) {
#pragma unroll 1
        for (int e = 0; e < 4; e++) {
          const uint32_t b4 = *reinterpret_cast<const uint32_t *>(sbs[dir][e]);
          if (b4 == 0u) continue;
          const b2dbk_mb &mp = e ? mq : (dir ? mt : ml);
          if (lane < 16) {
            if (!((e & 1) && mq.transform8x8)) {
              const int bs = (b4 >> (8 * (lane >> 2))) & 0xff;
              const int qp = (mp.qp + mq.qp + 1) >> 1, ia = iclamp(qp + mq.alpha_off, 0, 51), ib = iclamp(qp + mq.beta_off, 0, 51);
              const int alpha = c_dbk_alpha[ia], beta = c_dbk_beta[ib];
              if (bs && (alpha | beta))
                dbk_luma_line(dir ? ty + (4 + 4 * e) * DBK_YP + 4 + lane : ty + (4 + lane) * DBK_YP + 4 + 4 * e, dir ? DBK_YP : 1, bs, alpha, beta, c_dbk_clip[ia][bs]);
            }
          } else if (!(e & 1)) {
            const int pl = (lane - 16) >> 3, i = lane & 7;
            const int bs = (b4 >> (8 * (i >> 1))) & 0xff;
            const int qp = ((pl ? mp.qpc_v : mp.qpc_u) + (pl ? mq.qpc_v : mq.qpc_u) + 1) >> 1;
            const int ia = iclamp(qp + mq.alpha_off, 0, 51), ib = iclamp(qp + mq.beta_off, 0, 51);
            const int alpha = c_dbk_alpha[ia], beta = c_dbk_beta[ib];
            if (bs && (alpha | beta))
              dbk_chroma_line(dir ? tc[pl] + (4 + 2 * e) * DBK_CP + 4 + i : tc[pl] + (4 + i) * DBK_CP + 4 + 2 * e, dir ? DBK_CP : 1, bs, alpha, beta, c_dbk_clip[ia][bs]);
          }
          __syncwarp();
        }
      }
      // ---- write back: own rows with the left neighbour's four columns, the top neighbour's rows without the corner ----
      if (lane < 20) {
        const int gy = row * 16 - 4 + lane;
        if (gy >= 0) {
          uint8_t *dstp = a.y + (size_t)gy * a.yp + mbx * 16 - 4;
          const uint32_t *src = reinterpret_cast<const uint32_t *>(ty + lane * DBK_YP);
#pragma unroll
          for (int k = 0; k < 5; k++) if (k || (mbx && lane >= 4)) reinterpret_cast<uint32_t *>(dstp)[k] = src[k];
        }
      }
      if (lane < 24) {
        const int pl = lane / 12, r = lane - 12 * pl, gy = row * 8 - 4 + r;
        if (gy >= 0) {
          uint8_t *dstp = (pl ? a.v : a.u) + (size_t)gy * a.cp + mbx * 8 - 4;
          const uint32_t *src = reinterpret_cast<const uint32_t *>(tc[pl] + r * DBK_CP);
#pragma unroll
          for (int k = 0; k < 3; k++) if (k || (mbx && r >= 4)) reinterpret_cast<uint32_t *>(dstp)[k] = src[k];
        }
      }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile int *>(a.progress + row) = mbx + 1;
  }
}

}  // namespace b2

using namespace b2;

static thread_local char g_dbkerr[256];
extern "C" const char *b2dbk_last_error(void) { return g_dbkerr; }

extern "C" int b2dbk_frame_dev(int W, int H, uint8_t *y, int y_pitch, uint8_t *u, uint8_t *v, int c_pitch,
                               const b2dbk_mb *mbs, const b2dbk_blk *blks, int *progress, void *stream)
{
  if (W <= 0 || H <= 0 || (W & 15) || (H & 15) || !y || !u || !v || !mbs || !blks || !progress || y_pitch < W || c_pitch < W / 2 || (y_pitch & 3) || (c_pitch & 3) ||
      ((uintptr_t)y & 3) || ((uintptr_t)u & 3) || ((uintptr_t)v & 3)) {
    snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: picture size must be a multiple of 16, planes and pitches 4-byte aligned");
    return B2ME_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0, occ = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_deblock, 32, 0);
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  if (H / 16 > occ * sms) {                         // every macroblock row must be resident: a row waits on the row above
    snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %d macroblock rows exceed the %d resident row CTAs of this device", H / 16, occ * sms);
    return B2ME_EUNSUPPORTED;
  }
  DbkArgs a;
  a.W = W; a.H = H; a.mbw = W / 16; a.mbh = H / 16; a.y = y; a.yp = y_pitch; a.u = u; a.v = v; a.cp = c_pitch; a.mbs = mbs; a.blks = blks; a.progress = progress;
  e = cudaMemsetAsync(progress, 0, sizeof(int) * a.mbh, s);
  if (e == cudaSuccess) { k_deblock<<<a.mbh, 32, 0, s>>>(a); e = cudaGetLastError(); }
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  return B2ME_OK;
}

extern "C" int b2dbk_frame(int device, int W, int H, uint8_t *y, uint8_t *u, uint8_t *v, const b2dbk_mb *mbs, const b2dbk_blk *blks)
{
  if (W <= 0 || H <= 0 || (W & 15) || (H & 15) || !y || !u || !v || !mbs || !blks) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: bad arguments"); return B2ME_EINVAL; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  const size_t ny = (size_t)W * H, nc = ny / 4, nm = sizeof(b2dbk_mb) * (W / 16) * (H / 16), nb = sizeof(b2dbk_blk) * (W / 4) * (H / 4), np = sizeof(int) * (H / 16);
  const size_t o_u = ny, o_v = o_u + nc, o_m = (o_v + nc + 15) & ~(size_t)15, o_b = (o_m + nm + 15) & ~(size_t)15, o_p = (o_b + nb + 15) & ~(size_t)15;
  uint8_t *d = nullptr;
  e = cudaMalloc(&d, o_p + np);
  if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  cudaMemcpy(d, y, ny, cudaMemcpyHostToDevice); cudaMemcpy(d + o_u, u, nc, cudaMemcpyHostToDevice); cudaMemcpy(d + o_v, v, nc, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_m, mbs, nm, cudaMemcpyHostToDevice); cudaMemcpy(d + o_b, blks, nb, cudaMemcpyHostToDevice);
  int r = b2dbk_frame_dev(W, H, d, W, d + o_u, d + o_v, W / 2, reinterpret_cast<const b2dbk_mb *>(d + o_m), reinterpret_cast<const b2dbk_blk *>(d + o_b),
                          reinterpret_cast<int *>(d + o_p), 0);
  if (!r) {
    e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(y, d, ny, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(u, d + o_u, nc, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(v, d + o_v, nc, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { snprintf(g_dbkerr, sizeof(g_dbkerr), "b2dbk_frame: %s", cudaGetErrorString(e)); r = B2ME_ECUDA; }
  }
  cudaFree(d);
  return r;
}
