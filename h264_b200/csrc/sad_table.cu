// sad_table.cu -- the SAD tables of JM's fast full search for ONE (macroblock, reference): b2me_sad_table.
//
// Replaces   setup_fast_full_search            JM/lencod/src/me_fullfast.c:269-608 (the luma, unweighted branch :470-520)
//            update_full_search_large_blocks   JM/lencod/src/me_fullfast.c:195-260
// For every position of the spiral around the search centre: the sixteen 4x4 SADs of the macroblock (the reference reads the
// block at UMVLine4X's clamped 16x16 origin; in the integer plane the pad is edge replication, so that equals a per-pixel
// coordinate clamp) and their tree sums for the 41 partitions -- BlockSAD[blocktype][block][pos] in this library's partition
// numbering, uint16 (a 16x16 SAD is at most 65280).  The 41 table scans of fast_full_search_motion_estimation (:618-689), whose
// motion-vector costs depend on predictors that only exist one partition after the other, stay with the caller
// (integration/jm/b2me_jm_shim.c): this is the split the reference itself makes between set-up and search.
// One thread per position; a drop-in service (one small launch per (macroblock, reference)), not the throughput path.
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

__global__ void __launch_bounds__(128) k_sad_table(const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *__restrict__ plane, int Wp, int Hp,
                                                   int mbx, int mby, int cx, int cy, int npos, uint16_t *__restrict__ out)
{
  __shared__ uint8_t c[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) c[i] = cur[(size_t)(mby * 16 + (i >> 4)) * cur_pitch + mbx * 16 + (i & 15)];
  __syncthreads();
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= npos) return;
  int sx, sy;
  spiral_xy(pos, &sx, &sy);
  const int x0 = mbx * 16 + cx + sx + PADX, y0 = mby * 16 + cy + sy + PADY;        // block origin in the padded plane
  int s[16];
#pragma unroll
  for (int k = 0; k < 16; k++) s[k] = 0;
#pragma unroll 4
  for (int y = 0; y < 16; y++) {
    const uint8_t *row = plane + (size_t)iclamp(y0 + y, 0, Hp - 1) * Wp;
#pragma unroll
    for (int x = 0; x < 16; x++) {
      const int d = (int)row[iclamp(x0 + x, 0, Wp - 1)] - (int)c[y * 16 + x];
      s[(y >> 2) * 4 + (x >> 2)] += d < 0 ? -d : d;
    }
  }
  auto put = [&](int p, int v) { out[(size_t)p * npos + pos] = (uint16_t)v; };
  int q[4];
#pragma unroll
  for (int k = 0; k < 4; k++) q[k] = s[(k >> 1) * 8 + (k & 1) * 2] + s[(k >> 1) * 8 + (k & 1) * 2 + 1] + s[(k >> 1) * 8 + (k & 1) * 2 + 4] + s[(k >> 1) * 8 + (k & 1) * 2 + 5];
  put(0, q[0] + q[1] + q[2] + q[3]);
  put(1, q[0] + q[1]); put(2, q[2] + q[3]); put(3, q[0] + q[2]); put(4, q[1] + q[3]);
#pragma unroll
  for (int k = 0; k < 4; k++) put(5 + k, q[k]);
#pragma unroll
  for (int k = 0; k < 8; k++) put(9 + k, s[(k >> 1) * 4 + (k & 1) * 2] + s[(k >> 1) * 4 + (k & 1) * 2 + 1]);           // 8x4: row k>>1, half k&1
#pragma unroll
  for (int k = 0; k < 8; k++) put(17 + k, s[(k >> 2) * 8 + (k & 3)] + s[(k >> 2) * 8 + 4 + (k & 3)]);               // 4x8: rows 2(k>>2), +1, column k&3
#pragma unroll
  for (int k = 0; k < 16; k++) put(25 + k, s[k]);
}

}  // namespace b2

using namespace b2;

extern "C" int b2me_sad_table(b2me_ctx *c, int mb_x, int mb_y, int ref_idx, const int16_t center_mv[2], int search_range_pel, uint16_t *table)
{
  if (!c || !center_mv || !table || mb_x < 0 || mb_y < 0 || mb_x >= c->mbw || mb_y >= c->mbh || ref_idx < 0 || ref_idx >= c->nrefs ||
      search_range_pel < 0 || search_range_pel > c->R || ((center_mv[0] | center_mv[1]) & 3)) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  const int npos = (2 * search_range_pel + 1) * (2 * search_range_pel + 1);
  const size_t bytes = (size_t)NPART * npos * sizeof(uint16_t);
  if (c->sadtab_bytes < bytes) {
    if (c->d_sadtab) cudaFree(c->d_sadtab);
    c->d_sadtab = nullptr; c->sadtab_bytes = 0;
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_sadtab, bytes));
    c->sadtab_bytes = bytes;
  }
  cudaStream_t s = c->stream;                      // behind a plane build a host-pointer b2me_set_ref left running on this stream
  k_sad_table<<<(npos + 127) / 128, 128, 0, s>>>(c->d_cur, c->W, c->d_planes + (size_t)ref_idx * 16 * c->plane_size, c->Wp, c->Hp,
                                                  mb_x, mb_y, center_mv[0] >> 2, center_mv[1] >> 2, npos, c->d_sadtab);
  B2_CUDA_CHECK(c, cudaGetLastError());
  c->launches++;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(table, c->d_sadtab, bytes, cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaStreamSynchronize(s));
  return B2ME_OK;
}
