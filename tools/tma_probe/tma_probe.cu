// Stand-alone probe of cp.async.bulk.tensor with u8 elements (development aid for sad_fs.cu).
// usage: tma_probe boxw boxh x y rank
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const CUtensorMap *tm, uint8_t *out, int bytes, int x, int y, int z, int rank)
{
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv)
{
  int bw = atoi(argv[1]), bh = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), rank = atoi(argv[5]);
  const int W = 272, H = 240, N = 2;
  uint8_t *d, *o; cudaMalloc(&d, W * H * N); cudaMalloc(&o, 65536);
  uint8_t *h = (uint8_t *)malloc(W * H * N);
  for (int i = 0; i < W * H * N; i++) h[i] = (uint8_t)((i % W) * 3 + (i / W) * 7);
  cudaMemcpy(d, h, W * H * N, cudaMemcpyHostToDevice);
  CUtensorMap tm; CUtensorMap *dtm; cudaMalloc(&dtm, sizeof(tm));
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N}; cuuint64_t str[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode=%d ", (int)r);
  cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<<<1, 128, 65536>>>(dtm, o, bw * bh, x, y, 1, rank);
  cudaError_t e = cudaDeviceSynchronize();
  printf("box %dx%d at (%d,%d) rank %d: %s", bw, bh, x, y, rank, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    uint8_t *ho = (uint8_t *)malloc(bw * bh); cudaMemcpy(ho, o, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) {
      int gx = x + i, gy = y + j; uint8_t exp = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[(rank == 3 ? W * H : 0) + gy * W + gx];
      bad += ho[j * bw + i] != exp;
    }
    printf("  mismatches %d", bad);
  }
  printf("\n");
  return 0;
}
