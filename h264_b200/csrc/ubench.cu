// ubench.cu -- instruction-mix micro-benchmarks that give the integer roofline denominator:
// the measured chip-wide issue rate of VABSDIFF4.U8.ACC (the packed-byte SAD instruction the
// search kernel is built on), alone and co-issued with the other instruction classes the
// kernel needs.  MEASURED_PEAKS.json has no integer figure, so bench.py
// measures it live with this routine.
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

template <int KIND>
__global__ void __launch_bounds__(256) k_ubench(uint32_t *out, int iters, uint32_t seed)
{
  uint32_t a[8], x[8], y[8];
  const uint32_t b = seed * 0x01010101u + threadIdx.x;
  __shared__ uint32_t sm[256];
  sm[threadIdx.x] = b;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; x[i] = b ^ (i * 0x9e3779b9u); y[i] = x[i] + 7; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (KIND == 0 || KIND == 1 || KIND == 2 || KIND == 3 || KIND == 7 || (KIND >= 8 && KIND != 19 && KIND != 24 && KIND != 25)) a[i] = sad4(x[i], b, a[i]);
        // 24: VABSDIFF4 predicated OFF at run time (does a nullified instruction hold the ALU pipe?); 25: one live + one nullified
        if (KIND == 24 || KIND == 25) asm volatile("{.reg .pred p; setp.eq.u32 p, %3, 0x7fffffff; @p vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;}" : "+r"(y[i]) : "r"(x[i]), "r"(b), "r"(iters));
        if (KIND == 25) a[i] = sad4(x[i], b, a[i]);
        if (KIND == 8) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(8));          // SHF (funnel shift)
        if (KIND == 9) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(y[i]) : "r"(b));                      // PRMT
        if (KIND == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(b), "r"(seed));        // LOP3, independent chain
        if (KIND == 11) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(y[i]) : "r"(b + i));                        // VIMNMX.U16x2
        if (KIND == 12) y[i] = sm[(threadIdx.x + i * 32 + it) & 255];                                            // LDS.32, independent
        if (KIND == 13) asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));                               // IADD, independent chain
        if (KIND == 14) y[i] = y[i] * 5u + seed;                                                                 // IMAD, independent chain
        if (KIND == 15 || KIND == 19) y[i] = __viaddmin_s16x2(x[i], b, y[i]);                                     // VIADDMNMX.S16x2
        if (KIND == 16) y[i] = __dp2a_lo(y[i], 0x0101u, b);                                                      // IDP.2A
        if (KIND == 17) { uint2 v = *reinterpret_cast<const uint2 *>(&sm[(2 * (threadIdx.x + i * 32 + it)) & 254]); y[i] = v.x ^ v.y; }   // LDS.64 (+LOP)
        if (KIND == 18) y[i] = x[i] * 65536u + y[i];                                                             // IMAD (pack)
        if (KIND == 20) y[i] = __vimin3_s16x2(x[i], b + i, y[i]);                                                // VIMNMX3.S16x2
        if (KIND == 21) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(y[i]) : "r"(b));                             // VIADD.16x2
        if (KIND == 22) asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(y[i]) : "r"(b));                         // add on the FMA pipe?
        if (KIND == 23) asm volatile("min.s16x2 %0, %0, %1;" : "+r"(y[i]) : "r"(b + i));                         // VIMNMX.S16x2
        if (KIND == 1 || KIND == 5) x[i] = x[i] * 3u + b;                                   // IMAD
        if (KIND == 2 || KIND == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b)); // IADD3
        if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(a[(i + 1) & 7]));
        if (KIND == 6) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(b + i));
        if (KIND == 7 && (i & 3) == 0) x[i] = sm[(threadIdx.x + x[i]) & 255];
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += a[i] ^ x[i] ^ y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

cudaError_t ubench(int kind, int iters, double *gops)
{
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, block = 256;
  uint32_t *out = nullptr;
  cudaError_t e = cudaMalloc(&out, (size_t)grid * block * 4);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    switch (kind) {
      case 0: k_ubench<0><<<grid, block>>>(out, iters, rep); break;
      case 1: k_ubench<1><<<grid, block>>>(out, iters, rep); break;
      case 2: k_ubench<2><<<grid, block>>>(out, iters, rep); break;
      case 3: k_ubench<3><<<grid, block>>>(out, iters, rep); break;
      case 4: k_ubench<4><<<grid, block>>>(out, iters, rep); break;
      case 5: k_ubench<5><<<grid, block>>>(out, iters, rep); break;
      case 6: k_ubench<6><<<grid, block>>>(out, iters, rep); break;
      case 7: k_ubench<7><<<grid, block>>>(out, iters, rep); break;
      case 8: k_ubench<8><<<grid, block>>>(out, iters, rep); break;
      case 9: k_ubench<9><<<grid, block>>>(out, iters, rep); break;
      case 10: k_ubench<10><<<grid, block>>>(out, iters, rep); break;
      case 11: k_ubench<11><<<grid, block>>>(out, iters, rep); break;
      case 12: k_ubench<12><<<grid, block>>>(out, iters, rep); break;
      case 13: k_ubench<13><<<grid, block>>>(out, iters, rep); break;
      case 14: k_ubench<14><<<grid, block>>>(out, iters, rep); break;
      case 15: k_ubench<15><<<grid, block>>>(out, iters, rep); break;
      case 16: k_ubench<16><<<grid, block>>>(out, iters, rep); break;
      case 17: k_ubench<17><<<grid, block>>>(out, iters, rep); break;
      case 18: k_ubench<18><<<grid, block>>>(out, iters, rep); break;
      case 19: k_ubench<19><<<grid, block>>>(out, iters, rep); break;
      case 20: k_ubench<20><<<grid, block>>>(out, iters, rep); break;
      case 21: k_ubench<21><<<grid, block>>>(out, iters, rep); break;
      case 22: k_ubench<22><<<grid, block>>>(out, iters, rep); break;
      case 23: k_ubench<23><<<grid, block>>>(out, iters, rep); break;
      case 24: k_ubench<24><<<grid, block>>>(out, iters, rep); break;
      default: k_ubench<25><<<grid, block>>>(out, iters, rep); break;
    }
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess) return e;
  const double ops = (double)grid * block * (double)iters * 32.0;   // first-op count per thread = iters*4*8
  *gops = ops / (best * 1e-3) / 1e9;
  return cudaGetLastError();
}

}  // namespace b2
