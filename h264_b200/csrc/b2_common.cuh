// b2_common.cuh -- shared device/host helpers for libb2me (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace b2 {

constexpr int PADX = 32;   // IMG_PAD_SIZE_X  JM/lencod/inc/defines.h:120
constexpr int PADY = 20;   // IMG_PAD_SIZE_Y  JM/lencod/inc/defines.h:121
constexpr int NPART = 41;

__host__ __device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Partition geometry tables (p = 0..40), see include/b2me.h.
struct PartGeom { uint8_t bt, ox, oy, w, h; };
__host__ __device__ __forceinline__ PartGeom part_geom(int p)
{
  PartGeom g;
  if (p == 0)      { g.bt = 1; g.w = 16; g.h = 16; g.ox = 0; g.oy = 0; }
  else if (p < 3)  { g.bt = 2; g.w = 16; g.h = 8;  g.ox = 0; g.oy = (uint8_t)((p - 1) * 8); }
  else if (p < 5)  { g.bt = 3; g.w = 8;  g.h = 16; g.ox = (uint8_t)((p - 3) * 8); g.oy = 0; }
  else if (p < 9)  { int k = p - 5;  g.bt = 4; g.w = 8; g.h = 8; g.ox = (uint8_t)((k & 1) * 8); g.oy = (uint8_t)((k >> 1) * 8); }
  else if (p < 17) { int k = p - 9;  g.bt = 5; g.w = 8; g.h = 4; g.ox = (uint8_t)((k & 1) * 8); g.oy = (uint8_t)((k >> 1) * 4); }
  else if (p < 25) { int k = p - 17; g.bt = 6; g.w = 4; g.h = 8; g.ox = (uint8_t)((k & 3) * 4); g.oy = (uint8_t)((k >> 2) * 8); }
  else             { int k = p - 25; g.bt = 7; g.w = 4; g.h = 4; g.ox = (uint8_t)((k & 3) * 4); g.oy = (uint8_t)((k >> 2) * 4); }
  return g;
}
// first partition index of blocktype bt (1..7)
__host__ __device__ __forceinline__ int part_first(int bt)
{ return bt == 1 ? 0 : bt == 2 ? 1 : bt == 3 ? 3 : bt == 4 ? 5 : bt == 5 ? 9 : bt == 6 ? 17 : 25; }

// get_search_range (JM/lencod/src/mv_search.c:70-92): per-block search range in pel.
__host__ __device__ __forceinline__ int block_search_range(int R, int mode, int ref, int bt)
{
  int q = R << 2, scale = 1;
  if (mode == 1) scale = (ref < 1 ? ref : 1) + 1;
  else if (mode != 2) scale = ((ref < 1 ? ref : 1) + 1) * (bt < 2 ? bt : 2);
  return (q / scale) >> 2;
}

// mvbits[d] = 1 for d == 0 else 2*floor(log2|d|)+3   (JM/lencod/src/mv_search.c:366-374)
__device__ __forceinline__ int mvbits(int d)
{
  int a = d < 0 ? -d : d;
  return a ? (65 - 2 * __clz(a)) : 1;
}

// Index of integer displacement (x,y) in the JM spiral (mv_search.c:406-442).
__host__ __device__ __forceinline__ int spiral_index(int x, int y)
{
  int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
  int l = ax > ay ? ax : ay;
  if (l == 0) return 0;
  int base = (2 * l - 1) * (2 * l - 1);
  if (ay == l && ax < l) return base + 2 * (x + l - 1) + (y > 0 ? 1 : 0);
  return base + 2 * (2 * l - 1) + 2 * (y + l) + (x > 0 ? 1 : 0);
}
// inverse of spiral_index
__host__ __device__ __forceinline__ void spiral_xy(int pos, int *x, int *y)
{
  if (pos == 0) { *x = 0; *y = 0; return; }
  int l = (int)((sqrtf((float)pos) + 1.0f) * 0.5f);          // ring: (2l-1)^2 <= pos < (2l+1)^2
  while ((2 * l + 1) * (2 * l + 1) <= pos) l++;
  while ((2 * l - 1) * (2 * l - 1) > pos) l--;
  int k = pos - (2 * l - 1) * (2 * l - 1);
  if (k < 2 * (2 * l - 1)) { *x = (k >> 1) - l + 1; *y = (k & 1) ? l : -l; }
  else { k -= 2 * (2 * l - 1); *y = (k >> 1) - l; *x = (k & 1) ? l : -l; }
}

// acc + sum_i |a.b[i] - b.b[i]| on four packed unsigned bytes: one VABSDIFF4.U8.ACC.
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc)
{
  uint32_t r;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
  return r;
}

}  // namespace b2

#define B2_CUDA_CHECK(ctx, expr)                                                          \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #expr, \
               cudaGetErrorString(_e));                                                   \
      return B2ME_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)
