// subpel_planes.cu -- the 16 quarter-pel luma planes of one reference picture.
//
// Replaces getSubImagesLuma (JM/lencod/src/img_luma.c:611-680) and its helpers
// (getSubImageInteger :40, getHorSubImageSixTap :151, getVerSubImageSixTap :257,
//  getVerSubImageSixTapTmp :347, get*SubImageBiLinear :440-600), bit-exactly.
//
// Layout in HBM: planes16[yy*4+xx][Hp][Wp] u8, Hp = H+40, Wp = W+64 (JM's own padding:
// IMG_PAD_SIZE_Y/X), so plane [y&3][x&3] row (y>>2)+20, column (x>>2)+32 is what
// UMVLine4X (refbuf.h:22-26) addresses.
//
// The reference filters the *padded* plane and clamps tap indices to the padded row/column
// ends; because the pad is pure edge replication this equals filtering the unpadded luma with
// every tap coordinate clamped to [0,W-1]x[0,H-1], which is what kernel 1 does (no halo
// exchange between thread blocks is needed).  Kernel 2 forms the 12 bilinear planes.
//
// Roofline: HBM-bound. Algorithmic bytes per reference = W*H read + 16*Wp*Hp written.
#include "b2_common.cuh"
#include "b2_ctx.h"

namespace b2 {

constexpr int TX = 32, TY = 8;   // output tile per CTA (kernel 1)

// Kernel 1: planes [0][0], [0][2], [2][0], [2][2].
__global__ void __launch_bounds__(TX * TY) k_half_planes(const uint8_t *__restrict__ luma, int pitch, int W, int H,
                                                          uint8_t *__restrict__ planes, int Wp, int Hp, int ybase, int yend)
{
  // tile of clamped luma: rows y-2..y+3, cols x-2..x+3 around each output sample
  __shared__ uint8_t sl[TY + 5][TX + 5 + 3];
  __shared__ int sh[TY + 5][TX];           // unrounded horizontal 6-tap for rows y-2..y+3
  const int x0 = blockIdx.x * TX, y0 = ybase + blockIdx.y * TY;      // plane rows [ybase, yend): MB-row bands build only what they read
  const int tid = threadIdx.y * TX + threadIdx.x;
  for (int i = tid; i < (TY + 5) * (TX + 5); i += TX * TY) {
    int r = i / (TX + 5), c = i % (TX + 5);
    // padded-plane coordinates (x0+c-2, y0+r-2) -> luma coordinates, clamped.  The reference
    // clamps first to the padded plane and then the pad replicates the picture edge; the
    // composition is a single clamp to the picture.
    int lx = iclamp(x0 + c - 2 - PADX, 0, W - 1), ly = iclamp(iclamp(y0 + r - 2, 0, Hp - 1) - PADY, 0, H - 1);
    // horizontal clamp to the padded row ends happens before the picture clamp; identical result
    sl[r][c] = luma[(size_t)ly * pitch + lx];
  }
  __syncthreads();
  for (int i = tid; i < (TY + 5) * TX; i += TX * TY) {
    int r = i / TX, c = i % TX;
    const uint8_t *s = &sl[r][c];        // s[2] is the sample at column x0+c
    sh[r][c] = 20 * (s[2] + s[3]) - 5 * (s[1] + s[4]) + (s[0] + s[5]);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= Wp || y >= yend) return;
  const int r = threadIdx.y + 2, c = threadIdx.x;
  const size_t PS = (size_t)Wp * Hp, o = (size_t)y * Wp + x;
  int p00 = sl[r][c + 2];
  int ih = sh[r][c];
  int iv = 20 * (sl[r][c + 2] + sl[r + 1][c + 2]) - 5 * (sl[r - 1][c + 2] + sl[r + 2][c + 2]) + (sl[r - 2][c + 2] + sl[r + 3][c + 2]);
  int id = 20 * (sh[r][c] + sh[r + 1][c]) - 5 * (sh[r - 1][c] + sh[r + 2][c]) + (sh[r - 2][c] + sh[r + 3][c]);
  planes[0 * PS + o]  = (uint8_t)p00;
  planes[2 * PS + o]  = (uint8_t)iclamp((ih + 16) >> 5, 0, 255);
  planes[8 * PS + o]  = (uint8_t)iclamp((iv + 16) >> 5, 0, 255);
  planes[10 * PS + o] = (uint8_t)iclamp((id + 512) >> 10, 0, 255);
}

// Kernel 2: the twelve quarter planes, (a+b+1)>>1 of two of the four planes above
// (img_luma.c:653-678).  One thread per 4 horizontally adjacent samples.
__global__ void __launch_bounds__(256) k_quarter_planes(uint8_t *__restrict__ planes, int Wp, int Hp, int ybase)
{
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = ybase + blockIdx.y;
  if (x >= Wp) return;
  const size_t PS = (size_t)Wp * Hp;
  const size_t r = (size_t)y * Wp, rn = (size_t)min(y + 1, Hp - 1) * Wp;
  const uint8_t *p00 = planes, *p02 = planes + 2 * PS, *p20 = planes + 8 * PS, *p22 = planes + 10 * PS;
  uint32_t o[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int xx = x + k, xn = min(xx + 1, Wp - 1);
    const int a00 = p00[r + xx], a02 = p02[r + xx], a20 = p20[r + xx], a22 = p22[r + xx];
    const int b00 = p00[r + xn], b20 = p20[r + xn];       // x+1
    const int c00 = p00[rn + xx], c02 = p02[rn + xx];      // y+1
    const int sft = 8 * k;
    o[0]  |= (uint32_t)((a00 + a02 + 1) >> 1) << sft;   // [0][1]
    o[1]  |= (uint32_t)((a02 + b00 + 1) >> 1) << sft;   // [0][3]
    o[2]  |= (uint32_t)((a00 + a20 + 1) >> 1) << sft;   // [1][0]
    o[3]  |= (uint32_t)((a02 + a20 + 1) >> 1) << sft;   // [1][1]
    o[4]  |= (uint32_t)((a02 + a22 + 1) >> 1) << sft;   // [1][2]
    o[5]  |= (uint32_t)((a02 + b20 + 1) >> 1) << sft;   // [1][3]
    o[6]  |= (uint32_t)((a20 + a22 + 1) >> 1) << sft;   // [2][1]
    o[7]  |= (uint32_t)((a22 + b20 + 1) >> 1) << sft;   // [2][3]
    o[8]  |= (uint32_t)((a20 + c00 + 1) >> 1) << sft;   // [3][0]
    o[9]  |= (uint32_t)((a20 + c02 + 1) >> 1) << sft;   // [3][1]
    o[10] |= (uint32_t)((a22 + c02 + 1) >> 1) << sft;   // [3][2]
    o[11] |= (uint32_t)((c02 + b20 + 1) >> 1) << sft;   // [3][3]
  }
  const int idx[12] = {1, 3, 4, 5, 6, 7, 9, 11, 12, 13, 14, 15};
#pragma unroll
  for (int i = 0; i < 12; i++)
    *reinterpret_cast<uint32_t *>(planes + idx[i] * PS + r + x) = o[i];   // Wp % 4 == 0
}

// Search planes of the integer full search: the reconstructed luma with an edge-replicated pad of `spad`
// on every side (the [0][0] plane of getSubImagesLuma continued further out; UMVLine4X's clamp of the block
// origin, refbuf.h:25, equals this replication, see sad_fs.cu), stored 16 times, plane s shifted left by s
// bytes: out[s][y][x] = P[y][x+s].  TMA boxes must start on 16-byte boundaries; the shifted planes make
// every byte column reachable.  HBM-bound: W*H read, 16*Wq*Hq written.
__global__ void __launch_bounds__(256) k_search_plane(const uint8_t *__restrict__ luma, int pitch, int W, int H,
                                                       uint8_t *__restrict__ out, int Wq, int Hq, int spad, int ybase, int yend)
{
  const int wq4 = Wq >> 2;
  const int i = ybase * wq4 + blockIdx.x * blockDim.x + threadIdx.x;      // plane rows [ybase, yend)
  if (i >= wq4 * yend) return;
  const int y = i / wq4, x = (i - y * wq4) * 4;
  const uint8_t *row = luma + (size_t)iclamp(y - spad, 0, H - 1) * pitch;
  uint32_t b[19];
#pragma unroll
  for (int k = 0; k < 19; k++) b[k] = row[iclamp(x + k - spad, 0, W - 1)];
#pragma unroll
  for (int s = 0; s < 16; s++)
    reinterpret_cast<uint32_t *>(out + (size_t)s * Wq * Hq)[i] = b[s] | (b[s + 1] << 8) | (b[s + 2] << 16) | (b[s + 3] << 24);
}

// Explicit weighted prediction of the single-list search (computeSADWP / SATDWP / SSEWP, me_distortion.c:434-517,
// 833-935, 1262-1345): the distortion replaces every reference sample v it reads by
// clip1(((weight*v + round) >> log_denom) + offset).  The mapping is pointwise on the fetched sample, so it is
// applied once to the reference's planes (16 quarter-pel planes and 16 search planes) instead of per candidate.
__global__ void __launch_bounds__(256) k_apply_wp(uint32_t *__restrict__ buf, size_t nwords, int weight, int offset, int log_denom, int rnd)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t w = buf[i];
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int v = (int)((w >> (8 * k)) & 255u);
      o |= (uint32_t)iclamp(((weight * v + rnd) >> log_denom) + offset, 0, 255) << (8 * k);
    }
    buf[i] = o;
  }
}
cudaError_t launch_apply_wp(uint8_t *buf, size_t bytes, int weight, int offset, int log_denom, cudaStream_t s)
{
  const size_t nw = bytes / 4;
  const int grid = (int)((nw + 255) / 256 < 148 * 16 ? (nw + 255) / 256 : 148 * 16);
  k_apply_wp<<<grid, 256, 0, s>>>(reinterpret_cast<uint32_t *>(buf), nw, weight, offset, log_denom, log_denom ? 1 << (log_denom - 1) : 0);
  return cudaGetLastError();
}

// row_lo / row_hi: luma rows [row_lo, row_hi) whose planes are (re)built; the pad above / below goes with the first /
// last picture row.  The whole picture: 0, H.
cudaError_t launch_search_plane(const uint8_t *luma, int pitch, int W, int H, uint8_t *out, int Wq, int Hq, int spad, int row_lo, int row_hi, cudaStream_t s)
{
  const int y0 = row_lo <= 0 ? 0 : row_lo + spad, y1 = row_hi >= H ? Hq : row_hi + spad;
  const int n = (Wq >> 2) * (y1 - y0);
  if (n <= 0) return cudaSuccess;
  k_search_plane<<<(n + 255) / 256, 256, 0, s>>>(luma, pitch, W, H, out, Wq, Hq, spad, y0, y1);
  return cudaGetLastError();
}

cudaError_t launch_subpel_planes(const uint8_t *luma, int pitch, int W, int H, uint8_t *planes16, int row_lo, int row_hi, cudaStream_t s)
{
  const int Wp = W + 2 * PADX, Hp = H + 2 * PADY;
  const int y0 = row_lo <= 0 ? 0 : row_lo + PADY, y1 = row_hi >= H ? Hp : row_hi + PADY;
  if (y1 <= y0) return cudaSuccess;
  dim3 g1((Wp + TX - 1) / TX, (y1 - y0 + TY - 1) / TY), b1(TX, TY);
  k_half_planes<<<g1, b1, 0, s>>>(luma, pitch, W, H, planes16, Wp, Hp, y0, y1);
  dim3 g2((Wp / 4 + 255) / 256, y1 - y0);
  k_quarter_planes<<<g2, 256, 0, s>>>(planes16, Wp, Hp, y0);
  return cudaGetLastError();
}

}  // namespace b2
