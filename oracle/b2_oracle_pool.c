/* b2_oracle_pool.c -- TEST INFRASTRUCTURE ONLY: CPU statement of the fractal range x domain-POOL matching
 * (BASELINE.json config 5: 8x8 range blocks against a pool of 2:1-averaged 16x16 domain blocks, 8 isometries).
 *
 * PARITY UNPINNED BY THE REFERENCE: version1 has no pool / decimated-domain / isometry search (SURVEY Q-F1:
 * its full_search matches same-size blocks at integer displacements).  This file therefore DEFINES the
 * pool-mode semantics; what it takes from the reference is compute_rms's per-pair fit
 * (V1/src/compute.c:156-182, QUAN_A V1/inc/defines_enc.h:591-601, limits :19-22):
 *     alpha = (n*Srd - Sr*Sd) / det,  det = n*Sd2 - Sd^2   (0 when det == 0)
 *     a     = (int)(alpha*100), QUAN_A(a);  beta = QUAN_A((int)(Sr/n));  reject unless -2.35 <= a/100 <= 4.0
 *     rms   = sum (r - a/100*(d - Sd/n) - beta)^2      (the reference's expanded expression, :181-182)
 * evaluated here in EXACT integer arithmetic (n = 64, num = n*Srd - Sr*Sd):
 *     a = trunc(100*num/det);  aq = QUAN_A(a);  G = 200*aq*num - aq^2*det;
 *     640000 * rms = 640000*(Sr2 - 2*beta*Sr + 64*beta^2) - G        (err_num, an int64)
 * so that "best domain" is an exact integer argmax of G with a defined tie order (isometry ascending, then
 * pool index ascending: the first maximum in that scan order wins).  orc_pool_rms_double() is the reference's
 * floating-point expression for the same pair; tests check err_num/640000 against it within 1e-6 relative.
 *
 * Pool: nd domain blocks on a uniform grid of the domain plane (orc_pool_positions), each the 2x2 average
 * ((a+b+c+d+2)>>2) of a 16x16 block.  Isometries (of the 8x8 RANGE block, orc_pool_iso): 0 identity, 1 mirror
 * x, 2 mirror y, 3 rotate 180, 4 transpose, 5 rotate 90 cw, 6 rotate 90 ccw, 7 anti-transpose.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int quan_a_pool(int x)
{
  int b = x % 10, c = x / 10;
  if (b > 2 && b < 8) b = 5;
  else if (b > 7) { b = 0; c += 1; }
  else b = 0;
  return c * 10 + b;
}

int orc_pool_quan_a(int x) { return quan_a_pool(x); }

/* pixel (i row, j col) of isometry `iso` of block src[8][8] */
static int iso_src_index(int iso, int i, int j)
{
  switch (iso) {
    case 0: return i * 8 + j;
    case 1: return i * 8 + (7 - j);
    case 2: return (7 - i) * 8 + j;
    case 3: return (7 - i) * 8 + (7 - j);
    case 4: return j * 8 + i;
    case 5: return (7 - j) * 8 + i;
    case 6: return j * 8 + (7 - i);
    default: return (7 - j) * 8 + (7 - i);
  }
}
void orc_pool_iso(const uint8_t *src, int iso, uint8_t *dst)
{
  int i, j;
  for (i = 0; i < 8; i++) for (j = 0; j < 8; j++) dst[i * 8 + j] = src[iso_src_index(iso, i, j)];
}

/* grid of nd top-left corners of 16x16 domain blocks in a dw x dh plane: nx columns x ny rows, raster order */
void orc_pool_positions(int dw, int dh, int nd, int32_t *xy)
{
  int nx = 1, ny, p;
  while ((int64_t)nx * nx * (dh - 15) < (int64_t)nd * (dw - 15)) nx++;
  ny = (nd + nx - 1) / nx;
  for (p = 0; p < nd; p++) {
    int ix = p % nx, iy = p / nx;
    xy[2 * p]     = nx > 1 ? (int)((int64_t)ix * (dw - 16) / (nx - 1)) : 0;
    xy[2 * p + 1] = ny > 1 ? (int)((int64_t)iy * (dh - 16) / (ny - 1)) : 0;
  }
}

void orc_pool_domain_block(const uint8_t *plane, int stride, int x, int y, uint8_t *blk)
{
  int i, j;
  for (i = 0; i < 8; i++)
    for (j = 0; j < 8; j++) {
      const uint8_t *p = plane + (size_t)(y + 2 * i) * stride + x + 2 * j;
      blk[i * 8 + j] = (uint8_t)((p[0] + p[1] + p[stride] + p[stride + 1] + 2) >> 2);
    }
}

/* the reference's floating-point expression for one (range, domain) pair, compute_rms :156-182 */
double orc_pool_rms_double(const uint8_t *r, const uint8_t *d, double *alpha_out, double *beta_out)
{
  double no = 64.0, rsum1 = 0, rsum2 = 0, dsum1 = 0, dsum2 = 0, rdsum = 0, det, alpha, beta, rms = 1e30;
  int k, a;
  for (k = 0; k < 64; k++) { rsum1 += r[k]; rsum2 += r[k] * r[k]; dsum1 += d[k]; dsum2 += d[k] * d[k]; rdsum += r[k] * d[k]; }
  det = no * dsum2 - dsum1 * dsum1;
  if (det == 0.0) alpha = 0.0; else alpha = (no * rdsum - rsum1 * dsum1) / det;
  a = (int)(alpha * 100);
  beta = rsum1 / no;
  a = quan_a_pool(a);
  beta = quan_a_pool((int)beta);
  alpha = (double)a / 100;
  *alpha_out = alpha; *beta_out = beta;
  if (alpha < -2.35 || alpha > 4.0) return rms;
  if (beta < -60 || beta > 255) return rms;
  rms = rsum2 + alpha * (alpha * dsum2 - 2.0 * rdsum + 2.0 * (beta - alpha * dsum1 / no) * dsum1)
      + (beta - alpha * dsum1 / no) * ((beta - alpha * dsum1 / no) * no - 2.0 * rsum1);
  return rms;
}

/* exact integer fit of one pair; returns 0 when rejected */
int orc_pool_pair(const uint8_t *r, const uint8_t *d, int *aq_out, int64_t *G_out)
{
  int64_t Sr = 0, Sd = 0, Sd2 = 0, Srd = 0, num, det, a;
  int k, aq;
  for (k = 0; k < 64; k++) { Sr += r[k]; Sd += d[k]; Sd2 += d[k] * d[k]; Srd += r[k] * d[k]; }
  num = 64 * Srd - Sr * Sd;
  det = 64 * Sd2 - Sd * Sd;
  a = det == 0 ? 0 : (100 * num) / det;          /* C division truncates toward zero, like the (int) cast */
  aq = quan_a_pool((int)a);
  *aq_out = aq;
  if (aq < -235 || aq > 400) return 0;
  *G_out = 200 * (int64_t)aq * num - (int64_t)aq * aq * det;
  return 1;
}

/* Search of every 8x8 range block (raster order) of the range plane against the pool.
 * out: best_dom[nr] (-1: every pair rejected), best_iso[nr], aq[nr], beta[nr], err_num[nr] (= 640000*rms). */
void orc_pool_search(const uint8_t *rplane, int rw, int rh, int rstride, const uint8_t *dplane, int dw, int dh, int dstride, int nd,
                     int32_t *best_dom, uint8_t *best_iso, int16_t *aq_out, int16_t *beta_out, int64_t *err_num)
{
  uint8_t *pool = (uint8_t *)malloc((size_t)nd * 64);
  int32_t *xy = (int32_t *)malloc((size_t)nd * 2 * sizeof(int32_t));
  int p, bx, by, iso, k, i, j;
  orc_pool_positions(dw, dh, nd, xy);
  for (p = 0; p < nd; p++) orc_pool_domain_block(dplane, dstride, xy[2 * p], xy[2 * p + 1], pool + (size_t)p * 64);
  for (by = 0; by < rh / 8; by++)
    for (bx = 0; bx < rw / 8; bx++) {
      uint8_t r0[64], r[64];
      int64_t bestG = -1, Sr = 0, Sr2 = 0; int bd = -1, bi = 0, ba = 0, beta;
      const int ri = by * (rw / 8) + bx;
      for (i = 0; i < 8; i++) for (j = 0; j < 8; j++) r0[i * 8 + j] = rplane[(size_t)(by * 8 + i) * rstride + bx * 8 + j];
      for (k = 0; k < 64; k++) { Sr += r0[k]; Sr2 += r0[k] * r0[k]; }
      beta = quan_a_pool((int)(Sr / 64));
      for (iso = 0; iso < 8; iso++) {
        orc_pool_iso(r0, iso, r);
        for (p = 0; p < nd; p++) {
          int aq; int64_t G;
          if (!orc_pool_pair(r, pool + (size_t)p * 64, &aq, &G)) continue;
          if (G > bestG) { bestG = G; bd = p; bi = iso; ba = aq; }
        }
      }
      best_dom[ri] = bd; best_iso[ri] = (uint8_t)bi; aq_out[ri] = (int16_t)ba; beta_out[ri] = (int16_t)beta;
      err_num[ri] = bd < 0 ? -1 : 640000 * (Sr2 - 2 * (int64_t)beta * Sr + 64 * (int64_t)beta * beta) - bestG;
    }
  free(pool); free(xy);
}
