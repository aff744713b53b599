/* b2_oracle_v1.c -- TEST INFRASTRUCTURE ONLY: CPU restatement of version1's fractal block search.
 * (V1/ = 2.论文程序/ZhangLing_Yu_version1/H264Fractal under the reference root.)
 *
 * Restates, function by function:
 *   orc_v1_box_table   <- compute_domain_Sum   V1/src/compute.c:277-684 (sliding sums; 0 where unset: calloc)
 *   orc_v1_range_table <- compute_range_Sum    V1/src/compute.c:686-1091
 *   orc_v1_rdsum       <- compute_rdSum        V1/src/compute.c:192-215
 *   orc_v1_rms         <- compute_rms          V1/src/compute.c:6-189 (QUAN_A: V1/inc/defines_enc.h:591-601,
 *                                              limits :19-22)
 *   orc_v1_bound_chk   <- bound_chk            V1/src/block_enc.c:2894-2919
 *   orc_v1_full_search <- full_search          V1/src/block_enc.c:1933-1977
 * Pinned against the unmodified reference (oracle/_ref/libv1ref.so) by tests/test_oracle_v1.py and
 * against tests/golden/v1_harness_cif.npz (generated from libv1ref.so by oracle/gen_golden_v1.py).
 * Compile with plain -O2 on x86-64: doubles are SSE2, no contraction (liborc is built that way).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int w, h, R;
  const uint8_t *org, *ref;     /* range plane, domain plane (w x h) */
  int have_sums;                /* 0: the domain sum tables were never built (all zero, SURVEY Q-F3) */
} orc_v1_plane;

/* sliding box sum of block bw x bh at (x,y); squares: sum of squares */
static double box(const orc_v1_plane *p, const uint8_t *img, int x, int y, int bw, int bh, int squares)
{
  double s = 0.0;
  int i, j;
  for (i = 0; i < bh; i++)
    for (j = 0; j < bw; j++) {
      int v = img[(size_t)(y + i) * p->w + x + j];
      s += squares ? v * v : v;
    }
  return s;
}

void orc_v1_box_table(const uint8_t *img, int w, int h, int bw, int bh, int squares, double *out)
{
  orc_v1_plane p; int x, y;
  p.w = w; p.h = h;
  memset(out, 0, (size_t)w * h * sizeof(double));
  for (y = 0; y + bh <= h; y++)
    for (x = 0; x + bw <= w; x++) out[(size_t)y * w + x] = box(&p, img, x, y, bw, bh, squares);
}

void orc_v1_range_table(const uint8_t *img, int w, int h, int squares, double *out)
{
  orc_v1_plane p; int gx, gy;
  p.w = w; p.h = h;
  for (gy = 0; gy < h / 4; gy++)
    for (gx = 0; gx < w / 4; gx++) out[(size_t)gy * (w / 4) + gx] = box(&p, img, gx * 4, gy * 4, 4, 4, squares);
}

double orc_v1_rdsum(const orc_v1_plane *p, int bx, int by, int m, int n, int bw, int bh)
{
  double rdsum = 0.0;
  int i, j;
  for (i = 0; i < bh; i++)
    for (j = 0; j < bw; j++) rdsum += p->org[(size_t)(by + i) * p->w + bx + j] * p->ref[(size_t)(n + i) * p->w + m + j];
  return rdsum;
}

static int quan_a(int x)
{
  int b = x % 10, c = x / 10;
  if (b > 2 && b < 8) b = 5;
  else if (b > 7) { b = 0; c += 1; }
  else b = 0;
  return c * 10 + b;
}

double orc_v1_rms(const orc_v1_plane *p, int bx, int by, int m, int n, int bw, int bh, double *alpha, double *beta)
{
  double rms = 1e30, det;
  int a, no = bw * bh;
  double dsum1 = 0, dsum2 = 0, rsum1, rsum2, rdsum;
  if (p->have_sums) { dsum1 = box(p, p->ref, m, n, bw, bh, 0); dsum2 = box(p, p->ref, m, n, bw, bh, 1); }
  rsum1 = box(p, p->org, bx, by, bw, bh, 0); rsum2 = box(p, p->org, bx, by, bw, bh, 1);
  rdsum = orc_v1_rdsum(p, bx, by, m, n, bw, bh);
  det = no * dsum2 - dsum1 * dsum1;
  if (det == 0.0) *alpha = 0.0;
  else *alpha = (no * rdsum - rsum1 * dsum1) / det;
  a = (int)(*alpha * 100);
  *beta = rsum1 / no;
  a = quan_a(a);
  *beta = quan_a((int)(*beta));
  *alpha = (double)(a) / 100;
  if (*alpha < -2.35 || *alpha > 4.0) return rms;
  if (*beta < -60 || *beta > 255) return rms;
  rms = rsum2 + (*alpha) * ((*alpha) * dsum2 - 2.0 * rdsum + 2.0 * ((*beta) - (*alpha) * dsum1 / no) * dsum1)
      + ((*beta) - (*alpha) * dsum1 / no) * (((*beta) - (*alpha) * dsum1 / no) * no - 2.0 * rsum1);
  return rms;
}

int orc_v1_bound_chk(const orc_v1_plane *p, int m, int n, int cx, int cy, int bw, int bh)
{
  int ilow = cx - p->R, ihigh = cx + p->R, jlow = cy - p->R, jhigh = cy + p->R;
  if (ilow < 0) ilow = 0;
  if (ihigh > p->w - bw) ihigh = p->w - bw;
  if (jlow < 0) jlow = 0;
  if (jhigh > p->h - bh) jhigh = p->h - bh;
  return m <= ihigh && m >= ilow && n <= jhigh && n >= jlow;
}

double orc_v1_full_search(const orc_v1_plane *p, int bx, int by, int bw, int bh, int *xy, double *so)
{
  double best, rms, alpha, beta;
  int l, k, i, j;
  best = orc_v1_rms(p, bx, by, bx, by, bw, bh, &alpha, &beta);
  so[0] = alpha; so[1] = beta;
  for (l = 1; l <= p->R; l++) {
    i = j = -l;
    for (k = 0; k < 8 * l; k++) {
      int m = bx + i, n = by + j;
      if (orc_v1_bound_chk(p, m, n, bx, by, bw, bh)) {
        rms = orc_v1_rms(p, bx, by, m, n, bw, bh, &alpha, &beta);
        if (rms < best) { best = rms; xy[0] = m - bx; xy[1] = n - by; so[0] = alpha; so[1] = beta; }
      }
      if (k < 2 * l) i++;
      else if (k < 4 * l) j++;
      else if (k < 6 * l) i--;
      else j--;
    }
  }
  return best;
}

/* every range block of one plane in the 41-partition numbering of include/b2me.h:
 * xy [nmb][41][2], so [nmb][41][2], rms [nmb][41]; macroblock grid mbw x mbh */
void orc_partition_geometry(int p, int *bt, int *ox, int *oy, int *w, int *h);
void orc_v1_search_plane(const uint8_t *org, const uint8_t *ref, int w, int h, int mbw, int mbh, int R, int have_sums,
                         int32_t *xy, double *so, double *rms)
{
  orc_v1_plane pl; int mb, p;
  pl.w = w; pl.h = h; pl.R = R; pl.org = org; pl.ref = ref; pl.have_sums = have_sums;
  for (mb = 0; mb < mbw * mbh; mb++)
    for (p = 0; p < 41; p++) {
      int bt, ox, oy, bw, bh, v[2] = {0, 0};
      size_t o = (size_t)mb * 41 + p;
      orc_partition_geometry(p, &bt, &ox, &oy, &bw, &bh);
      rms[o] = orc_v1_full_search(&pl, (mb % mbw) * 16 + ox, (mb / mbw) * 16 + oy, bw, bh, v, so + 2 * o);
      xy[2 * o] = v[0]; xy[2 * o + 1] = v[1];
    }
}
