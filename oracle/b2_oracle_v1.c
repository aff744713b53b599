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
#include <math.h>

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

/* ------------------------------------------------------------------------------------------------
 * F5: the partition cascade of one macroblock -- encode_one_macroblock (V1/src/block_enc.c:508-1051),
 * encode_block_rect (:1072-1334), encode_block_8 (:1337-1675), encode_block_4 (:1676-1930) for the shipped
 * configuration (num_regions == 1, search_mode == 0, currentVideo == 'C').
 *
 * Every block is searched on the four plane sets C, H, M, N (strict '<' keeps the earlier set, :652-733);
 * the cascade itself only compares those rms values with tol^2 * n and, at the macroblock level, the squared
 * normalised cross-correlation `chun` of the range block with the CO-LOCATED block of plane set C (:811-848).
 * Restated as a replay of the reference's writes to its TRANS_NODE tree (the nodes are reused between the 16x8,
 * 8x16 and 8x8 attempts and several fields survive from one attempt to the next -- e.g. `reference` of a 4x4 node
 * is only rewritten by the H comparison, and encode_block_4 also sets partition = 1 when H wins, :1773), from the
 * per-partition search results  xy/so/rms [4 sets][nmb][41].  Quirks kept: after the 16x8 / 8x16 attempts the
 * macroblock ALWAYS goes on to four 8x8 blocks (`if(mode<4)` after `for(mode=1;mode<3;mode++)`, :915), while an 8x8
 * block does stop at 8x4 / 4x8 (`mode=4`, :1641).
 * nodes: [nmb][21] in the harness' pre-order (root, child 0, its 4 children, child 1, ...).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int32_t block_type, partition, reference, x, y, pad; double scale, offset; } OrcV1Node;   /* = b2fr_node */

typedef struct { const int32_t *xy[4]; const double *so[4], *rms[4]; size_t base; } OrcV1Res;

/* the four searches of one block (node = trans of the reference): returns the best rms */
static double v1_search4(const OrcV1Res *R, int p, OrcV1Node *t, int is4x4)
{
  size_t o = R->base + p; int s;
  double rms = R->rms[0][o];
  if (R->xy[0][2 * o] || R->xy[0][2 * o + 1]) { t->x = R->xy[0][2 * o]; t->y = R->xy[0][2 * o + 1]; }   /* Q-F11 */
  t->scale = R->so[0][2 * o]; t->offset = R->so[0][2 * o + 1];
  for (s = 1; s < 4; s++) {
    if (R->rms[s][o] < rms) {
      if (s == 1 && is4x4) t->partition = 1;        /* block_enc.c:1773 */
      t->reference = s; rms = R->rms[s][o];
      t->x = R->xy[s][2 * o]; t->y = R->xy[s][2 * o + 1];
      t->offset = R->so[s][2 * o + 1]; t->scale = R->so[s][2 * o];
      t->block_type = 0;
    } else if (s == 1) t->reference = 0;
  }
  return rms;
}

/* chun of encode_one_macroblock (:811-848): doubles, column-major walk, in the reference's operation order */
double orc_v1_chun(const uint8_t *org, const uint8_t *ref, int w, int bx, int by)
{
  double Rv[256], Dv[256], sumR = 0, sumD = 0, r, d, sR = 0, sD = 0, mr = 0;
  int i, j, ii = 0;
  for (j = bx; j < bx + 16; j++)
    for (i = by; i < by + 16; i++) {
      Rv[ii] = org[(size_t)i * w + j]; Dv[ii] = ref[(size_t)i * w + j];
      sumR += Rv[ii]; sumD += Dv[ii]; ii++;
    }
  r = sumR / 256; d = sumD / 256;
  for (ii = 0; ii < 256; ii++) { sR += (Rv[ii] - r) * (Rv[ii] - r); sD += (Dv[ii] - d) * (Dv[ii] - d); }
  for (ii = 0; ii < 256; ii++) mr += ((Rv[ii] - r) / (sqrt(sR))) * ((Dv[ii] - d) / (sqrt(sD)));
  return mr * mr;
}

static int v1_rect(const OrcV1Res *R, OrcV1Node *t, int p, double tol8, int n)
{ return !(v1_search4(R, p, t, 0) > tol8 * tol8 * n); }

void orc_v1_encode_plane(const uint8_t *org, const uint8_t *refC, int w, int mbw, int mbh,
                         const int32_t *xy /*[4][nmb][41][2]*/, const double *so, const double *rms,
                         const double *tol /*[3]: tol_16, tol_8, tol_4*/, OrcV1Node *nodes /*[nmb][21]*/)
{
  const int nmb = mbw * mbh; int mb, s;
  OrcV1Res R;
  for (s = 0; s < 4; s++) { R.xy[s] = xy + (size_t)s * nmb * 82; R.so[s] = so + (size_t)s * nmb * 82; R.rms[s] = rms + (size_t)s * nmb * 41; }
  memset(nodes, 0, sizeof(OrcV1Node) * 21 * (size_t)nmb);
  for (mb = 0; mb < nmb; mb++) {
    OrcV1Node *root = nodes + (size_t)mb * 21;
    const int bx = (mb % mbw) * 16, by = (mb / mbw) * 16;
    double r16, chun;
    int mode, i, j, k;
    R.base = (size_t)mb * 41;
    root->partition = 0; root->block_type = 0; root->x = root->y = 0; root->reference = 0;
    r16 = v1_search4(&R, 0, root, 0);
    chun = orc_v1_chun(org, refC, w, bx, by);
    if (!(chun <= 1 && chun >= 0.9 && r16 > tol[0] * tol[0] * 256)) continue;       /* 16x16 accepted */
    for (mode = 1; mode < 3; mode++) {
      root->partition = mode;
      for (i = 0; i < 2; i++) {
        OrcV1Node *c = root + 1 + 5 * i;
        c->x = 0; c->y = 0;
        if (!v1_rect(&R, c, (mode == 1 ? 1 : 3) + i, tol[1], 128)) break;
      }
    }
    root->partition = 3;
    for (k = 0; k < 4; k++) {                         /* encode_block_8 */
      OrcV1Node *c = root + 1 + 5 * k;
      const int by2 = k >> 1, bx2 = k & 1;
      c->partition = 0; c->reference = 0; c->x = 0; c->y = 0;
      if (!(v1_search4(&R, 5 + k, c, 0) > tol[1] * tol[1] * 64)) continue;
      for (mode = 1; mode < 3; mode++) {
        int ok = 0;
        c->partition = mode;
        for (i = 0; i < 2; i++) {
          OrcV1Node *g = c + 1 + i;
          const int p = mode == 1 ? 9 + 2 * (2 * by2 + i) + bx2 : 17 + 4 * by2 + 2 * bx2 + i;
          g->x = 0; g->y = 0;
          if (!v1_rect(&R, g, p, tol[1], 32)) break;
          ok++;
        }
        if (ok == 2) { mode = 4; }
      }
      if (mode < 4) {
        c->partition = 3;
        for (i = 0; i < 2; i++)
          for (j = 0; j < 2; j++) {
            OrcV1Node *g = c + 1 + i * 2 + j;
            g->x = g->y = 0;
            v1_search4(&R, 25 + 4 * (2 * by2 + i) + 2 * bx2 + j, g, 1);
          }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * F8: the fractal prediction -- decode_one_macroblock (V1/src/block_dec.c:20-283), decode_block_rect (:285-758),
 * decode_block_8 (:760-976), decode_block_4 (:978-) for num_regions == 1: every leaf block of the TRANS_NODE tree is
 *     rec = (unsigned char) bound(0.5 + scale * d + offset - scale * mean_d)          (:232, :726, :913)
 * with d the displaced block of plane set `reference` (0 C, 1 H, 2 M, 3 N) and mean_d its sum table entry / n.
 * A macroblock / 8x8 block with partition 0 is one leaf, 1 or 2 two rectangles, anything else four quarters.
 * ------------------------------------------------------------------------------------------------ */
static void v1_leaf(const uint8_t *const *sets, int w, const OrcV1Node *t, int bx, int by, int bw, int bh, int mb_level, uint8_t *out)
{
  /* macroblock level: reference > 3 selects the *_temp set = C (block_dec.c:148-152); below: anything but 0,1,2 is N (:410-414) */
  const uint8_t *ref = sets[t->reference >= 0 && t->reference <= 3 ? t->reference : (mb_level ? 0 : 3)];
  const int ox = bx + t->x, oy = by + t->y;
  double sum = 0, avg; int i, j;
  for (j = 0; j < bh; j++) for (i = 0; i < bw; i++) sum += ref[(size_t)(oy + j) * w + ox + i];
  avg = sum / (double)(bh * bw);
  for (j = 0; j < bh; j++)
    for (i = 0; i < bw; i++) {
      const double a = 0.5 + t->scale * ref[(size_t)(oy + j) * w + ox + i] + t->offset - t->scale * avg;
      out[(size_t)(by + j) * w + bx + i] = (unsigned char)(a < 0.0 ? 0 : (a > 255.0 ? 255 : a));
    }
}

void orc_v1_decode_plane(const uint8_t *refC, const uint8_t *refH, const uint8_t *refM, const uint8_t *refN,
                         int w, int mbw, int mbh, const OrcV1Node *nodes /*[nmb][21]*/, uint8_t *out /* [h][w], zero where no macroblock */)
{
  const uint8_t *sets[4]; int mb, k, i;
  sets[0] = refC; sets[1] = refH; sets[2] = refM; sets[3] = refN;
  for (mb = 0; mb < mbw * mbh; mb++) {
    const OrcV1Node *root = nodes + (size_t)mb * 21;
    const int bx = (mb % mbw) * 16, by = (mb / mbw) * 16;
    if (root->partition == 0) { v1_leaf(sets, w, root, bx, by, 16, 16, 1, out); continue; }
    if (root->partition == 1 || root->partition == 2) {
      for (i = 0; i < 2; i++)
        if (root->partition == 1) v1_leaf(sets, w, root + 1 + 5 * i, bx, by + 8 * i, 16, 8, 0, out);
        else v1_leaf(sets, w, root + 1 + 5 * i, bx + 8 * i, by, 8, 16, 0, out);
      continue;
    }
    for (k = 0; k < 4; k++) {
      const OrcV1Node *c = root + 1 + 5 * k;
      const int cx = bx + (k & 1) * 8, cy = by + (k >> 1) * 8;
      if (c->partition == 0) v1_leaf(sets, w, c, cx, cy, 8, 8, 0, out);
      else if (c->partition == 1 || c->partition == 2)
        for (i = 0; i < 2; i++) {
          if (c->partition == 1) v1_leaf(sets, w, c + 1 + i, cx, cy + 4 * i, 8, 4, 0, out);
          else v1_leaf(sets, w, c + 1 + i, cx + 4 * i, cy, 4, 8, 0, out);
        }
      else
        for (i = 0; i < 4; i++) v1_leaf(sets, w, c + 1 + i, cx + (i & 1) * 4, cy + (i >> 1) * 4, 4, 4, 0, out);
    }
  }
}
