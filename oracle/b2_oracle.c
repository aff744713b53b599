/* b2_oracle.c -- CPU restatement of the reference's block-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this; the product (libb2me.so) never does.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_jm.py,
 * run in the build container where /root/reference exists) against the unmodified
 * reference objects through oracle/_ref/libjmref.so / libv1ref.so, and against the golden
 * vectors under tests/golden/ that were generated from those objects and from a
 * boundary-logged run of the stock `lencod` (scripts in oracle/gen_golden_*.py).
 *
 * Path aliases: JM/ = /root/reference/4.对比程序/jm18.5/JM/,
 *               V1/ = /root/reference/2.论文程序/ZhangLing_Yu_version1/H264Fractal/
 *
 * Plain C, no dependencies beyond libc/libm.  Build: gcc -O2 -shared -fPIC (no -march=native,
 * no FMA contraction) so the double arithmetic of the fractal part is reproducible.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define PAD_X 32   /* IMG_PAD_SIZE_X  JM/lencod/inc/defines.h:120 */
#define PAD_Y 20   /* IMG_PAD_SIZE_Y  JM/lencod/inc/defines.h:121 */
#define DISTBLK_MAX_ORC (((int64_t)INT_MAX) << 5)   /* defines.h:135 */

static inline int iclip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs_(int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------------------------
 * Tables: spiral order and MV bit lengths.   JM/lencod/src/mv_search.c:406-442, :366-374
 * ---------------------------------------------------------------------------------- */
/* out: (2R+1)^2 (x,y) pairs in integer-pel units (the reference keeps x1, x2, x4 copies). */
void orc_spiral(int R, int16_t *out_xy)
{
  int k = 1, l, i;
  out_xy[0] = out_xy[1] = 0;
  for (l = 1; l <= (R > 1 ? R : 1); l++) {
    for (i = -l + 1; i < l; i++) {
      out_xy[2*k] = (int16_t)i;  out_xy[2*k+1] = (int16_t)-l; k++;
      out_xy[2*k] = (int16_t)i;  out_xy[2*k+1] = (int16_t) l; k++;
    }
    for (i = -l; i <= l; i++) {
      out_xy[2*k] = (int16_t)-l; out_xy[2*k+1] = (int16_t)i; k++;
      out_xy[2*k] = (int16_t) l; out_xy[2*k+1] = (int16_t)i; k++;
    }
  }
}

/* mvbits[d]: Exp-Golomb length of a quarter-pel MV difference (mv_search.c:366-374):
 * 1 for d==0, else 2*floor(log2|d|)+3. */
int orc_mvbits(int d)
{
  int a = iabs_(d), n = 0;
  if (!a) return 1;
  while (a >>= 1) n++;
  return 2 * n + 3;
}

static inline int64_t orc_mv_cost(int lambda, int cx, int cy, int px, int py)
{ /* JM/lencod/inc/mv_search.h:100-104 (JCOST_CALC_SCALEUP) */
  return (int64_t)lambda * (int64_t)(orc_mvbits(cx - px) + orc_mvbits(cy - py));
}

/* ------------------------------------------------------------------------------------
 * 16 quarter-pel planes.   JM/lencod/src/img_luma.c:611-680 (+ helpers :40-600)
 * planes: [4][4][Hp][Wp] u8 with Hp = H+40, Wp = W+64; index [yy][xx] = (y&3, x&3).
 * ---------------------------------------------------------------------------------- */
static inline int clip1_255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

void orc_subpel_planes(const uint8_t *luma, int W, int H, uint8_t *planes)
{
  const int Wp = W + 2 * PAD_X, Hp = H + 2 * PAD_Y;
  const size_t PS = (size_t)Wp * Hp;
#define PL(yy, xx) (planes + ((size_t)((yy) * 4 + (xx))) * PS)
  uint8_t *p00 = PL(0,0), *p02 = PL(0,2), *p20 = PL(2,0), *p22 = PL(2,2);
  int32_t *tmp = (int32_t *)malloc(PS * sizeof(int32_t));
  int x, y;
  /* [0][0]: integer samples, edge replicated into the pad (img_luma.c:40-86) */
  for (y = 0; y < Hp; y++) {
    int sy = iclip3(0, H - 1, y - PAD_Y);
    for (x = 0; x < Wp; x++)
      p00[(size_t)y * Wp + x] = luma[(size_t)sy * W + iclip3(0, W - 1, x - PAD_X)];
  }
  /* [0][2]: horizontal 6-tap (20,-5,1) over the padded plane, neighbours clamped to the
   * padded row ends; unrounded value kept in tmp (img_luma.c:151-237) */
  for (y = 0; y < Hp; y++) {
    const uint8_t *s = p00 + (size_t)y * Wp;
    for (x = 0; x < Wp; x++) {
#define CX(v) ((v) < 0 ? 0 : ((v) > Wp - 1 ? Wp - 1 : (v)))
      int is = 20 * (s[x] + s[CX(x + 1)]) - 5 * (s[CX(x - 1)] + s[CX(x + 2)]) + (s[CX(x - 2)] + s[CX(x + 3)]);
      tmp[(size_t)y * Wp + x] = is;
      p02[(size_t)y * Wp + x] = (uint8_t)clip1_255((is + 16) >> 5);
    }
  }
  /* [2][0]: vertical 6-tap of [0][0] (img_luma.c:257-331); [2][2]: vertical 6-tap of the
   * unrounded horizontal intermediates, (is+512)>>10 (img_luma.c:347-422) */
  for (y = 0; y < Hp; y++) {
#define CY(v) ((v) < 0 ? 0 : ((v) > Hp - 1 ? Hp - 1 : (v)))
    size_t a = (size_t)y * Wp, d = (size_t)CY(y + 1) * Wp, b = (size_t)CY(y - 1) * Wp,
           e = (size_t)CY(y + 2) * Wp, c = (size_t)CY(y - 2) * Wp, f = (size_t)CY(y + 3) * Wp;
    for (x = 0; x < Wp; x++) {
      int is = 20 * (p00[a + x] + p00[d + x]) - 5 * (p00[b + x] + p00[e + x]) + (p00[c + x] + p00[f + x]);
      int it = 20 * (tmp[a + x] + tmp[d + x]) - 5 * (tmp[b + x] + tmp[e + x]) + (tmp[c + x] + tmp[f + x]);
      p20[a + x] = (uint8_t)clip1_255((is + 16) >> 5);
      p22[a + x] = (uint8_t)clip1_255((it + 512) >> 10);
    }
  }
  /* twelve quarter planes: (a+b+1)>>1 (img_luma.c:440-600, call list :653-678) */
  for (y = 0; y < Hp; y++) {
    size_t r = (size_t)y * Wp, rn = (size_t)CY(y + 1) * Wp;
    for (x = 0; x < Wp; x++) {
      int xn = CX(x + 1);
#define AVG(a, b) ((uint8_t)(((int)(a) + (int)(b) + 1) >> 1))
      PL(0,1)[r + x] = AVG(p00[r + x], p02[r + x]);
      PL(1,0)[r + x] = AVG(p00[r + x], p20[r + x]);
      PL(1,1)[r + x] = AVG(p02[r + x], p20[r + x]);
      PL(1,2)[r + x] = AVG(p02[r + x], p22[r + x]);
      PL(2,1)[r + x] = AVG(p20[r + x], p22[r + x]);
      PL(0,3)[r + x] = AVG(p02[r + x], p00[r + xn]);
      PL(1,3)[r + x] = AVG(p02[r + x], p20[r + xn]);
      PL(2,3)[r + x] = AVG(p22[r + x], p20[r + xn]);
      PL(3,0)[r + x] = AVG(p20[r + x], p00[rn + x]);
      PL(3,1)[r + x] = AVG(p20[r + x], p02[rn + x]);
      PL(3,2)[r + x] = AVG(p22[r + x], p02[rn + x]);
      PL(3,3)[r + x] = AVG(p02[rn + x], p20[r + xn]);
    }
  }
  free(tmp);
#undef PL
#undef CX
#undef CY
#undef AVG
}

/* UMVLine4X (JM/lencod/inc/refbuf.h:22-26): plane [y&3][x&3], block ORIGIN clamped to
 * [-20, H+3] x [-32, W+15] (size_y_pad/size_x_pad, mbuffer.c:549-550). Returns pointer
 * into the padded plane (row stride Wp). */
static inline const uint8_t *orc_umv_line4x(const uint8_t *planes, int W, int H, int y, int x)
{
  const int Wp = W + 2 * PAD_X, Hp = H + 2 * PAD_Y;
  const uint8_t *pl = planes + ((size_t)((y & 3) * 4 + (x & 3))) * ((size_t)Wp * Hp);
  int yy = iclip3(-PAD_Y, H + 3, y >> 2), xx = iclip3(-PAD_X, W + 15, x >> 2);
  return pl + (size_t)(yy + PAD_Y) * Wp + (xx + PAD_X);
}

/* computeSAD (JM/lencod/src/me_distortion.c:349-426), luma only, no early exit (the early
 * exit is result-neutral, SURVEY Q-J1).  cand = absolute quarter-pel coordinates.
 * Returns the plain SAD (caller shifts <<5). */
int orc_sad(const uint8_t *planes, int W, int H, const uint8_t *cur, int cur_stride,
            int bsx, int bsy, int cand_x, int cand_y)
{
  const int Wp = W + 2 * PAD_X;
  const uint8_t *ref = orc_umv_line4x(planes, W, H, cand_y, cand_x);
  int x, y, s = 0;
  for (y = 0; y < bsy; y++)
    for (x = 0; x < bsx; x++)
      s += iabs_((int)cur[y * cur_stride + x] - (int)ref[(size_t)y * Wp + x]);
  return s;
}

/* computeSSE (JM/lencod/src/me_distortion.c:1190-1255), luma only: as computeSAD with squared differences. */
int orc_sse(const uint8_t *planes, int W, int H, const uint8_t *cur, int cur_stride,
            int bsx, int bsy, int cand_x, int cand_y)
{
  const int Wp = W + 2 * PAD_X;
  const uint8_t *ref = orc_umv_line4x(planes, W, H, cand_y, cand_x);
  int x, y, s = 0;
  for (y = 0; y < bsy; y++)
    for (x = 0; x < bsx; x++) {
      const int d = (int)cur[y * cur_stride + x] - (int)ref[(size_t)y * Wp + x];
      s += d * d;
    }
  return s;
}

/* HadamardSAD4x4 (me_distortion.c:175-258): 2-D 4-point Hadamard, sum |coef|, (s+1)>>1.
 * The butterfly network is sign/permutation-equivalent to H4 * D * H4; the absolute sum is
 * invariant to row/column sign flips and permutations of the transform, so a plain
 * separable Hadamard is used here.  diff: 16 shorts, raster. */
int orc_hadamard4x4(const int16_t *d)
{
  int m[16], i, s = 0;
  for (i = 0; i < 4; i++) {               /* rows */
    int a = d[4*i], b = d[4*i+1], c = d[4*i+2], e = d[4*i+3];
    m[4*i] = a + b + c + e; m[4*i+1] = a - b + c - e; m[4*i+2] = a + b - c - e; m[4*i+3] = a - b - c + e;
  }
  for (i = 0; i < 4; i++) {               /* columns */
    int a = m[i], b = m[4+i], c = m[8+i], e = m[12+i];
    s += iabs_(a + b + c + e) + iabs_(a - b + c - e) + iabs_(a + b - c - e) + iabs_(a - b - c + e);
  }
  return (s + 1) >> 1;
}

/* HadamardSAD8x8 (me_distortion.c:266-341): (sum|H8 D H8| + 2) >> 2 */
int orc_hadamard8x8(const int16_t *d)
{
  int m[64], t[8], i, j, k, s = 0;
  for (j = 0; j < 8; j++) {
    for (i = 0; i < 8; i++) t[i] = d[8*j + i];
    for (k = 1; k < 8; k <<= 1)
      for (i = 0; i < 8; i++) if (!(i & k)) { int a = t[i], b = t[i | k]; t[i] = a + b; t[i | k] = a - b; }
    for (i = 0; i < 8; i++) m[8*j + i] = t[i];
  }
  for (i = 0; i < 8; i++) {
    for (j = 0; j < 8; j++) t[j] = m[8*j + i];
    for (k = 1; k < 8; k <<= 1)
      for (j = 0; j < 8; j++) if (!(j & k)) { int a = t[j], b = t[j | k]; t[j] = a + b; t[j | k] = a - b; }
    for (j = 0; j < 8; j++) s += iabs_(t[j]);
  }
  return (s + 2) >> 2;
}

/* computeSATD (me_distortion.c:745-825), no early exit.  Each 4x4 (8x8) tile origin goes
 * through UMVLine4X separately (per-tile clamp, :771 / :801). */
int orc_satd(const uint8_t *planes, int W, int H, const uint8_t *cur, int cur_stride,
             int bsx, int bsy, int cand_x, int cand_y, int test8x8)
{
  const int Wp = W + 2 * PAD_X, T = test8x8 ? 8 : 4;
  int bx, by, i, j, s = 0; int16_t diff[64];
  for (by = 0; by < bsy; by += T)
    for (bx = 0; bx < bsx; bx += T) {
      const uint8_t *ref = orc_umv_line4x(planes, W, H, cand_y + (by << 2), cand_x + (bx << 2));
      for (j = 0; j < T; j++)
        for (i = 0; i < T; i++)
          diff[j * T + i] = (int16_t)((int)cur[(by + j) * cur_stride + bx + i] - (int)ref[(size_t)j * Wp + i]);
      s += test8x8 ? orc_hadamard8x8(diff) : orc_hadamard4x4(diff);
    }
  return s;
}

/* distortion4x4/8x8{SAD,SSE,SATD} (me_distortion.c:38-134) on a precomputed difference block: dist_scale(d) = d << 5.
 * kind 0 SAD, 1 SSE, 2 SATD; n = 4 or 8; diff raster n*n shorts. */
int64_t orc_distortion(int kind, int n, const int16_t *diff)
{
  int64_t d = 0; int k;
  if (kind == 2) d = n == 4 ? orc_hadamard4x4(diff) : orc_hadamard8x8(diff);
  else for (k = 0; k < n * n; k++) d += kind == 0 ? iabs_(diff[k]) : (int64_t)diff[k] * diff[k];
  return d << 5;
}

/* ------------------------------------------------------------------------------------
 * Searches.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int W, H, nrefs, R;          /* frame size (multiple of 16), refs, max search range (pel) */
  const uint8_t *cur;          /* W x H */
  const uint8_t *planes;       /* [nrefs][16][Hp][Wp] */
  int16_t *spiral;             /* (2R+1)^2 x 2, integer pel */
} OrcFrame;

void *orc_frame_create(int W, int H, int nrefs, int R, const uint8_t *cur, const uint8_t *refs /*[nrefs][H][W]*/)
{
  OrcFrame *f = (OrcFrame *)calloc(1, sizeof(OrcFrame));
  const size_t PS = (size_t)(W + 2 * PAD_X) * (H + 2 * PAD_Y);
  int r, np = (2 * R + 1) * (2 * R + 1); if (np < 9) np = 9;
  uint8_t *pl = (uint8_t *)malloc(PS * 16 * nrefs);
  uint8_t *c = (uint8_t *)malloc((size_t)W * H);
  memcpy(c, cur, (size_t)W * H);
  for (r = 0; r < nrefs; r++) orc_subpel_planes(refs + (size_t)r * W * H, W, H, pl + PS * 16 * r);
  f->W = W; f->H = H; f->nrefs = nrefs; f->R = R; f->cur = c; f->planes = pl;
  f->spiral = (int16_t *)malloc(sizeof(int16_t) * 2 * np);
  orc_spiral(R, f->spiral);
  return f;
}
/* Explicit weighted prediction of the single-list search (computeSADWP / SATDWP / SSEWP, me_distortion.c:434-517,
 * :833-935, :1262-1345): every reference sample v read by the distortion is replaced by
 *     clip1(((weight * v + round) >> log_denom) + offset),  round = log_denom ? 1 << (log_denom-1) : 0.
 * The mapping is pointwise on the fetched (integer or interpolated) sample, so applying it to the 16 quarter-pel
 * planes of the reference is the same computation. */
void orc_frame_set_weights(void *h, int ref, int weight, int offset, int log_denom)
{
  OrcFrame *f = (OrcFrame *)h;
  const size_t n = (size_t)16 * (f->W + 2 * PAD_X) * (f->H + 2 * PAD_Y);
  uint8_t *p = (uint8_t *)f->planes + (size_t)ref * n;
  const int rnd = log_denom ? 1 << (log_denom - 1) : 0;
  uint8_t lut[256]; size_t i; int v;
  for (v = 0; v < 256; v++) lut[v] = (uint8_t)clip1_255(((weight * v + rnd) >> log_denom) + offset);
  for (i = 0; i < n; i++) p[i] = lut[p[i]];
}

void orc_frame_destroy(void *h)
{ OrcFrame *f = (OrcFrame *)h; free((void *)f->cur); free((void *)f->planes); free(f->spiral); free(f); }
const uint8_t *orc_frame_planes(void *h, int r)
{ OrcFrame *f = (OrcFrame *)h; return f->planes + (size_t)r * 16 * (size_t)(f->W + 2*PAD_X) * (f->H + 2*PAD_Y); }

/* full_search_motion_estimation (JM/lencod/src/me_fullsearch.c:39-103), RDO build
 * (check_for_00 is only live when !rdopt).  mv: in = centre (relative quarter-pel MV),
 * out = best integer MV.  search_range in pel = min(max_x,max_y)>>2 of the block's window. */
int64_t orc_full_search(const OrcFrame *f, int ref, int pos_x, int pos_y, int bsx, int bsy,
                        const int16_t *pred_mv, int16_t *mv, int search_range,
                        int64_t min_mcost, int lambda_factor)
{
  const uint8_t *pl = orc_frame_planes((void *)f, ref);
  const uint8_t *cur = f->cur + (size_t)pos_y * f->W + pos_x;
  int max_pos = (2 * search_range + 1) * (2 * search_range + 1), pos, best_pos = 0;
  int cx = (pos_x << 2) + mv[0], cy = (pos_y << 2) + mv[1];
  int px = (pos_x << 2) + pred_mv[0], py = (pos_y << 2) + pred_mv[1];
  for (pos = 0; pos < max_pos; pos++) {
    int candx = cx + 4 * f->spiral[2*pos], candy = cy + 4 * f->spiral[2*pos+1];
    int64_t mcost = orc_mv_cost(lambda_factor, candx, candy, px, py);
    if (mcost >= min_mcost) continue;
    mcost += ((int64_t)orc_sad(pl, f->W, f->H, cur, f->W, bsx, bsy, candx, candy)) << 5;
    if (mcost < min_mcost) { best_pos = pos; min_mcost = mcost; }
  }
  if (best_pos) { mv[0] = (int16_t)(mv[0] + 4 * f->spiral[2*best_pos]); mv[1] = (int16_t)(mv[1] + 4 * f->spiral[2*best_pos+1]); }
  return min_mcost;
}

/* sub_pel_motion_estimation (me_fullsearch.c:186-289), RDO build.  start_hp/start_qp =
 * p_Vid->start_me_refinement_hp/_qp (mv_search.c:445-446); metric_h/q: 0 SAD, 2 SATD. */
int64_t orc_sub_pel(const OrcFrame *f, int ref, int pos_x, int pos_y, int bsx, int bsy,
                    const int16_t *pred_mv, int16_t *mv, int64_t min_mcost,
                    const int *lambda /*[3]*/, int start_hp, int start_qp,
                    int metric_h, int metric_q, int test8x8)
{
  const uint8_t *pl = orc_frame_planes((void *)f, ref);
  const uint8_t *cur = f->cur + (size_t)pos_y * f->W + pos_x;
  int pos, best_pos, lam = lambda[1];
  int max_pos2 = 9; /* search_pos2 (init_mv_block, mv_search.c:718) ; imax(1,9) */
  for (best_pos = 0, pos = start_hp; pos < max_pos2; pos++) {
    int cx = mv[0] + 2 * f->spiral[2*pos], cy = mv[1] + 2 * f->spiral[2*pos+1];
    int64_t mcost = orc_mv_cost(lam, cx, cy, pred_mv[0], pred_mv[1]);
    int d;
    if (mcost >= min_mcost) continue;
    d = (metric_h == 2) ? orc_satd(pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2), test8x8)
      : (metric_h == 1) ? orc_sse (pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2))
                        : orc_sad (pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2));
    mcost += ((int64_t)d) << 5;
    if (mcost < min_mcost) { min_mcost = mcost; best_pos = pos; }
  }
  if (best_pos) { mv[0] = (int16_t)(mv[0] + 2 * f->spiral[2*best_pos]); mv[1] = (int16_t)(mv[1] + 2 * f->spiral[2*best_pos+1]); }
  if (!start_qp) min_mcost = DISTBLK_MAX_ORC;
  lam = lambda[2];
  for (best_pos = 0, pos = start_qp; pos < 9; pos++) {
    int cx = mv[0] + f->spiral[2*pos], cy = mv[1] + f->spiral[2*pos+1];
    int64_t mcost = orc_mv_cost(lam, cx, cy, pred_mv[0], pred_mv[1]);
    int d;
    if (mcost >= min_mcost) continue;
    d = (metric_q == 2) ? orc_satd(pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2), test8x8)
      : (metric_q == 1) ? orc_sse (pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2))
                        : orc_sad (pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2));
    mcost += ((int64_t)d) << 5;
    if (mcost < min_mcost) { min_mcost = mcost; best_pos = pos; }
  }
  if (best_pos) { mv[0] = (int16_t)(mv[0] + f->spiral[2*best_pos]); mv[1] = (int16_t)(mv[1] + f->spiral[2*best_pos+1]); }
  return min_mcost;
}

/* partition p (0..40) -> blocktype (1..7, JM block_size[] macroblock.h:58) and offset in MB */
/* full_sub_pel_motion_estimation (me_fullsearch.c:409-469), RDO build (check_position0 is a !rdopt rule):
 * the 81 quarter-pel positions of spiral_search (radius 4) around mv, computePredQPel, strict '<' in scan order. */
int64_t orc_full_sub_pel(const OrcFrame *f, int ref, int pos_x, int pos_y, int bsx, int bsy,
                         const int16_t *pred_mv, int16_t *mv, int64_t min_mcost, int lambda_q, int metric_q, int test8x8)
{
  const uint8_t *pl = orc_frame_planes((void *)f, ref);
  const uint8_t *cur = f->cur + (size_t)pos_y * f->W + pos_x;
  int pos, best_pos;
  for (best_pos = 0, pos = 0; pos < 81; pos++) {
    int cx = mv[0] + f->spiral[2*pos], cy = mv[1] + f->spiral[2*pos+1];
    int64_t mcost = orc_mv_cost(lambda_q, cx, cy, pred_mv[0], pred_mv[1]);
    int d;
    if (mcost >= min_mcost) continue;
    d = (metric_q == 2) ? orc_satd(pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2), test8x8)
      : (metric_q == 1) ? orc_sse (pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2))
                        : orc_sad (pl, f->W, f->H, cur, f->W, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2));
    mcost += ((int64_t)d) << 5;
    if (mcost < min_mcost) { min_mcost = mcost; best_pos = pos; }
  }
  if (best_pos) { mv[0] = (int16_t)(mv[0] + f->spiral[2*best_pos]); mv[1] = (int16_t)(mv[1] + f->spiral[2*best_pos+1]); }
  return min_mcost;
}

static const uint8_t ORC_BS[8][2] = {{0,0},{16,16},{16,8},{8,16},{8,8},{8,4},{4,8},{4,4}};
static const int ORC_FIRST[8] = {0, 0, 1, 3, 5, 9, 17, 25};
void orc_partition_geometry(int p, int *bt, int *ox, int *oy, int *w, int *h)
{
  int t = 7, k, per_row;
  while (ORC_FIRST[t] > p) t--;
  k = p - ORC_FIRST[t]; *w = ORC_BS[t][0]; *h = ORC_BS[t][1]; per_row = 16 / *w;
  *bt = t; *ox = (k % per_row) * *w; *oy = (k / per_row) * *h;
}

/* get_search_range (mv_search.c:70-92): per-block search range in pel.
 * mode = p_Inp->full_search (RestrictSearchRange): 2 = no restriction. */
int orc_block_search_range(int R, int mode, int ref, int blocktype)
{
  int q = R << 2, scale = 1;
  if (mode == 1) scale = (ref < 1 ? ref : 1) + 1;
  else if (mode != 2) scale = ((ref < 1 ? ref : 1) + 1) * (blocktype < 2 ? blocktype : 2);
  return (q / scale) >> 2;
}

/* Batch form == product b2me_search_frame() semantics == jmh_search_frame().
 * metric_f is SAD; metric_h/q SAD(0) or SATD(2). */
void orc_search_frame(void *h, int mb_first, int mb_count, const int16_t *pred, const int16_t *center,
                      const int *lambda_factor, int restrict_mode, int metric_h, int metric_q, int do_subpel,
                      int16_t *mv_int, int64_t *cost_int, int16_t *mv_sub, int64_t *cost_sub)
{
  OrcFrame *f = (OrcFrame *)h;
  int mbw = f->W / 16, m, r, p;
  int start_hp = (0 != metric_h) ? 0 : 1, start_qp = (metric_h != metric_q) ? 0 : 1;
  for (m = mb_first; m < mb_first + mb_count; m++)
    for (r = 0; r < f->nrefs; r++)
      for (p = 0; p < 41; p++) {
        int bt, ox, oy, w, hh; size_t i = ((size_t)m * f->nrefs + r) * 41 + p;
        int16_t mv[2]; int64_t c; int sr;
        orc_partition_geometry(p, &bt, &ox, &oy, &w, &hh);
        sr = orc_block_search_range(f->R, restrict_mode, r, bt);
        mv[0] = center[2*i]; mv[1] = center[2*i+1];
        c = orc_full_search(f, r, (m % mbw) * 16 + ox, (m / mbw) * 16 + oy, w, hh, pred + 2*i, mv, sr,
                            DISTBLK_MAX_ORC, lambda_factor[0]);
        mv_int[2*i] = mv[0]; mv_int[2*i+1] = mv[1]; cost_int[i] = c;
        if (do_subpel) {
          if (!start_hp) c = DISTBLK_MAX_ORC;
          if (do_subpel == 2)          /* SubPelME = full_sub_pel_motion_estimation (EPZSSubPelME == 2, me_epzs_common.c:157) */
            c = orc_full_sub_pel(f, r, (m % mbw) * 16 + ox, (m / mbw) * 16 + oy, w, hh, pred + 2*i, mv, c,
                                 lambda_factor[2], metric_q, 0);
          else
            c = orc_sub_pel(f, r, (m % mbw) * 16 + ox, (m / mbw) * 16 + oy, w, hh, pred + 2*i, mv, c,
                            lambda_factor, start_hp, start_qp, metric_h, metric_q, 0);
          mv_sub[2*i] = mv[0]; mv_sub[2*i+1] = mv[1]; cost_sub[i] = c;
        }
      }
}

/* Call-level entry points (one call of full_search_motion_estimation / sub_pel_motion_estimation)
 * used against the boundary-logged golden vectors. */
int64_t orc_call_full_search(void *h, int ref, int pos_x, int pos_y, int blocktype, const int16_t *pred_mv,
                             int16_t *mv_inout, int search_range, int64_t min_mcost, int lambda_factor)
{
  return orc_full_search((OrcFrame *)h, ref, pos_x, pos_y, ORC_BS[blocktype][0], ORC_BS[blocktype][1], pred_mv,
                         mv_inout, search_range, min_mcost, lambda_factor);
}
int64_t orc_call_sub_pel(void *h, int ref, int pos_x, int pos_y, int blocktype, const int16_t *pred_mv,
                         int16_t *mv_inout, int64_t min_mcost, int lambda_h, int lambda_q,
                         int metric_h, int metric_q)
{
  int lam[3]; int start_hp = (0 != metric_h) ? 0 : 1, start_qp = (metric_h != metric_q) ? 0 : 1;
  lam[0] = 0; lam[1] = lambda_h; lam[2] = lambda_q;
  return orc_sub_pel((OrcFrame *)h, ref, pos_x, pos_y, ORC_BS[blocktype][0], ORC_BS[blocktype][1], pred_mv,
                     mv_inout, min_mcost, lam, start_hp, start_qp, metric_h, metric_q, 0);
}


/* ------------------------------------------------------------------------------------
 * Motion-compensated luma prediction of list 0: luma_prediction with p_dir == 0, no weighting
 * (JM/lencod/src/mc_prediction.c:144-236) = OneComponentLumaPrediction (:117-136): block_size_y rows of block_size_x
 * samples copied from UMVLine4X(ref, 4*(pix_y+block_y) + mv_y, 4*(pix_x+block_x) + mv_x) -- the same plane address
 * (and whole-block origin clamp) the distortions read, which the SAD/SATD tests pin against the reference.
 *   mb_mode [nmb]: 1 16x16, 2 16x8, 3 8x16, 8 P8x8;  b8mode [nmb][4]: 4 8x8, 5 8x4, 6 4x8, 7 4x4 (per 8x8 quadrant)
 *   ref8 [nmb][4]: reference of each quadrant;  mv [nmb][nrefs][41][2] quarter-pel (a search result array)
 *   orig_blk / pred_blk [nmb*16][16]: the sixteen 4x4 blocks of every MB in raster order, each raster (b2tq layout)
 * ---------------------------------------------------------------------------------- */
static void orc_block_partition(int mode, const uint8_t *b8, int bx, int by, int *p, int *ox, int *oy)
{
  const int qd = (by >> 1) * 2 + (bx >> 1);
  if (mode == 1) { *p = 0; *ox = 0; *oy = 0; }
  else if (mode == 2) { *p = 1 + (by >> 1); *ox = 0; *oy = 8 * (by >> 1); }
  else if (mode == 3) { *p = 3 + (bx >> 1); *ox = 8 * (bx >> 1); *oy = 0; }
  else switch (b8[qd]) {
    case 4:  *p = 5 + qd; *ox = 8 * (bx >> 1); *oy = 8 * (by >> 1); break;
    case 5:  *p = 9 + by * 2 + (bx >> 1); *ox = 8 * (bx >> 1); *oy = 4 * by; break;
    case 6:  *p = 17 + (by >> 1) * 4 + bx; *ox = 4 * bx; *oy = 8 * (by >> 1); break;
    default: *p = 25 + by * 4 + bx; *ox = 4 * bx; *oy = 4 * by; break;
  }
}
void orc_mc_luma(void *h, const uint8_t *mb_mode, const uint8_t *b8mode, const int8_t *ref8, const int16_t *mv,
                 uint8_t *orig_blk, uint8_t *pred_blk)
{
  OrcFrame *f = (OrcFrame *)h;
  const int mbw = f->W / 16, nmb = mbw * (f->H / 16), Wp = f->W + 2 * PAD_X;
  int m, k, i, j;
  for (m = 0; m < nmb; m++)
    for (k = 0; k < 16; k++) {
      const int bx = k & 3, by = k >> 2, qd = (by >> 1) * 2 + (bx >> 1), r = ref8[m * 4 + qd];
      int p, ox, oy;
      const int16_t *v;
      const uint8_t *ref, *pl = orc_frame_planes(h, r);
      orc_block_partition(mb_mode[m], b8mode + m * 4, bx, by, &p, &ox, &oy);
      v = mv + (((size_t)m * f->nrefs + r) * 41 + p) * 2;
      ref = orc_umv_line4x(pl, f->W, f->H, 4 * ((m / mbw) * 16 + oy) + v[1], 4 * ((m % mbw) * 16 + ox) + v[0]);
      for (i = 0; i < 4; i++)
        for (j = 0; j < 4; j++) {
          pred_blk[((size_t)m * 16 + k) * 16 + i * 4 + j] = ref[(size_t)(4 * by - oy + i) * Wp + (4 * bx - ox + j)];
          orig_blk[((size_t)m * 16 + k) * 16 + i * 4 + j] = f->cur[(size_t)((m / mbw) * 16 + 4 * by + i) * f->W + (m % mbw) * 16 + 4 * bx + j];
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Bi-predictive search (B slices): distortion against the average of two reference blocks.
 * ---------------------------------------------------------------------------------- */
typedef struct {            /* mirrors b2me_bipred_job (include/b2me.h) */
  int64_t min_mcost;
  int16_t pos_x, pos_y, blocktype, ref1, ref2, search_range;
  int16_t pred1[2], pred2[2], mv1[2], mv2[2];
  int16_t weight1, weight2, offset_bi, reserved;
} OrcBiJob;
typedef struct { int64_t cost_int, cost_sub; int16_t mv_int[2], mv_sub[2]; } OrcBiResult;

/* the prediction sample of computeBiPred*1 (me_distortion.c:556: (r1 + r2 + 1) >> 1) and computeBiPred*2
 * (:655-657: iClip1(max, ((w1*r1 + w2*r2 + 2*wp_luma_round) >> (denom+1)) + offsetBi)) */
static inline int orc_bi_pel(int r1, int r2, int wp, int w1, int w2, int off, int denom)
{
  if (!wp) return (r1 + r2 + 1) >> 1;
  { const int lround = 2 * (denom ? 1 << (denom - 1) : 0);
    return clip1_255(((w1 * r1 + w2 * r2 + lround) >> (denom + 1)) + off); }
}

/* computeBiPredSAD1/2 (me_distortion.c:525-620, :628-737), computeBiPredSSE1/2 (:1353-1442, :1450-1549),
 * computeBiPredSATD1/2 (:943-1040, :1048-1182), luma only, no early exit (result-neutral as in computeSAD).
 * metric 0 SAD, 1 SSE, 2 SATD.  cand1/cand2 = absolute quarter-pel coordinates.  SAD/SSE clamp the block origin
 * once per reference (:546-547); SATD clamps every 4x4 (8x8) tile origin (:971-972). */
int orc_bipred_dist(const uint8_t *pl1, const uint8_t *pl2, int W, int H, const uint8_t *cur, int cur_stride,
                    int bsx, int bsy, int c1x, int c1y, int c2x, int c2y, int metric, int test8x8,
                    int wp, int w1, int w2, int off, int denom)
{
  const int Wp = W + 2 * PAD_X;
  int x, y, s = 0;
  if (metric != 2) {
    const uint8_t *r1 = orc_umv_line4x(pl1, W, H, c1y, c1x), *r2 = orc_umv_line4x(pl2, W, H, c2y, c2x);
    for (y = 0; y < bsy; y++)
      for (x = 0; x < bsx; x++) {
        const int d = (int)cur[y * cur_stride + x] - orc_bi_pel(r1[(size_t)y * Wp + x], r2[(size_t)y * Wp + x], wp, w1, w2, off, denom);
        s += metric == 0 ? iabs_(d) : d * d;
      }
    return s;
  }
  { const int T = test8x8 ? 8 : 4; int bx, by, i, j; int16_t diff[64];
    for (by = 0; by < bsy; by += T)
      for (bx = 0; bx < bsx; bx += T) {
        const uint8_t *r1 = orc_umv_line4x(pl1, W, H, c1y + (by << 2), c1x + (bx << 2));
        const uint8_t *r2 = orc_umv_line4x(pl2, W, H, c2y + (by << 2), c2x + (bx << 2));
        for (j = 0; j < T; j++)
          for (i = 0; i < T; i++)
            diff[j * T + i] = (int16_t)((int)cur[(by + j) * cur_stride + bx + i] -
                                        orc_bi_pel(r1[(size_t)j * Wp + i], r2[(size_t)j * Wp + i], wp, w1, w2, off, denom));
        s += test8x8 ? orc_hadamard8x8(diff) : orc_hadamard4x4(diff);
      }
  }
  return s;
}

/* full_search_bipred_motion_estimation (me_fullsearch.c:112-176) followed, if do_subpel, by
 * sub_pel_bipred_motion_estimation (:300-399) sequenced as BiPredBlockMotionSearch does (mv_search.c:1117-1126:
 * the bound is reset to DISTBLK_MAX when the half-pel stage re-examines the centre).  F_PEL metric = SAD.
 * lambda[3] = F, H, Q;  the weighted 8x8-Hadamard variant is not restated (reference bug Q-J5). */
void orc_bipred_search(void *h, int njobs, const OrcBiJob *jobs, const int *lambda, int metric_h, int metric_q,
                       int do_subpel, int test8x8, int wp, int denom, OrcBiResult *out)
{
  const OrcFrame *f = (const OrcFrame *)h;
  int n;
  for (n = 0; n < njobs; n++) {
    const OrcBiJob *J = &jobs[n];
    const int bsx = ORC_BS[J->blocktype][0], bsy = ORC_BS[J->blocktype][1];
    const uint8_t *pl1 = orc_frame_planes(h, J->ref1), *pl2 = orc_frame_planes(h, J->ref2);
    const uint8_t *cur = f->cur + (size_t)J->pos_y * f->W + J->pos_x;
    const int ox = J->pos_x << 2, oy = J->pos_y << 2;
    const int sr = J->search_range, max_pos = (2 * sr + 1) * (2 * sr + 1);
    const int start_hp_cfg = (0 != metric_h) ? 0 : 1, start_qp = (metric_h != metric_q) ? 0 : 1;
    int16_t mv1[2] = {J->mv1[0], J->mv1[1]};
    int64_t min_mcost = J->min_mcost;
    int pos, best_pos = 0;
    const int64_t c2f = orc_mv_cost(lambda[0], J->mv2[0], J->mv2[1], J->pred2[0], J->pred2[1]);
    for (pos = 0; sr >= 0 && pos < max_pos; pos++) {       /* search_range == -1: the sub-pel call alone */
      const int cx = mv1[0] + 4 * f->spiral[2*pos], cy = mv1[1] + 4 * f->spiral[2*pos+1];
      int64_t mcost = orc_mv_cost(lambda[0], cx, cy, J->pred1[0], J->pred1[1]) + c2f;
      if (mcost >= min_mcost) continue;
      mcost += ((int64_t)orc_bipred_dist(pl1, pl2, f->W, f->H, cur, f->W, bsx, bsy, ox + cx, oy + cy, ox + J->mv2[0], oy + J->mv2[1],
                                         0, test8x8, wp, J->weight1, J->weight2, J->offset_bi, denom)) << 5;
      if (mcost < min_mcost) { best_pos = pos; min_mcost = mcost; }
    }
    if (best_pos) { mv1[0] = (int16_t)(mv1[0] + 4 * f->spiral[2*best_pos]); mv1[1] = (int16_t)(mv1[1] + 4 * f->spiral[2*best_pos+1]); }
    out[n].mv_int[0] = mv1[0]; out[n].mv_int[1] = mv1[1]; out[n].cost_int = min_mcost;
    out[n].mv_sub[0] = mv1[0]; out[n].mv_sub[1] = mv1[1]; out[n].cost_sub = min_mcost;
    if (!do_subpel) continue;
    if (!start_hp_cfg && sr >= 0) min_mcost = DISTBLK_MAX_ORC;           /* the caller's reset, mv_search.c:1119-1120 */
    {
      /* do_subpel == 2: full_sub_pel_bipred_motion_estimation (me_fullsearch.c:478-538), one stage of 81 quarter-pel positions */
      const int full81 = do_subpel == 2;
      int stage;
      for (stage = 0; stage < (full81 ? 1 : 2); stage++) {
        const int lam = full81 ? lambda[2] : lambda[1 + stage], metric = (stage || full81) ? metric_q : metric_h, step = (stage || full81) ? 1 : 2;
        const int64_t c2 = orc_mv_cost(lam, J->mv2[0], J->mv2[1], J->pred2[0], J->pred2[1]);
        int start = full81 ? 0 : (stage ? start_qp : ((min_mcost == DISTBLK_MAX_ORC) ? 0 : start_hp_cfg));
        if (stage && !start_qp) min_mcost = DISTBLK_MAX_ORC;             /* me_fullsearch.c:364-365 */
        for (best_pos = 0, pos = start; pos < (full81 ? 81 : 9); pos++) {
          const int cx = mv1[0] + step * f->spiral[2*pos], cy = mv1[1] + step * f->spiral[2*pos+1];
          int64_t mcost = orc_mv_cost(lam, cx, cy, J->pred1[0], J->pred1[1]) + c2;
          if (mcost >= min_mcost) continue;
          mcost += ((int64_t)orc_bipred_dist(pl1, pl2, f->W, f->H, cur, f->W, bsx, bsy, ox + cx, oy + cy, ox + J->mv2[0], oy + J->mv2[1],
                                             metric, test8x8, wp, J->weight1, J->weight2, J->offset_bi, denom)) << 5;
          if (mcost < min_mcost) { min_mcost = mcost; best_pos = pos; }
        }
        if (best_pos) { mv1[0] = (int16_t)(mv1[0] + step * f->spiral[2*best_pos]); mv1[1] = (int16_t)(mv1[1] + step * f->spiral[2*best_pos+1]); }
      }
    }
    out[n].mv_sub[0] = mv1[0]; out[n].mv_sub[1] = mv1[1]; out[n].cost_sub = min_mcost;
  }
}

/* computeSAD / computeSSE / computeSATD at explicit candidates (mirrors b2me_distortion_candidates): out = distortion << 5 */
typedef struct { int16_t pos_x, pos_y, blocktype, ref, mv[2]; } OrcCand;
void orc_distortion_candidates(void *h, int metric, int test8x8, int n, const OrcCand *c, int64_t *out)
{
  const OrcFrame *f = (const OrcFrame *)h; int i;
  for (i = 0; i < n; i++) {
    const int bsx = ORC_BS[c[i].blocktype][0], bsy = ORC_BS[c[i].blocktype][1];
    const uint8_t *pl = orc_frame_planes(h, c[i].ref), *cur = f->cur + (size_t)c[i].pos_y * f->W + c[i].pos_x;
    const int qx = (c[i].pos_x << 2) + c[i].mv[0], qy = (c[i].pos_y << 2) + c[i].mv[1];
    const int d = metric == 2 ? orc_satd(pl, f->W, f->H, cur, f->W, bsx, bsy, qx, qy, test8x8)
                : metric == 1 ? orc_sse(pl, f->W, f->H, cur, f->W, bsx, bsy, qx, qy)
                              : orc_sad(pl, f->W, f->H, cur, f->W, bsx, bsy, qx, qy);
    out[i] = ((int64_t)d) << 5;
  }
}

/* ------------------------------------------------------------------------------------
 * BIDPartitionCost (JM/lencod/src/mv_search.c:1159-1250): motion cost of the bi-predictive direction of one partition.
 *   mcost  = weighted_cost(lambda_factor, mvd_bits)              (lcommon/inc/ifunctions.h:224: factor * bits, JCOST_CALC_SCALEUP)
 *   mb_pred: for every sub-block (h, v) of the region, luma_prediction(.., p_dir = 2, ..) (mc_prediction.c:144-236): both lists'
 *            blocks through OneComponentLumaPrediction (:117-136, ONE UMVLine4X origin clamp per sub-block and list), then
 *            bi_prediction ((a + b + 1) >> 1) or weighted_bi_prediction with wbp_weight / offsets (:207-213)
 *   mcost += distortion4x4 of every 4x4 block of the region, or distortion8x8 of every 8x8 block when Transform8x8Mode is on and
 *            blocktype <= 4 (select_distortion, me_distortion.c:148-166; each returns dist_scale(.) = . << 5)
 * mvd_bits (mv_bit_cost, mv_search.c:559-581) needs the encoder's motion field: it is an input here, as at the device boundary.
 * Job layout = b2me_bid_job of include/b2me.h.
 * ---------------------------------------------------------------------------------- */
typedef struct { int16_t mb_x, mb_y, blocktype, block8x8, ref_l0, ref_l1, mv_l0[4][2], mv_l1[4][2], weight_l0, weight_l1, offset_bi, reserved;
                 int32_t mvd_bits, lambda_factor; } OrcBidJob;
static const short ORC_BX0[5][4] = {{0,0,0,0}, {0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,2,0,2}};
static const short ORC_BY0[5][4] = {{0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,0,0,0}, {0,0,2,2}};
void orc_bid_partition_cost(void *h, int metric, int transform8x8, int wp, int denom, int n, const OrcBidJob *jobs, int64_t *out)
{
  const OrcFrame *f = (const OrcFrame *)h;
  const int Wp = f->W + 2 * PAD_X;
  int k;
  for (k = 0; k < n; k++) {
    const OrcBidJob *J = &jobs[k];
    const int bt = J->blocktype, pt = bt < 4 ? bt : 4;
    const int bx = ORC_BX0[pt][J->block8x8] << 2, by = ORC_BY0[pt][J->block8x8] << 2;
    const int w0 = ORC_BS[pt][0], h0 = ORC_BS[pt][1], sw = ORC_BS[bt][0], sh = ORC_BS[bt][1];
    const uint8_t *pl0 = orc_frame_planes(h, J->ref_l0), *pl1 = orc_frame_planes(h, J->ref_l1);
    uint8_t pred[16][16];
    int v, hh, x, y, sb = 0;
    int64_t mcost = (int64_t)J->lambda_factor * J->mvd_bits;
    for (v = 0; v < h0; v += sh)
      for (hh = 0; hh < w0; hh += sw, sb++) {
        const int qx = (J->mb_x + bx + hh) << 2, qy = (J->mb_y + by + v) << 2;
        const uint8_t *r0 = orc_umv_line4x(pl0, f->W, f->H, qy + J->mv_l0[sb][1], qx + J->mv_l0[sb][0]);
        const uint8_t *r1 = orc_umv_line4x(pl1, f->W, f->H, qy + J->mv_l1[sb][1], qx + J->mv_l1[sb][0]);
        for (y = 0; y < sh; y++)
          for (x = 0; x < sw; x++)
            pred[v + y][hh + x] = (uint8_t)orc_bi_pel(r0[(size_t)y * Wp + x], r1[(size_t)y * Wp + x], wp, J->weight_l0, J->weight_l1, J->offset_bi, denom);
      }
    { const int T = (transform8x8 && bt <= 4) ? 8 : 4; int16_t diff[64]; int i, j;
      for (v = 0; v < h0; v += T)
        for (hh = 0; hh < w0; hh += T) {
          for (j = 0; j < T; j++)
            for (i = 0; i < T; i++)
              diff[j * T + i] = (int16_t)((int)f->cur[(size_t)(J->mb_y + by + v + j) * f->W + J->mb_x + bx + hh + i] - (int)pred[v + j][hh + i]);
          mcost += orc_distortion(metric, T, diff);
        } }
    out[k] = mcost;
  }
}

/* ------------------------------------------------------------------------------------
 * list_prediction_cost, list 0 (JM/lencod/src/mode_decision.c:275-300) with update_mcost (:256-267) and ref_cost
 * (JM/lencod/inc/mv_search.h:114-131, refbits of mv_search.c:377-385 = the ue(v) length 2*floor(log2(ref+1))+1):
 * for the 21 (mode, block) entries of a macroblock -- mode 1 (1 block), 2 and 3 (2 blocks), 4..7 (the four 8x8 quadrants;
 * the motion cost of a quadrant is the SUM over its sub-partitions, PartitionMotionSearch mv_search.c:1601-1843) -- the
 * reference that minimises motion cost + lambda * refbits, first minimum in reference order.
 * cost [nmb][nrefs][41] (cost_sub of the search) -> best_ref [nmb][21], best_cost [nmb][21].
 * ---------------------------------------------------------------------------------- */
static const signed char ORC_ENTRY_PARTS[21][4] = {
  {0, -1, -1, -1}, {1, -1, -1, -1}, {2, -1, -1, -1}, {3, -1, -1, -1}, {4, -1, -1, -1},
  {5, -1, -1, -1}, {6, -1, -1, -1}, {7, -1, -1, -1}, {8, -1, -1, -1},
  {9, 11, -1, -1}, {10, 12, -1, -1}, {13, 15, -1, -1}, {14, 16, -1, -1},          /* 8x4: two rows of the quadrant */
  {17, 18, -1, -1}, {19, 20, -1, -1}, {21, 22, -1, -1}, {23, 24, -1, -1},          /* 4x8: two columns */
  {25, 26, 29, 30}, {27, 28, 31, 32}, {33, 34, 37, 38}, {35, 36, 39, 40}};         /* 4x4 */
int orc_refbits(int ref) { int b = 1, v = ref + 1; while (v > 1) { v >>= 1; b += 2; } return b; }
void orc_select_refs_list(int nmb, int nrefs, int lsize, const int64_t *cost, int ref_lambda, int8_t *best_ref, int64_t *best_cost);
void orc_select_refs(int nmb, int nrefs, const int64_t *cost, int ref_lambda, int8_t *best_ref, int64_t *best_cost)
{ orc_select_refs_list(nmb, nrefs, nrefs, cost, ref_lambda, best_ref, best_cost); }
/* either list of a B slice: lsize = listXsize[cur_list] bounds the loop (mode_decision.c:290) and ref_cost (mv_search.h:116) */
void orc_select_refs_list(int nmb, int nrefs, int lsize, const int64_t *cost, int ref_lambda, int8_t *best_ref, int64_t *best_cost)
{
  int mb, e, r, k;
  for (mb = 0; mb < nmb; mb++)
    for (e = 0; e < 21; e++) {
      int64_t bm = DISTBLK_MAX_ORC; int br = 0;
      for (r = 0; r < lsize; r++) {
        int64_t mc = 0;
        for (k = 0; k < 4 && ORC_ENTRY_PARTS[e][k] >= 0; k++) mc += cost[((size_t)mb * nrefs + r) * 41 + ORC_ENTRY_PARTS[e][k]];
        if (mc < bm) {                                   /* update_mcost */
          mc += lsize <= 1 ? 0 : (int64_t)ref_lambda * orc_refbits(r);
          if (mc < bm) { bm = mc; br = r; }
        }
      }
      best_ref[(size_t)mb * 21 + e] = (int8_t)br; best_cost[(size_t)mb * 21 + e] = bm;
    }
}

/* ------------------------------------------------------------------------------------
 * EPZS integer-pel search: EPZS_motion_estimation JM/lencod/src/me_epzs.c:54-407 and its sub-macroblock twin
 * EPZS_subMB_motion_estimation :417-750, restated over the job the C ABI defines (include/b2me.h b2me_epzs_job): what the two
 * functions read from encoder state -- predictor vectors, EPZSDetermineStopCriterion (me_epzs_common.c:1764-1780), prevSad,
 * the EPZSStructure pattern tables (:46-145) -- arrives as data; everything they compute is here:
 *   median candidate (:104-110), prevSad exit (:118), stop test and half-threshold exit (:133-156), predictor scan with the
 *   EPZSMap test, best / second-best bookkeeping (:213-252), [sub-MB: three-quarter exit :586], pattern choice (:262-285),
 *   refinement loop with pattern chaining (:290-345), prevSad exit (:351), the second-best round (:366-390).
 * The EPZSMap (visited in this call) is a byte map over the window here.
 * ---------------------------------------------------------------------------------- */
typedef struct { int16_t dx, dy, start_nmbr, next_points; } OrcEpzsPoint;
typedef struct { int32_t npoints, stop_search, next_last, next_pattern; OrcEpzsPoint pt[12]; } OrcEpzsPattern;
typedef struct {
  int16_t pos_x, pos_y, blocktype, ref, mv[2], pred[2], range[2], mv_range, flags;
  int32_t lambda_factor;
  int64_t stop0, stop, medthres, prev_sad;
  int32_t pred_first;
  int16_t npred[4], cond_host[4], fixed_edge, pat_init, pat_sd, pat_sq, pat_else, pat_dual, pad_;
} OrcEpzsJob;
typedef struct { int64_t cost; int16_t mv[2], early, npoints; } OrcEpzsResult;

typedef struct { const OrcFrame *f; const uint8_t *pl; const uint8_t *cur; const OrcEpzsJob *J; int bsx, bsy, npts; uint8_t *map; int mw, mh; } OrcEpzsCtx;
static int64_t oe_sad(OrcEpzsCtx *c, int tx, int ty)
{
  c->npts++;
  return ((int64_t)orc_sad(c->pl, c->f->W, c->f->H, c->cur, c->f->W, c->bsx, c->bsy, (c->J->pos_x << 2) + tx, (c->J->pos_y << 2) + ty)) << 5;
}
static int64_t oe_mvc(const OrcEpzsJob *J, int tx, int ty) { return (int64_t)J->lambda_factor * (orc_mvbits(tx - J->pred[0]) + orc_mvbits(ty - J->pred[1])); }
static int oe_in(const OrcEpzsJob *J, int tx, int ty) { return abs(tx - J->mv[0]) <= J->range[0] && abs(ty - J->mv[1]) <= J->range[1]; }
static int oe_visit(OrcEpzsCtx *c, int tx, int ty)       /* 1: first visit (now marked) */
{
  uint8_t *m = &c->map[(size_t)(ty - c->J->mv[1] + c->J->range[1]) * c->mw + (tx - c->J->mv[0] + c->J->range[0])];
  if (*m) return 0;
  *m = 1; return 1;
}
void orc_epzs_search(void *h, int njobs, const OrcEpzsJob *jobs, const int16_t *preds, int npats, const OrcEpzsPattern *pats, OrcEpzsResult *out)
{
  const OrcFrame *f = (const OrcFrame *)h; int ji;
  for (ji = 0; ji < njobs; ji++) {
    const OrcEpzsJob *J = &jobs[ji];
    OrcEpzsCtx c;
    int tx = J->mv[0], ty = J->mv[1], t2x = 0, t2y = 0, done = 0;
    int64_t minc;
    const int pgate = (J->flags & 1) != 0;
    c.f = f; c.J = J; c.pl = orc_frame_planes(h, J->ref); c.cur = f->cur + (size_t)J->pos_y * f->W + J->pos_x;
    c.bsx = ORC_BS[J->blocktype][0]; c.bsy = ORC_BS[J->blocktype][1]; c.npts = 0;
    c.mw = 2 * J->range[0] + 1; c.mh = 2 * J->range[1] + 1;
    c.map = (uint8_t *)calloc((size_t)c.mw * c.mh, 1);
    oe_visit(&c, tx, ty);
    minc = oe_mvc(J, tx, ty) + oe_sad(&c, tx, ty);
    if (pgate && J->prev_sad < (J->stop0 < minc ? J->stop0 : minc)) done = 1;
    if (!done && minc > J->stop0) {
      const int64_t stop = J->stop;
      if (minc < (stop >> 1)) done = 1;
      if (!done) {
        int checkMedian = 0, g, k, pi = J->pred_first;
        int64_t second = DISTBLK_MAX_ORC;
        int use[4];
        use[0] = 1; use[1] = J->cond_host[1] && minc > stop; use[2] = J->fixed_edge || (J->cond_host[2] && minc > 3 * stop);
        use[3] = (J->cond_host[3] & 1) && ((J->cond_host[3] & 2) || minc > 2 * stop);
        for (g = 0; g < 4; g++)
          for (k = 0; k < J->npred[g]; k++, pi++) {
            int px, py; int64_t mc;
            if (!use[g]) continue;
            px = (int16_t)(preds[2 * pi] & 0xFFFC); py = (int16_t)(preds[2 * pi + 1] & 0xFFFC);     /* set_integer_mv :40-44 */
            if (!oe_in(J, px, py) || !oe_visit(&c, px, py)) continue;
            mc = oe_mvc(J, px, py);
            if (mc >= second) continue;
            mc += oe_sad(&c, px, py);
            if (mc < minc) { t2x = tx; t2y = ty; tx = px; ty = py; second = minc; minc = mc; checkMedian = 1; }
            else if (mc < second) { t2x = px; t2y = py; second = mc; checkMedian = 1; }
          }
        if ((J->flags & 2) && minc < ((3 * stop) >> 2)) done = 1;
        if (!done && minc > stop) {
          int pat = J->pat_init, cx, cy, round;
          if (J->flags & 4) {
            if (minc < stop + ((3 * J->medthres) >> 1))
              pat = ((tx == 0 && ty == 0) || (abs(tx - J->mv[0]) < J->mv_range && abs(ty - J->mv[1]) < J->mv_range)) ? J->pat_sd : J->pat_sq;
            else pat = J->pat_else;
          }
          cx = tx; cy = ty;
          for (round = 0; round < 2; round++) {
            const OrcEpzsPattern *P = &pats[pat];
            int total = P->npoints, point = 0, pstop = 0, last = 0, dir = 0;
            do {
              int n = total;
              while (n-- > 0) {
                const int qx = cx + P->pt[point].dx, qy = cy + P->pt[point].dy;
                if (oe_in(J, qx, qy) && oe_visit(&c, qx, qy)) {
                  int64_t mc = oe_mvc(J, qx, qy);
                  if (mc < minc) { mc += oe_sad(&c, qx, qy); if (mc < minc) { tx = qx; ty = qy; minc = mc; dir = point; } }
                }
                if (++point >= P->npoints) point -= P->npoints;
              }
              if (last || (tx == cx && ty == cy)) { pstop = P->stop_search; P = &pats[P->next_pattern]; total = P->npoints; last = P->next_last; dir = 0; point = 0; }
              else { total = P->pt[dir].next_points; point = P->pt[dir].start_nmbr; cx = tx; cy = ty; }
            } while (pstop != 1);
            if (pgate && ((4 * J->prev_sad < minc) || ((3 * J->prev_sad < minc) && (J->prev_sad <= stop)))) { done = 1; break; }
            if (!(checkMedian && (J->flags & 8) && minc > stop)) break;
            if ((tx == 0 && ty == 0) || (tx == J->mv[0] && ty == J->mv[1]))
              pat = (abs(tx - J->mv[0]) < J->mv_range && abs(ty - J->mv[1]) < J->mv_range) ? J->pat_sd : J->pat_sq;
            else pat = J->pat_dual;
            cx = t2x; cy = t2y; checkMedian = 0;
          }
        }
      }
    }
    free(c.map);
    out[ji].cost = minc; out[ji].mv[0] = (int16_t)tx; out[ji].mv[1] = (int16_t)ty; out[ji].early = (int16_t)done; out[ji].npoints = (int16_t)c.npts;
  }
  (void)npats;
}

/* ------------------------------------------------------------------------------------
 * Luma + chroma prediction of 4:2:0 macroblocks from list 0, list 1 or both (mirrors b2me_mc_mb_dev):
 * luma_prediction JM/lencod/src/mc_prediction.c:144-236 (bi_prediction :82-99), chroma_prediction :469-566 with
 * OneComponentChromaPrediction4x4_regenerate :292-353 -- restated below as orc_chroma_sample: the reference's own arithmetic,
 * C division of the (possibly negative) eighth-sample coordinate included.
 * ---------------------------------------------------------------------------------- */
int orc_chroma_sample(const uint8_t *plane, int Wc, int Hc, int ii, int jj)
{
  const int x0 = ii / 8, x1 = (ii + 7) / 8, y0 = jj / 8, y1 = (jj + 7) / 8;        /* C division, as the reference (:338-341) */
  const int cx0 = x0 < 0 ? 0 : (x0 > Wc - 1 ? Wc - 1 : x0), cx1 = x1 < 0 ? 0 : (x1 > Wc - 1 ? Wc - 1 : x1);
  const int cy0 = y0 < 0 ? 0 : (y0 > Hc - 1 ? Hc - 1 : y0), cy1 = y1 < 0 ? 0 : (y1 > Hc - 1 ? Hc - 1 : y1);
  const int f1 = ii & 7, f0 = 8 - f1, g1 = jj & 7, g0 = 8 - g1;
  return (f0 * g0 * plane[(size_t)cy0 * Wc + cx0] + f1 * g0 * plane[(size_t)cy0 * Wc + cx1] +
          f0 * g1 * plane[(size_t)cy1 * Wc + cx0] + f1 * g1 * plane[(size_t)cy1 * Wc + cx1] + 32) / 64;
}
void orc_mc_mb(void *h, const uint8_t *mb_mode, const uint8_t *b8mode, const uint8_t *pdir, const int8_t *ref8, const int16_t *mv0, const int16_t *mv1,
               const uint8_t *curc /* [2][Hc][Wc] */, const uint8_t *refc /* [nrefs][2][Hc][Wc] */,
               uint8_t *orig_y, uint8_t *pred_y, uint8_t *orig_c, uint8_t *pred_c)
{
  OrcFrame *f = (OrcFrame *)h;
  const int mbw = f->W / 16, nmb = mbw * (f->H / 16), Wp = f->W + 2 * PAD_X, Wc = f->W / 2, Hc = f->H / 2;
  int m, k, i, j, L;
  for (m = 0; m < nmb; m++) {
    const int mbx = m % mbw, mby = m / mbw;
    for (k = 0; k < 16; k++) {
      const int bx = k & 3, by = k >> 2, qd = (by >> 1) * 2 + (bx >> 1), dir = pdir[m * 4 + qd];
      int p, ox, oy, acc[16] = {0};
      orc_block_partition(mb_mode[m], b8mode + m * 4, bx, by, &p, &ox, &oy);
      for (L = 0; L < 2; L++) {
        if (!(dir == 2 || dir == L)) continue;
        {
          const int r = ref8[(m * 2 + L) * 4 + qd];
          const int16_t *v = (L ? mv1 : mv0) + (((size_t)m * f->nrefs + r) * 41 + p) * 2;
          const uint8_t *ref = orc_umv_line4x(orc_frame_planes(h, r), f->W, f->H, 4 * (mby * 16 + oy) + v[1], 4 * (mbx * 16 + ox) + v[0]);
          for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) acc[i * 4 + j] += ref[(size_t)(4 * by - oy + i) * Wp + (4 * bx - ox + j)];
        }
      }
      for (i = 0; i < 4; i++)
        for (j = 0; j < 4; j++) {
          pred_y[((size_t)m * 16 + k) * 16 + i * 4 + j] = (uint8_t)(dir == 2 ? (acc[i * 4 + j] + 1) >> 1 : acc[i * 4 + j]);
          orig_y[((size_t)m * 16 + k) * 16 + i * 4 + j] = f->cur[(size_t)(mby * 16 + 4 * by + i) * f->W + mbx * 16 + 4 * bx + j];
        }
    }
    for (k = 0; k < 8; k++) {
      const int pl = k >> 2, cb = k & 3, cbx = cb & 1, cby = cb >> 1, dir = pdir[m * 4 + cb];
      for (j = 0; j < 4; j++)
        for (i = 0; i < 4; i++) {
          const int ci = 4 * cbx + i, cj = 4 * cby + j;
          int p, ox, oy, acc = 0;
          orc_block_partition(mb_mode[m], b8mode + m * 4, ci >> 1, cj >> 1, &p, &ox, &oy);
          for (L = 0; L < 2; L++) {
            if (!(dir == 2 || dir == L)) continue;
            {
              const int r = ref8[(m * 2 + L) * 4 + cb];
              const int16_t *v = (L ? mv1 : mv0) + (((size_t)m * f->nrefs + r) * 41 + p) * 2;
              acc += orc_chroma_sample(refc + ((size_t)r * 2 + pl) * Wc * Hc, Wc, Hc, 8 * (mbx * 8 + ci) + v[0], 8 * (mby * 8 + cj) + v[1]);
            }
          }
          pred_c[(((size_t)m * 2 + pl) * 4 + cb) * 16 + j * 4 + i] = (uint8_t)(dir == 2 ? (acc + 1) >> 1 : acc);
          orig_c[(((size_t)m * 2 + pl) * 4 + cb) * 16 + j * 4 + i] = curc[(size_t)pl * Wc * Hc + (size_t)(mby * 8 + cj) * Wc + mbx * 8 + ci];
        }
    }
  }
}
