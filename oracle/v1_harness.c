/* v1_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Drives the UNMODIFIED version1 fractal hot path -- compute.c (compute_rms, compute_rdSum,
 * compute_domain_Sum, compute_range_Sum) and block_enc.c (full_search, bound_chk,
 * changeReferenceFrame, encode_one_macroblock ...) compiled from /root/reference where they lie --
 * the way V1/src/code.c:256-270 and V1/src/image.c:411-458,1108-1133 do:
 *     start_oneframe -> compute_domain_Sum() on the C plane set -> compute_range_Sum() ->
 *     encode_one_macroblock -> full_search -> compute_rms -> compute_rdSum.
 * The tentative-definition globals of V1/inc/global.h resolve here under -fcommon.  The H/M/N
 * plane sets are allocated zero-filled and their sum tables are never computed, exactly like the
 * shipped program (SURVEY Q-F3: the interpolation and the H/M/N table builds are commented out,
 * V1/src/code.c:220-265); v1h_set_ref() can fill them for tests that want live planes.
 */
#include "windows.h"
#include "global.h"
#include "i_global.h"
#include "image.h"
#include "mbuffer.h"
#include "compute.h"
#include "block_enc.h"
#include <stdint.h>

StorablePicture *enc_picture;
short *spiral_search_x, *spiral_search_y, *spiral_hpel_search_x, *spiral_hpel_search_y;

static InputParameters h_input;
static ImageParameters h_img;
static int g_W, g_H;

static double **alloc2d(int h, int w)
{
  int i;
  double **rows = (double **)calloc((size_t)h, sizeof(double *));
  double *data = (double *)calloc((size_t)h * w, sizeof(double));   /* zero-initialised, like get_mem2Ddouble (V1/src/memalloc.c:60-73) */
  for (i = 0; i < h; i++) rows[i] = data + (size_t)i * w;
  return rows;
}
static byte **alloc2b(int h, int w)
{
  int i;
  byte **rows = (byte **)calloc((size_t)h, sizeof(byte *));
  byte *data = (byte *)calloc((size_t)h * w, 1);
  for (i = 0; i < h; i++) rows[i] = data + (size_t)i * w;
  return rows;
}
static byte ***alloc_uv(int h, int w)
{
  byte ***p = (byte ***)calloc(2, sizeof(byte **));
  p[0] = alloc2b(h, w); p[1] = alloc2b(h, w);
  return p;
}

#define SIZES(X, C, S, h, w) X(16, C, S, h, w) X(8, C, S, h, w) X(4, C, S, h, w) X(16_8, C, S, h, w) X(8_16, C, S, h, w) X(8_4, C, S, h, w) X(4_8, C, S, h, w)
#define ALLOC_REF(SZ, C, S, h, w)  sum_##SZ##C##_ref##S = alloc2d(h, w); sum2_##SZ##C##_ref##S = alloc2d(h, w);
#define ALLOC_ORG(SZ, C, S, h, w)  sum_##SZ##C##_org = alloc2d(h, w); sum2_##SZ##C##_org = alloc2d(h, w);
#define ALLOC_SET(S) SIZES(ALLOC_REF, , S, g_H, g_W) SIZES(ALLOC_REF, _U, S, g_H / 2, g_W / 2) SIZES(ALLOC_REF, _V, S, g_H / 2, g_W / 2)

int v1h_init(int W, int H, int search_range_, double tol16, double tol8, double tol4)
{
  if (g_W) return -1;            /* one geometry per process (globals), like the reference */
  g_W = W; g_H = H;
  input = &h_input; img = &h_img;
  memset(&h_input, 0, sizeof(h_input)); memset(&h_img, 0, sizeof(h_img));
  input->imagewidth = W; input->imageheight = H; input->search_range = search_range_;
  input->tol_16 = tol16; input->tol_8 = tol8; input->tol_4 = tol4;
  search_range = search_range_; search_mode = 0; num_regions = 1; region = obj = 0;
  img->frmWidthInMbs = W / 16; img->frmHeightInMbs = H / 16; img->width = W; img->height = H;
  img->width_cr = W / 2; img->height_cr = H / 2;
  img->mb_data = (Macroblock *)calloc((size_t)(W / 16) * (H / 16), sizeof(Macroblock));
  img->current_mb_nr = 0;
  imgY_org = alloc2b(H, W); imgUV_org = alloc_uv(H / 2, W / 2);
  imgY_ref = alloc2b(H, W); imgUV_ref = alloc_uv(H / 2, W / 2);
  imgY_ref_h = alloc2b(H, W); imgUV_ref_h = alloc_uv(H / 2, W / 2);
  imgY_ref_m = alloc2b(H, W); imgUV_ref_m = alloc_uv(H / 2, W / 2);
  imgY_ref_n = alloc2b(H, W); imgUV_ref_n = alloc_uv(H / 2, W / 2);
  ALLOC_SET() ALLOC_SET(_H) ALLOC_SET(_M) ALLOC_SET(_N)
  SIZES(ALLOC_ORG, , , g_H / 4, g_W / 4) SIZES(ALLOC_ORG, _U, , g_H / 8, g_W / 8) SIZES(ALLOC_ORG, _V, , g_H / 8, g_W / 8)
  /* enc_picture->mv / ref_pic_id as SetRefAndMotionVectors_fract indexes them (block_enc.c:216-497):
   * [list][x4][y4], U and V stored behind Y */
  enc_picture = (StorablePicture *)calloc(1, sizeof(StorablePicture));
  {
    int l, x, nx = W / 4 + W / 4 + 8, ny = H / 4 + H / 4 + 8;
    enc_picture->mv = (int ****)calloc(2, sizeof(int ***));
    enc_picture->ref_pic_id = (int64 ***)calloc(2, sizeof(int64 **));
    for (l = 0; l < 2; l++) {
      enc_picture->mv[l] = (int ***)calloc((size_t)nx, sizeof(int **));
      enc_picture->ref_pic_id[l] = (int64 **)calloc((size_t)nx, sizeof(int64 *));
      for (x = 0; x < nx; x++) {
        int y;
        enc_picture->mv[l][x] = (int **)calloc((size_t)ny, sizeof(int *));
        enc_picture->ref_pic_id[l][x] = (int64 *)calloc((size_t)ny, sizeof(int64));
        for (y = 0; y < ny; y++) enc_picture->mv[l][x][y] = (int *)calloc(2, sizeof(int));
      }
    }
  }
  currentVideo = 'C';
  changeReferenceFrame('C');
  return 0;
}

static void copy_plane(byte **dst, const uint8_t *src, int h, int w)
{
  int i;
  for (i = 0; i < h; i++) memcpy(dst[i], src + (size_t)i * w, (size_t)w);
}

/* current (range) frame: V1/src/image.c:458 */
void v1h_set_cur(const uint8_t *Y, const uint8_t *U, const uint8_t *V)
{
  copy_plane(imgY_org, Y, g_H, g_W);
  copy_plane(imgUV_org[0], U, g_H / 2, g_W / 2);
  copy_plane(imgUV_org[1], V, g_H / 2, g_W / 2);
  compute_range_Sum();
}

static char plane_letter(int which) { return which == 0 ? 'C' : which == 1 ? 'H' : which == 2 ? 'M' : 'N'; }

/* reference (domain) plane set `which` (0 C, 1 H, 2 M, 3 N); build_sums=1 runs compute_domain_Sum on it
 * (V1/src/code.c:256-257 does so for C only). */
void v1h_set_ref(int which, const uint8_t *Y, const uint8_t *U, const uint8_t *V, int build_sums)
{
  changeReferenceFrame(plane_letter(which));
  copy_plane(imgY_ref_temp, Y, g_H, g_W);
  copy_plane(imgUV_ref_temp[0], U, g_H / 2, g_W / 2);
  copy_plane(imgUV_ref_temp[1], V, g_H / 2, g_W / 2);
  if (build_sums) compute_domain_Sum();
  changeReferenceFrame('C');
}

/* one call of the reference full_search (V1/src/block_enc.c:1933) against plane set `which`.
 * xy is pre-zeroed by the caller as the reference's callers do (Q-F11). */
double v1h_full_search(int which, int bx, int by, int bsx, int bsy, int con, int *xy, double *so)
{
  TRANS_NODE t;
  double rms;
  memset(&t, 0, sizeof(t));
  t.x = xy[0]; t.y = xy[1];
  changeReferenceFrame(plane_letter(which));
  rms = full_search(bx, by, bsx, bsy, con, &t);
  changeReferenceFrame('C');
  xy[0] = t.x; xy[1] = t.y; so[0] = t.scale; so[1] = t.offset;
  return rms;
}

/* every range block of one (plane set, component, block size), raster order over the block grid */
void v1h_full_search_all(int which, int bsx, int bsy, int con, int32_t *xy, double *so, double *rms)
{
  int w = g_W / (con > 1 ? 2 : 1), h = g_H / (con > 1 ? 2 : 1), bx, by, k = 0;
  for (by = 0; by + bsy <= h; by += bsy)
    for (bx = 0; bx + bsx <= w; bx += bsx, k++) {
      int v[2] = {0, 0};
      rms[k] = v1h_full_search(which, bx, by, bsx, bsy, con, v, so + 2 * k);
      xy[2 * k] = v[0]; xy[2 * k + 1] = v[1];
    }
}

double v1h_compute_rms(int which, int bx, int by, int m, int n, int bsx, int bsy, int con, double *ab)
{
  double r;
  changeReferenceFrame(plane_letter(which));
  r = compute_rms(bx, by, m, n, &ab[0], &ab[1], bsx, bsy, con);
  changeReferenceFrame('C');
  return r;
}

/* read back a domain sum table of plane set C: sz 0..6 = 16,8,4,16_8,8_16,8_4,4_8; con 1..3; sq 0 sum, 1 sum2 */
#define PICK(SZ, C, S, h, w) if (k++ == sz) t = sq ? sum2_##SZ##C##_ref : sum_##SZ##C##_ref;
void v1h_domain_table(int sz, int con, int sq, double *out, int h, int w)
{
  double **t = NULL;
  int k = 0, i;
  if (con == 1) { SIZES(PICK, , , 0, 0) } else if (con == 2) { SIZES(PICK, _U, , 0, 0) } else { SIZES(PICK, _V, , 0, 0) }
  for (i = 0; i < h; i++) memcpy(out + (size_t)i * w, t[i], (size_t)w * sizeof(double));
}

/* The partition cascade of one macroblock (V1/src/block_enc.c:508): the TRANS_NODE tree flattened
 * to rows {level, bx, by, bsx, bsy, reference, partition, x, y, scale*100, offset} via a pre-order walk
 * is left to the caller through v1h_node_*; here only the root call. */
static TRANS_NODE *g_trans[2];
static TRANS_NODE *new_nodes(int n) { return (TRANS_NODE *)calloc((size_t)n, sizeof(TRANS_NODE)); }
static void give_children(TRANS_NODE *t, int depth)
{
  int i;
  if (depth == 0) return;
  t->next = new_nodes(4);
  for (i = 0; i < 4; i++) give_children(&t->next[i], depth - 1);
}
/* fresh (all-zero) TRANS_NODE trees: the reference keeps one tree array per component, the harness one in all */
void v1h_reset_trans(void) { g_trans[0] = g_trans[1] = NULL; }
int v1h_encode_mb(int CurMb, int con, int32_t *out, double *out_d, int maxn)
{
  /* out rows: block_type, partition, reference, x, y ; out_d rows: scale, offset ; pre-order, 3 levels */
  int n = 0, i, j, k;
  int nmb = (g_W / 16) * (g_H / 16);
  TRANS_NODE *root;
  if (!g_trans[0]) {
    for (i = 0; i < 2; i++) {
      g_trans[i] = new_nodes(nmb);
      for (j = 0; j < nmb; j++) give_children(&g_trans[i][j], 2);
    }
  }
  partition_length[0] = partition_length[1] = 0;
  img->current_mb_nr = 0;
  currentVideo = 'C';
  changeReferenceFrame('C');
  encode_one_macroblock(CurMb, g_trans, con);
  root = &g_trans[0][CurMb];
#define EMIT(t) do { if (n < maxn) { out[5 * n] = (t)->block_type; out[5 * n + 1] = (t)->partition; out[5 * n + 2] = (t)->reference; \
    out[5 * n + 3] = (t)->x; out[5 * n + 4] = (t)->y; out_d[2 * n] = (t)->scale; out_d[2 * n + 1] = (t)->offset; } n++; } while (0)
  EMIT(root);
  for (i = 0; i < 4; i++) {
    EMIT(&root->next[i]);
    for (k = 0; k < 4; k++) EMIT(&root->next[i].next[k]);
  }
  return n;
}

/* ---- F8: the fractal prediction of decode_one_macroblock (V1/src/block_dec.c:20, decode_block_rect :285,
 * decode_block_8 :760, decode_block_4 :978) from the TRANS_NODE trees the cascade left.  block_dec.c also holds the
 * intra / inverse-transform code of the decoder; the symbols only that code needs are stubbed here. ---- */
#include "block_dec.h"
StorablePicture *dec_picture;
const byte QP_SCALE_CR[52] = {0};
const int dequant_coef[6][4][4] = {{{0}}};
void error(char *text, int code) { (void)text; (void)code; abort(); }
void getNeighbour(int curr_mb_nr, int xN, int yN, int luma, PixelPos *pix) { (void)curr_mb_nr; (void)xN; (void)yN; (void)luma; (void)pix; abort(); }

/* reconstruct every macroblock of component con from the current trees: out = the component plane */
void v1h_decode_plane(int con, uint8_t *out)
{
  int w = con == 1 ? g_W : g_W / 2, h = con == 1 ? g_H : g_H / 2, mb, i;
  int nmb = con == 1 ? (g_W / 16) * (g_H / 16) : (g_W / 32) * (g_H / 32);
  byte **rec;
  if (!imgY_rec) { imgY_rec = alloc2b(g_H, g_W); imgUV_rec = alloc_uv(g_H / 2, g_W / 2); }
  rec = con == 1 ? imgY_rec : imgUV_rec[con - 2];
  for (i = 0; i < h; i++) memset(rec[i], 0, (size_t)w);
  currentVideo = 'C';
  for (mb = 0; mb < nmb; mb++) decode_one_macroblock(mb, g_trans, con);
  for (i = 0; i < h; i++) memcpy(out + (size_t)i * w, rec[i], (size_t)w);
}

#ifdef V1H_B2
/* libv1b2.so: the same harness over the same unmodified objects, full_search from integration/v1/b2fr_v1_shim.c */
void b2fr_v1_new_frame(const int have[4]);
long b2fr_v1_calls(void);
void v1h_b2_new_frame(const int *have) { b2fr_v1_new_frame(have); }
long v1h_b2_calls(void) { return b2fr_v1_calls(); }
#endif
