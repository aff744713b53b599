/* jm_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C harness around the UNMODIFIED JM 18.5 objects compiled from /root/reference
 * (oracle/Makefile.jm -> oracle/_ref/libjmref.so).  It constructs the minimum JM state
 * (VideoParameters / InputParameters / Slice / Macroblock / MEBlock / StorablePicture)
 * and then calls the reference's own functions:
 *   init_motion_search_module   JM/lencod/src/mv_search.c:315   (spiral tables, mvbits)
 *   getSubImagesLuma            JM/lencod/src/img_luma.c:611    (16 quarter-pel planes)
 *   full_search_motion_estimation  JM/lencod/src/me_fullsearch.c:39
 *   sub_pel_motion_estimation      JM/lencod/src/me_fullsearch.c:186
 *   computeSAD / computeSATD       JM/lencod/src/me_distortion.c:349 / :745
 * so that the restated oracle (oracle/b2_oracle.c) and the CUDA path can be compared
 * against the reference itself on arbitrary seeded inputs.  The way the MEBlock is set
 * up follows init_mv_block / update_mv_block / get_original_block / get_search_range
 * (mv_search.c:698,675,786,70) and BlockMotionSearch (mv_search.c:858-976).
 *
 * Plain C ABI (ctypes-friendly): only ints, pointers to u8/i16/i32/i64.
 */
#include <limits.h>
#include <string.h>
#include <stdlib.h>
#include "global.h"
#include "image.h"
#include "memalloc.h"
#include "mbuffer.h"
#include "img_luma.h"
#include "me_distortion.h"
#include "me_fullsearch.h"
#include "mv_search.h"
#include "macroblock.h"

typedef struct {
  VideoParameters *p_Vid;
  InputParameters *p_Inp;
  Slice           *slice;
  Macroblock       mb;
  int W, H, nrefs, search_range;
  StorablePicture **refs;     /* nrefs pictures with 16 sub-planes each */
  StorablePicture **listX0;   /* list 0 */
  imgpel **cur;               /* current (original) luma */
  int wp_apply;               /* UseWeightedReferenceME && weighted_prediction: computeUniPred[x + 3] (the *WP variants) */
  short wp_weight[16], wp_offset[16];
} JMH;

/* metric codes follow JM: 0 SAD, 1 SSE, 2 SATD (lcommon/inc/types.h) */
void *jmh_create(int W, int H, int search_range, int nrefs,
                 int metric_f, int metric_h, int metric_q, int rdopt, int full_search_mode)
{
  JMH *h = (JMH *)calloc(1, sizeof(JMH));
  VideoParameters *p_Vid = (VideoParameters *)calloc(1, sizeof(VideoParameters));
  InputParameters *p_Inp = (InputParameters *)calloc(1, sizeof(InputParameters));
  int r;
  h->p_Vid = p_Vid; h->p_Inp = p_Inp; h->W = W; h->H = H; h->nrefs = nrefs;
  h->search_range = search_range;
  p_Vid->p_Inp = p_Inp;
  p_Inp->search_range[0] = p_Inp->search_range[1] = search_range;
  p_Inp->MEErrorMetric[F_PEL] = metric_f;
  p_Inp->MEErrorMetric[H_PEL] = metric_h;
  p_Inp->MEErrorMetric[Q_PEL] = metric_q;
  p_Inp->rdopt = rdopt;
  p_Inp->full_search = full_search_mode;   /* RestrictSearchRange, configfile.h:401 */
  p_Inp->SearchMode[0] = p_Inp->SearchMode[1] = FULL_SEARCH;
  p_Inp->IntraProfile = 1;                 /* skips fast-FS / UMHex table setup only */
  p_Vid->max_num_references = nrefs;
  p_Vid->max_pel_value_comp[0] = p_Vid->max_pel_value_comp[1] = p_Vid->max_pel_value_comp[2] = 255;
  p_Vid->max_imgpel_value = 255;
  p_Vid->bitdepth_luma = 8;
  p_Vid->nal_reference_idc = NALU_PRIORITY_HIGHEST;
  /* lencod.c:1795-1805 (init_global_buffers / coding params) */
  p_Vid->padded_size_x      = W + 2 * IMG_PAD_SIZE_X;
  p_Vid->padded_size_x_m8x8 = p_Vid->padded_size_x - BLOCK_SIZE_8x8;
  p_Vid->padded_size_x_m4x4 = p_Vid->padded_size_x - BLOCK_SIZE;
  get_mem2Dint_pad(&p_Vid->imgY_sub_tmp, H, W, IMG_PAD_SIZE_Y, IMG_PAD_SIZE_X);
  /* lencod.c:648-651: search window in quarter-pel units */
  p_Vid->searchRange.min_x = -(search_range << 2);
  p_Vid->searchRange.max_x =  (search_range << 2);
  p_Vid->searchRange.min_y = -(search_range << 2);
  p_Vid->searchRange.max_y =  (search_range << 2);

  init_motion_search_module(p_Vid, p_Inp);   /* the reference's own table builder */

  h->refs   = (StorablePicture **)calloc(nrefs, sizeof(StorablePicture *));
  h->listX0 = (StorablePicture **)calloc(nrefs + 1, sizeof(StorablePicture *));
  for (r = 0; r < nrefs; r++) {
    StorablePicture *s = alloc_storable_picture(p_Vid, FRAME, W, H, W / 2, H / 2);
    s->p_curr_img     = s->imgY;          /* mbuffer.c / image.c: luma plane selected */
    s->p_curr_img_sub = s->imgY_sub;
    s->p_img_sub[0]   = s->imgY_sub;
    h->refs[r] = s; h->listX0[r] = s;
  }
  get_mem2Dpel(&h->cur, H, W);
  h->slice = (Slice *)calloc(1, sizeof(Slice));
  h->slice->p_Vid = p_Vid; h->slice->p_Inp = p_Inp;
  h->slice->slice_type = P_SLICE;
  h->slice->listX[0] = h->listX0;
  h->slice->listXsize[0] = (char)nrefs;
  h->mb.p_Vid = p_Vid; h->mb.p_Inp = p_Inp; h->mb.p_Slice = h->slice; h->mb.list_offset = 0;
  return h;
}

int jmh_max_mvd(void *hh) { return ((JMH *)hh)->p_Vid->max_mvd; }
int jmh_mvbits(void *hh, int d) { return ((JMH *)hh)->p_Vid->mvbits[d]; }
void jmh_spiral(void *hh, int npos, short *out_xy_qpel)
{
  JMH *h = (JMH *)hh; int i;
  for (i = 0; i < npos; i++) {
    out_xy_qpel[2*i]   = h->p_Vid->spiral_qpel_search[i].mv_x;
    out_xy_qpel[2*i+1] = h->p_Vid->spiral_qpel_search[i].mv_y;
  }
}

/* upload reference luma (u8, W x H, stride W) and build the 16 sub-pel planes */
void jmh_set_ref(void *hh, int r, const unsigned char *luma)
{
  JMH *h = (JMH *)hh; int x, y;
  StorablePicture *s = h->refs[r];
  for (y = 0; y < h->H; y++)
    for (x = 0; x < h->W; x++)
      s->imgY[y][x] = luma[y * h->W + x];
  getSubImagesLuma(h->p_Vid, s);
}

void jmh_set_cur(void *hh, const unsigned char *luma)
{
  JMH *h = (JMH *)hh; int x, y;
  for (y = 0; y < h->H; y++)
    for (x = 0; x < h->W; x++)
      h->cur[y][x] = luma[y * h->W + x];
}

/* copy sub-plane [yy][xx] of reference r, padded size (H+40) x (W+64), as u8 */
void jmh_get_subplane(void *hh, int r, int yy, int xx, unsigned char *out)
{
  JMH *h = (JMH *)hh; int x, y, Wp = h->W + 2 * IMG_PAD_SIZE_X, Hp = h->H + 2 * IMG_PAD_SIZE_Y;
  imgpel **p = h->refs[r]->imgY_sub[yy][xx];
  for (y = 0; y < Hp; y++)
    for (x = 0; x < Wp; x++)
      out[y * Wp + x] = (unsigned char)p[y - IMG_PAD_SIZE_Y][x - IMG_PAD_SIZE_X];
}

/* explicit weighted prediction for the single-list search (PrepareMEParams, mv_search.c:183-188; init_mv_block
 * :731-748): per-reference luma weight / offset, slice-level log2 denominator; wp_luma_round as in
 * JM/lencod/src/weighted_prediction.c (denom ? 1 << (denom - 1) : 0). */
void jmh_set_weights(void *hh, int apply, int log_denom, const short *weight, const short *offset)
{
  JMH *h = (JMH *)hh; int r;
  h->wp_apply = apply;
  h->slice->luma_log_weight_denom = (short)log_denom;
  h->slice->wp_luma_round = log_denom ? 1 << (log_denom - 1) : 0;
  for (r = 0; r < h->nrefs && r < 16; r++) { h->wp_weight[r] = weight[r]; h->wp_offset[r] = offset[r]; }
}

static const short jmh_bs[8][2] = {{0,0},{16,16},{16,8},{8,16},{8,8},{8,4},{4,8},{4,4}};

static void jmh_setup_block(JMH *h, MEBlock *b, imgpel *orig, int pos_x, int pos_y, int blocktype, int ref)
{
  int j;
  memset(b, 0, sizeof(*b));
  b->p_Vid = h->p_Vid; b->p_Slice = h->slice;
  b->blocktype = (short)blocktype;
  b->blocksize_x = jmh_bs[blocktype][0];
  b->blocksize_y = jmh_bs[blocktype][1];
  b->pos_x = (short)pos_x; b->pos_y = (short)pos_y;
  b->pos_x2 = (short)(pos_x >> 2); b->pos_y2 = (short)(pos_y >> 2);
  b->pos_x_padded = (short)(pos_x << 2); b->pos_y_padded = (short)(pos_y << 2);
  b->list = 0; b->ref_idx = (char)ref;
  b->search_pos2 = 9; b->search_pos4 = 9;
  b->cost = INT_MAX;
  b->computePredFPel = h->p_Vid->computeUniPred[F_PEL];
  b->computePredHPel = h->p_Vid->computeUniPred[H_PEL];
  b->computePredQPel = h->p_Vid->computeUniPred[Q_PEL];
  if (h->wp_apply) {
    b->apply_weights = 1;
    b->computePredFPel = h->p_Vid->computeUniPred[F_PEL + 3];
    b->computePredHPel = h->p_Vid->computeUniPred[H_PEL + 3];
    b->computePredQPel = h->p_Vid->computeUniPred[Q_PEL + 3];
    b->weight_luma = h->wp_weight[ref]; b->offset_luma = h->wp_offset[ref];
  }
  get_search_range(b, h->p_Inp, (short)ref, blocktype);
  b->orig_pic = (imgpel **)malloc(sizeof(imgpel *));
  b->orig_pic[0] = orig;
  for (j = 0; j < b->blocksize_y; j++)
    memcpy(orig + j * b->blocksize_x, &h->cur[pos_y + j][pos_x], b->blocksize_x * sizeof(imgpel));
}

/* One call of the reference's integer full search followed (optionally) by its
 * sub-pel refinement, exactly as BlockMotionSearch sequences them (mv_search.c:960-976).
 *   center_mv[2] : in  = search centre (quarter-pel, relative MV), normally the
 *                        int-rounded predictor; out = integer-pel result
 *   out[0..1] = int-pel mv, out[2] = int cost lo, ... returned via pointers below. */
void jmh_block_search(void *hh, int pos_x, int pos_y, int blocktype, int ref,
                      const short *pred_mv, const short *center_mv,
                      const int *lambda_factor /*[3] F,H,Q*/, long long min_mcost_in,
                      int do_subpel, int test8x8,
                      short *mv_int_out, long long *cost_int_out,
                      short *mv_sub_out, long long *cost_sub_out)
{
  JMH *h = (JMH *)hh;
  MEBlock b; imgpel orig[256];
  MotionVector pred; distblk c;
  int lam[3];
  jmh_setup_block(h, &b, orig, pos_x, pos_y, blocktype, ref);
  b.test8x8 = test8x8;
  pred.mv_x = pred_mv[0]; pred.mv_y = pred_mv[1];
  b.mv[0].mv_x = center_mv[0]; b.mv[0].mv_y = center_mv[1];
  lam[0] = lambda_factor[0]; lam[1] = lambda_factor[1]; lam[2] = lambda_factor[2];
  c = full_search_motion_estimation(&h->mb, &pred, &b, (distblk)min_mcost_in, lam[F_PEL]);
  mv_int_out[0] = b.mv[0].mv_x; mv_int_out[1] = b.mv[0].mv_y; *cost_int_out = (long long)c;
  if (do_subpel) {
    if (!h->p_Vid->start_me_refinement_hp) c = DISTBLK_MAX;     /* mv_search.c:971-974 */
    c = do_subpel == 2 ? full_sub_pel_motion_estimation(&h->mb, &pred, &b, c, lam)      /* EPZSSubPelME == 2 */
                       : sub_pel_motion_estimation(&h->mb, &pred, &b, c, lam);
    mv_sub_out[0] = b.mv[0].mv_x; mv_sub_out[1] = b.mv[0].mv_y; *cost_sub_out = (long long)c;
  }
  free(b.orig_pic);
}

/* Batch form with the same semantics as the product's b2me_search_frame():
 * for every MB (raster), every ref, every one of the 41 partitions (blocktype 1..7,
 * partitions in raster order inside the MB) run jmh_block_search.
 *   pred  : [nmb][nrefs][41][2] i16 quarter-pel predictors
 *   center: same shape, search centres (relative MV, multiples of 4)
 *   outputs: mv_int/mv_sub [nmb][nrefs][41][2] i16, cost_int/cost_sub [nmb][nrefs][41] i64
 * mb_first/mb_count select a sub-range of MBs (bounded CPU-baseline samples). */
static const unsigned char part_bt[41] = {1, 2,2, 3,3, 4,4,4,4, 5,5,5,5,5,5,5,5, 6,6,6,6,6,6,6,6,
                                          7,7,7,7,7,7,7,7,7,7,7,7,7,7,7,7};
void jmh_partition_geometry(int p, int *bt, int *ox, int *oy)
{
  static const int first[8] = {0, 0, 1, 3, 5, 9, 17, 25};
  int t = part_bt[p], k = p - first[t], w = jmh_bs[t][0], hgt = jmh_bs[t][1], per_row = 16 / w;
  *bt = t; *ox = (k % per_row) * w; *oy = (k / per_row) * hgt;
}

void jmh_search_frame(void *hh, int mb_first, int mb_count,
                      const short *pred, const short *center, const int *lambda_factor,
                      int do_subpel,
                      short *mv_int, long long *cost_int, short *mv_sub, long long *cost_sub)
{
  JMH *h = (JMH *)hh;
  int mbw = h->W / 16, m, r, p;
  for (m = mb_first; m < mb_first + mb_count; m++) {
    int mbx = m % mbw, mby = m / mbw;
    for (r = 0; r < h->nrefs; r++)
      for (p = 0; p < 41; p++) {
        int bt, ox, oy; size_t i = ((size_t)m * h->nrefs + r) * 41 + p;
        jmh_partition_geometry(p, &bt, &ox, &oy);
        jmh_block_search(hh, mbx * 16 + ox, mby * 16 + oy, bt, r, pred + 2 * i, center + 2 * i,
                         lambda_factor, (long long)DISTBLK_MAX, do_subpel, 0,
                         mv_int + 2 * i, cost_int + i, mv_sub + 2 * i, cost_sub + i);
      }
  }
}

/* Direct distortion calls (absolute quarter-pel candidate, i.e. pos*4 + mv). */
long long jmh_sad(void *hh, int pos_x, int pos_y, int blocktype, int ref, int cand_x, int cand_y)
{
  JMH *h = (JMH *)hh; MEBlock b; imgpel orig[256]; MotionVector c; distblk d;
  jmh_setup_block(h, &b, orig, pos_x, pos_y, blocktype, ref);
  c.mv_x = (short)cand_x; c.mv_y = (short)cand_y;
  d = computeSAD(h->refs[ref], &b, DISTBLK_MAX, &c);
  free(b.orig_pic); return (long long)d;
}
long long jmh_satd(void *hh, int pos_x, int pos_y, int blocktype, int ref, int cand_x, int cand_y, int test8x8)
{
  JMH *h = (JMH *)hh; MEBlock b; imgpel orig[256]; MotionVector c; distblk d;
  jmh_setup_block(h, &b, orig, pos_x, pos_y, blocktype, ref);
  b.test8x8 = test8x8;
  c.mv_x = (short)cand_x; c.mv_y = (short)cand_y;
  d = computeSATD(h->refs[ref], &b, DISTBLK_MAX, &c);
  free(b.orig_pic); return (long long)d;
}
/* mode-decision distortions on a precomputed difference block (me_distortion.c:38-166): kind 0 SAD, 1 SSE, 2 SATD */
long long jmh_distortion(int kind, int n, short *diff)
{
  if (n == 4) return (long long)(kind == 0 ? distortion4x4SAD(diff, DISTBLK_MAX) : kind == 1 ? distortion4x4SSE(diff, DISTBLK_MAX) : distortion4x4SATD(diff, DISTBLK_MAX));
  return (long long)(kind == 0 ? distortion8x8SAD(diff, DISTBLK_MAX) : kind == 1 ? distortion8x8SSE(diff, DISTBLK_MAX) : distortion8x8SATD(diff, DISTBLK_MAX));
}
int jmh_hadamard4x4(short *diff) { return HadamardSAD4x4(diff); }
int jmh_hadamard8x8(short *diff) { return HadamardSAD8x8(diff); }

/* Bi-predictive search: the reference's full_search_bipred_motion_estimation (me_fullsearch.c:112) followed by
 * sub_pel_bipred_motion_estimation (:300), sequenced as BiPredBlockMotionSearch does (mv_search.c:1100-1126).
 * ref1 = searched picture (listX[list][ref]), ref2 = the static one (listX[list ^ 1][0]); the MEBlock carries the
 * weights PrepareBiPredMEParams (mv_search.c:196-262) would set.  job/result layouts = b2me_bipred_job / _result. */
typedef struct {
  long long min_mcost;
  short pos_x, pos_y, blocktype, ref1, ref2, search_range;
  short pred1[2], pred2[2], mv1[2], mv2[2];
  short weight1, weight2, offset_bi, reserved;
} JmhBiJob;
typedef struct { long long cost_int, cost_sub; short mv_int[2], mv_sub[2]; } JmhBiResult;

void jmh_bipred_search(void *hh, int njobs, const JmhBiJob *jobs, const int *lambda_factor, int do_subpel, int test8x8,
                       int wp, int log_denom, JmhBiResult *out)
{
  JMH *h = (JMH *)hh; int n;
  StorablePicture *l0[2], *l1[2];
  StorablePicture **save0 = h->slice->listX[0], **save1 = h->slice->listX[1];
  l0[1] = l1[1] = NULL;
  h->slice->luma_log_weight_denom = (short)log_denom;
  h->slice->wp_luma_round = log_denom ? 1 << (log_denom - 1) : 0;
  if (!h->p_Vid->mb_data) h->p_Vid->mb_data = (Macroblock *)calloc(1, sizeof(Macroblock));   /* sub_pel_bipred reads mb_data[mbAddrX].list_offset */
  for (n = 0; n < njobs; n++) {
    const JmhBiJob *J = &jobs[n];
    MEBlock b; imgpel orig[256]; MotionVector p1, p2, m1, m2; distblk c; int lam[3];
    jmh_setup_block(h, &b, orig, J->pos_x, J->pos_y, J->blocktype, 0);
    b.test8x8 = test8x8; b.ref_idx = 0;
    l0[0] = h->refs[J->ref1]; l1[0] = h->refs[J->ref2];
    h->slice->listX[0] = l0; h->slice->listX[1] = l1;
    b.apply_weights = wp;
    b.computeBiPredFPel = wp ? h->p_Vid->computeBiPred2[F_PEL] : h->p_Vid->computeBiPred1[F_PEL];   /* init_mv_block, mv_search.c:752-768 */
    b.computeBiPredHPel = wp ? h->p_Vid->computeBiPred2[H_PEL] : h->p_Vid->computeBiPred1[H_PEL];
    b.computeBiPredQPel = wp ? h->p_Vid->computeBiPred2[Q_PEL] : h->p_Vid->computeBiPred1[Q_PEL];
    b.weight1 = J->weight1; b.weight2 = J->weight2; b.offsetBi = J->offset_bi;
    p1.mv_x = J->pred1[0]; p1.mv_y = J->pred1[1]; p2.mv_x = J->pred2[0]; p2.mv_y = J->pred2[1];
    m1.mv_x = J->mv1[0]; m1.mv_y = J->mv1[1]; m2.mv_x = J->mv2[0]; m2.mv_y = J->mv2[1];
    lam[0] = lambda_factor[0]; lam[1] = lambda_factor[1]; lam[2] = lambda_factor[2];
    c = (distblk)J->min_mcost;
    if (J->search_range >= 0)
      c = full_search_bipred_motion_estimation(&h->mb, 0, &p1, &p2, &m1, &m2, &b, J->search_range << 2, (distblk)J->min_mcost, lam[F_PEL]);
    out[n].mv_int[0] = m1.mv_x; out[n].mv_int[1] = m1.mv_y; out[n].cost_int = (long long)c;
    out[n].mv_sub[0] = m1.mv_x; out[n].mv_sub[1] = m1.mv_y; out[n].cost_sub = (long long)c;
    if (do_subpel) {
      if (!h->p_Vid->start_me_refinement_hp && J->search_range >= 0) c = DISTBLK_MAX;   /* mv_search.c:1119-1120 */
      c = do_subpel == 2 ? full_sub_pel_bipred_motion_estimation(&h->mb, &b, 0, &p1, &p2, &m1, &m2, c, lam)
                         : sub_pel_bipred_motion_estimation(&h->mb, &b, 0, &p1, &p2, &m1, &m2, c, lam);
      out[n].mv_sub[0] = m1.mv_x; out[n].mv_sub[1] = m1.mv_y; out[n].cost_sub = (long long)c;
    }
    free(b.orig_pic);
  }
  h->slice->listX[0] = save0; h->slice->listX[1] = save1;
}

/* list_prediction_cost (mode_decision.c:275) for one (mode, block): p_Vid->motion_cost[mode][LIST_0][ref][block] = costs[ref] */
#include "mode_decision.h"
void jmh_list_prediction_cost(void *hh, int list, int mode, int block, int nrefs, const long long *costs, int ref_lambda, int *best_ref, long long *bmcost_out)
{
  JMH *h = (JMH *)hh; RD_PARAMS enc_mb; distblk bmcost[5]; char bref[2] = {0, 0}; int r;
  char save = h->slice->listXsize[list];
  memset(&enc_mb, 0, sizeof(enc_mb));
  enc_mb.lambda_mf[Q_PEL] = ref_lambda;
  if (!h->p_Vid->motion_cost) get_mem4Ddistblk(&h->p_Vid->motion_cost, 8, 2, h->p_Vid->max_num_references, 4);
  for (r = 0; r < nrefs; r++) h->p_Vid->motion_cost[mode][list][r][block] = (distblk)costs[r];
  h->slice->listXsize[list] = (char)nrefs;
  h->p_Vid->checkref = 0;
  bmcost[list] = DISTBLK_MAX;
  list_prediction_cost(&h->mb, list, block, mode, &enc_mb, bmcost, bref);
  h->slice->listXsize[list] = save;
  *best_ref = bref[list]; *bmcost_out = (long long)bmcost[list];
}
int jmh_refbits(void *hh, int r) { return ((JMH *)hh)->p_Vid->refbits[r]; }

/* One chroma 4x4 block through the UNMODIFIED OneComponentChromaPrediction4x4_regenerate (mc_prediction.c:292-353):
 * plane [Hc][Wc] = one chroma component of the reference picture, (pix_c_x, opix_c_y) = the macroblock's chroma origin,
 * (block_c_x, block_c_y) in {0, 4}^2, mv16 [4][4][2] = the vectors of the macroblock's sixteen luma 4x4 blocks. */
#include "mc_prediction.h"
void jmh_chroma_pred4x4(void *hh, int Wc, int Hc, const unsigned char *plane, int pix_c_x, int opix_c_y, int block_c_x, int block_c_y,
                        const short *mv16, unsigned char *out16)
{
  JMH *h = (JMH *)hh; VideoParameters *p_Vid = h->p_Vid;
  StorablePicture pic; Macroblock mb = h->mb; Macroblock mbd; seq_parameter_set_rbsp_t sps;
  MotionVector mvs[4][4], *rows[4]; imgpel mpred[16]; imgpel **uv2[2]; int x, y, i;
  memset(&pic, 0, sizeof(pic)); memset(&mbd, 0, sizeof(mbd)); memset(&sps, 0, sizeof(sps));
  sps.chroma_format_idc = 1;
  get_mem2Dpel(&uv2[0], Hc, Wc); uv2[1] = uv2[0];
  for (y = 0; y < Hc; y++) for (x = 0; x < Wc; x++) uv2[0][y][x] = plane[(size_t)y * Wc + x];
  pic.imgUV = uv2; pic.chroma_vector_adjustment = 0;
  for (y = 0; y < 4; y++) { rows[y] = mvs[y]; for (x = 0; x < 4; x++) { mvs[y][x].mv_x = mv16[(y * 4 + x) * 2]; mvs[y][x].mv_y = mv16[(y * 4 + x) * 2 + 1]; } }
  {
    seq_parameter_set_rbsp_t *save_sps = p_Vid->active_sps; Macroblock *save_md = p_Vid->mb_data;
    const int sx = p_Vid->mb_cr_size_x, sy = p_Vid->mb_cr_size_y, wc = p_Vid->width_cr, hc = p_Vid->height_cr;
    p_Vid->active_sps = &sps; p_Vid->mb_data = &mbd; p_Vid->mb_cr_size_x = p_Vid->mb_cr_size_y = 8; p_Vid->width_cr = Wc; p_Vid->height_cr = Hc;
    mb.mbAddrX = 0; mb.pix_c_x = (short)pix_c_x; mb.opix_c_y = (short)opix_c_y;
    OneComponentChromaPrediction4x4_regenerate(&mb, mpred, block_c_x, block_c_y, rows, &pic, 0);
    p_Vid->active_sps = save_sps; p_Vid->mb_data = save_md; p_Vid->mb_cr_size_x = sx; p_Vid->mb_cr_size_y = sy; p_Vid->width_cr = wc; p_Vid->height_cr = hc;
  }
  for (i = 0; i < 16; i++) out16[i] = (unsigned char)mpred[i];
  free_mem2Dpel(uv2[0]);
}
