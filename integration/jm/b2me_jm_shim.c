/* b2me_jm_shim.c -- the reference-side binding: JM 18.5's motion-search entry points served by
 * libb2me.so (hand-written sm_100a CUDA behind include/b2me.h).
 *
 * Link-level object replacement: JM's me_fullsearch.o is left out of the link and this
 * file defines all six symbols of that object (P and B slices),
 *     distblk full_search_motion_estimation(Macroblock*, MotionVector*, MEBlock*, distblk, int)
 *     distblk sub_pel_motion_estimation    (Macroblock*, MotionVector*, MEBlock*, distblk, int*)
 *     distblk full_search_bipred_motion_estimation(Macroblock*, int, MotionVector*, MotionVector*, MotionVector*, MotionVector*, MEBlock*, int, distblk, int)
 *     distblk sub_pel_bipred_motion_estimation    (Macroblock*, MEBlock*, int, MotionVector*, MotionVector*, MotionVector*, MotionVector*, distblk, int*)
 *     distblk full_sub_pel_motion_estimation / full_sub_pel_bipred_motion_estimation (same argument lists as the sub_pel pair)
 * with the exact signatures of JM/lencod/inc/me_fullsearch.h:20-25: nothing of me_fullsearch.c is linked.
 *
 * Ownership / threading follow the reference: the caller owns every buffer, the callee writes only
 * mv_block->mv[list] and returns the cost; one thread; state hangs off a process-wide context that
 * is created on first use from p_Vid / p_Inp and keyed on picture identity:
 *   - the current original picture is uploaded when p_Vid->enc_picture changes,
 *   - a reference picture is uploaded (and its 16 quarter-pel planes rebuilt on the GPU) the first
 *     time a (StorablePicture*, poc) pair is searched; slots are recycled least-recently-used.
 * Configurations the CUDA path does not cover (bit depth > 8, field/MBAFF pictures, weighted ME,
 * chroma ME, non-RDO (0,0) bias) stop the encoder through JM's own error()
 * -- there is no silent CPU fallback.
 *
 * Built only where the JM headers are available (it includes the reference's global.h).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mbuffer.h"
#include "me_fullsearch.h"
#include "mv_search.h"
#include "b2me.h"

#define B2_MAX_SLOTS 16

typedef struct {
  StorablePicture *pic;
  int poc;
  unsigned sig;          /* samples of the picture: JM re-uses StorablePicture addresses, and POC restarts at every IDR */
  int wp, weight, offset, denom;   /* planes stored weighted (UseWeightedReferenceME: computeSADWP & co read weighted samples) */
  long stamp;
} B2Slot;

static b2me_ctx *g_ctx;
static int g_W, g_H, g_nslots;
static B2Slot g_slot[B2_MAX_SLOTS];
static long g_clock;
static StorablePicture *g_cur_pic;
static int g_cur_poc = -0x7fffffff;
static unsigned g_cur_sig;
static unsigned char *g_stage;
static long g_calls_int, g_calls_sub, g_calls_bi, g_calls_dist, g_uploads, g_calls_epzs, g_points_epzs, g_calls_bidist, g_calls_blk;
static int g_in_bipred;

static void b2_fail(const char *what)
{
  char msg[600];
  snprintf(msg, sizeof(msg), "b2me shim: %s (%s)", what, g_ctx ? b2me_last_error(g_ctx) : "no context");
  error(msg, 500);
}

static void b2_report(void)
{
  if (getenv("B2ME_SHIM_VERBOSE"))
    fprintf(stderr, "b2me shim: %ld integer searches, %ld sub-pel refinements, %ld bi-predictive calls, %ld distortion calls, %ld bi-predictive distortion calls, %ld block distortions, %ld EPZS searches (%ld search points), %ld picture uploads, %lld kernel launches\n",
            g_calls_int, g_calls_sub, g_calls_bi, g_calls_dist, g_calls_bidist, g_calls_blk, g_calls_epzs, g_points_epzs, g_uploads, g_ctx ? (long long)b2me_launch_count(g_ctx) : 0LL);
  if (g_ctx) b2me_destroy(g_ctx);
  g_ctx = NULL;
}

static void b2_check_config(Macroblock *currMB, MEBlock *mv_block)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  Slice *currSlice = currMB->p_Slice;
  if (p_Vid->bitdepth_luma != 8) b2_fail("only 8-bit luma is supported");
  if (currSlice->structure != FRAME || currMB->list_offset != 0) b2_fail("field / MBAFF pictures are not supported");
  if (mv_block->list != 0 && mv_block->list != 1) b2_fail("only list 0 / list 1 frame references are supported");
  if (p_Inp->ChromaMEEnable || mv_block->ChromaMEEnable) b2_fail("ChromaMEEnable is not supported");
  if (!p_Inp->rdopt) b2_fail("RDOptimization=0 ((0,0)-bias path) is not supported");
  if (p_Inp->MEErrorMetric[F_PEL] != ERROR_SAD) b2_fail("MEDistortionFPel must be SAD");
  if (p_Inp->OnTheFlyFractMCP) b2_fail("OnTheFlyFractMCP must be 0");
}

/* Transform8x8Mode sets mv_block->test8x8 for 8x8 and larger blocks (mv_search.c:1630, 1770) and computeSATD then takes the 8x8
 * Hadamard (me_distortion.c:762).  The single-list sub-pel kernel has the 4x4 Hadamard only: stop instead of diverging. */
static void b2_check_test8x8(Macroblock *currMB, MEBlock *mv_block)
{
  InputParameters *p_Inp = currMB->p_Inp;
  if (mv_block->test8x8 && (p_Inp->MEErrorMetric[H_PEL] == ERROR_SATD || p_Inp->MEErrorMetric[Q_PEL] == ERROR_SATD))
    b2_fail("Transform8x8Mode with SATD sub-pel refinement (8x8 Hadamard, test8x8) is not supported by the single-list sub-pel search");
}

static void b2_ensure_ctx2(VideoParameters *p_Vid, InputParameters *p_Inp);
static void b2_ensure_ctx(Macroblock *currMB) { b2_ensure_ctx2(currMB->p_Vid, currMB->p_Inp); }
static void b2_ensure_ctx2(VideoParameters *p_Vid, InputParameters *p_Inp)
{
  int dev = 0, R = p_Inp->search_range[p_Vid->view_id];
  const char *e = getenv("B2ME_DEVICE");
  if (g_ctx) return;
  if (e) dev = atoi(e);
  g_W = p_Vid->width; g_H = p_Vid->height;
  g_nslots = imin(B2_MAX_SLOTS, imax(2, 2 * (p_Vid->max_num_references + 1)));
  if (b2me_create(&g_ctx, dev, g_W, g_H, g_nslots, R) != B2ME_OK) b2_fail("b2me_create failed");
  g_stage = (unsigned char *)malloc((size_t)g_W * g_H);
  if (!g_stage) no_mem_exit("b2me shim: staging plane");
  memset(g_slot, 0, sizeof(g_slot));
  atexit(b2_report);
}

/* imgpel (uint16 holding 8-bit samples, SURVEY Q-J3) -> packed bytes */
static void b2_narrow(imgpel **src)
{
  int x, y;
  for (y = 0; y < g_H; y++) {
    const imgpel *s = src[y];
    unsigned char *d = g_stage + (size_t)y * g_W;
    for (x = 0; x < g_W; x++) d[x] = (unsigned char)s[x];
  }
}

/* 256 samples spread over the picture: with (address, poc) a key that a re-allocated picture of another GOP does not repeat */
static unsigned b2_signature(imgpel **img)
{
  unsigned h = 2166136261u;
  int k;
  for (k = 0; k < 256; k++) {
    const int y = (int)(((long)k * 2654435761u) % (unsigned)g_H), x = (int)(((long)k * 40503u + 17) % (unsigned)g_W);
    h = (h ^ (unsigned)img[y][x]) * 16777619u;
  }
  return h;
}

static void b2_ensure_cur(VideoParameters *p_Vid)
{
  StorablePicture *enc = p_Vid->enc_picture;
  const unsigned sig = b2_signature(p_Vid->pCurImg);
  if (enc == g_cur_pic && enc->poc == g_cur_poc && sig == g_cur_sig) return;
  g_cur_sig = sig;
  b2_narrow(p_Vid->pCurImg);
  if (b2me_set_cur(g_ctx, g_stage, g_W) != B2ME_OK) b2_fail("b2me_set_cur failed");
  g_cur_pic = enc; g_cur_poc = enc->poc; g_uploads++;
  /* a new coded picture: reference slots whose picture is gone are simply aged out by the LRU */
}

static int b2_ref_slot_wp(StorablePicture *ref, int wp, int weight, int offset, int denom)
{
  int i, victim = 0;
  const unsigned sig = b2_signature(ref->imgY);
  if (!wp) weight = offset = denom = 0;
  for (i = 0; i < g_nslots; i++)
    if (g_slot[i].pic == ref && g_slot[i].poc == ref->poc && g_slot[i].sig == sig && g_slot[i].stamp && g_slot[i].wp == wp &&
        g_slot[i].weight == weight && g_slot[i].offset == offset && g_slot[i].denom == denom) { g_slot[i].stamp = ++g_clock; return i; }
  for (i = 1; i < g_nslots; i++)
    if (g_slot[i].stamp < g_slot[victim].stamp) victim = i;
  b2_narrow(ref->imgY);                 /* reconstructed (deblocked) luma; the GPU rebuilds getSubImagesLuma's planes */
  if (b2me_set_ref_weights(g_ctx, victim, wp, weight, offset, denom) != B2ME_OK) b2_fail("b2me_set_ref_weights failed");
  if (b2me_set_ref(g_ctx, victim, g_stage, g_W) != B2ME_OK) b2_fail("b2me_set_ref failed");
  g_slot[victim].pic = ref; g_slot[victim].poc = ref->poc; g_slot[victim].sig = sig; g_slot[victim].stamp = ++g_clock; g_uploads++;
  g_slot[victim].wp = wp; g_slot[victim].weight = weight; g_slot[victim].offset = offset; g_slot[victim].denom = denom;
  return victim;
}
static int b2_ref_slot(StorablePicture *ref) { return b2_ref_slot_wp(ref, 0, 0, 0, 0); }
/* the slot of a single-list search / distortion: weighted planes when the block asks for them (PrepareMEParams, mv_search.c:183-188) */
static int b2_ref_slot_for(Macroblock *currMB, MEBlock *mv_block, StorablePicture *ref)
{
  if (!mv_block->apply_weights) return b2_ref_slot(ref);
  return b2_ref_slot_wp(ref, 1, mv_block->weight_luma, mv_block->offset_luma, currMB->p_Slice->luma_log_weight_denom);
}

static void b2_params(InputParameters *p_Inp, b2me_search_params *P, int lam_f, int lam_h, int lam_q, distblk min_mcost)
{
  memset(P, 0, sizeof(*P));
  P->lambda_factor[0] = lam_f; P->lambda_factor[1] = lam_h; P->lambda_factor[2] = lam_q;
  P->restrict_mode = p_Inp->full_search;
  P->metric_h = p_Inp->MEErrorMetric[H_PEL]; P->metric_q = p_Inp->MEErrorMetric[Q_PEL];
  P->do_subpel = 0;
  P->min_mcost = (int64_t)min_mcost;
}

distblk full_search_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{
  Slice *currSlice = currMB->p_Slice;
  MotionVector *mv = &mv_block->mv[(short)mv_block->list];
  StorablePicture *ref_picture = currSlice->listX[mv_block->list + currMB->list_offset][mv_block->ref_idx];
  int search_range = imin(mv_block->searchRange.max_x, mv_block->searchRange.max_y) >> 2;
  b2me_search_params P;
  int16_t pm[2], cm[2], out[2];
  int64_t cost = 0;
  int slot;

  b2_ensure_ctx(currMB);
  b2_check_config(currMB, mv_block);
  b2_ensure_cur(currMB->p_Vid);
  slot = b2_ref_slot_for(currMB, mv_block, ref_picture);
  b2_params(currMB->p_Inp, &P, lambda_factor, lambda_factor, lambda_factor, min_mcost);
  pm[0] = pred_mv->mv_x; pm[1] = pred_mv->mv_y; cm[0] = mv->mv_x; cm[1] = mv->mv_y;
  if (b2me_block_search(g_ctx, mv_block->pos_x, mv_block->pos_y, mv_block->blocktype, slot, pm, cm, &P, search_range,
                        out, &cost, NULL, NULL) != B2ME_OK)
    b2_fail("b2me_block_search failed");
  g_calls_int++;
  mv->mv_x = out[0]; mv->mv_y = out[1];       /* centre + spiral[best_pos]; unchanged when best_pos == 0 */
  return (distblk)cost;
}

distblk sub_pel_motion_estimation(Macroblock *currMB, MotionVector *pred, MEBlock *mv_block, distblk min_mcost, int *lambda)
{
  Slice *currSlice = currMB->p_Slice;
  int list = mv_block->list;
  MotionVector *mv = &mv_block->mv[list];
  StorablePicture *ref_picture = currSlice->listX[list + currMB->list_offset][mv_block->ref_idx];
  b2me_search_params P;
  int16_t pm[2], in[2], out[2];
  int64_t cost = 0;
  int slot;

  b2_ensure_ctx(currMB);
  b2_check_config(currMB, mv_block);
  if (mv_block->search_pos2 != 9 || mv_block->search_pos4 != 9) b2_fail("SubPelSearch position counts other than 9/9 are not supported");
  b2_check_test8x8(currMB, mv_block);
  b2_ensure_cur(currMB->p_Vid);
  slot = b2_ref_slot_for(currMB, mv_block, ref_picture);
  b2_params(currMB->p_Inp, &P, lambda[F_PEL], lambda[H_PEL], lambda[Q_PEL], min_mcost);
  P.do_subpel = 1;
  pm[0] = pred->mv_x; pm[1] = pred->mv_y; in[0] = mv->mv_x; in[1] = mv->mv_y;
  if (b2me_block_subpel(g_ctx, mv_block->pos_x, mv_block->pos_y, mv_block->blocktype, slot, pm, in, &P, (int64_t)min_mcost,
                        out, &cost) != B2ME_OK)
    b2_fail("b2me_block_subpel failed");
  g_calls_sub++;
  mv->mv_x = out[0]; mv->mv_y = out[1];
  return (distblk)cost;
}

/* EPZSSubPelME == 2 (me_epzs_common.c:157) and SubPelME overrides: the brute-force 81-position refinement */
distblk full_sub_pel_motion_estimation(Macroblock *currMB, MotionVector *pred, MEBlock *mv_block, distblk min_mcost, int *lambda)
{
  Slice *currSlice = currMB->p_Slice;
  int list = mv_block->list;
  MotionVector *mv = &mv_block->mv[list];
  StorablePicture *ref_picture = currSlice->listX[list + currMB->list_offset][mv_block->ref_idx];
  b2me_search_params P;
  int16_t pm[2], in[2], out[2];
  int64_t cost = 0;
  int slot;

  b2_ensure_ctx(currMB);
  b2_check_config(currMB, mv_block);
  b2_check_test8x8(currMB, mv_block);
  b2_ensure_cur(currMB->p_Vid);
  slot = b2_ref_slot_for(currMB, mv_block, ref_picture);
  b2_params(currMB->p_Inp, &P, lambda[F_PEL], lambda[H_PEL], lambda[Q_PEL], min_mcost);
  P.do_subpel = 1; P.subpel_full = 1;
  pm[0] = pred->mv_x; pm[1] = pred->mv_y; in[0] = mv->mv_x; in[1] = mv->mv_y;
  if (b2me_block_subpel(g_ctx, mv_block->pos_x, mv_block->pos_y, mv_block->blocktype, slot, pm, in, &P, (int64_t)min_mcost,
                        out, &cost) != B2ME_OK)
    b2_fail("b2me_block_subpel (81 positions) failed");
  g_calls_sub++;
  mv->mv_x = out[0]; mv->mv_y = out[1];
  return (distblk)cost;
}

/* ---- B slices: the bi-predictive twins (JM/lencod/inc/me_fullsearch.h:21,24) on b2me_bipred_search ---------------- */
static distblk b2_bipred(Macroblock *currMB, int list, MotionVector *pred_mv1, MotionVector *pred_mv2, MotionVector *mv1,
                         MotionVector *mv2, MEBlock *mv_block, int search_range_pel, distblk min_mcost, int *lambda, int do_subpel /* 2: 81 positions */)
{
  Slice *currSlice = currMB->p_Slice;
  StorablePicture *ref_picture1 = currSlice->listX[list + currMB->list_offset][mv_block->ref_idx];
  StorablePicture *ref_picture2 = currSlice->listX[(list ^ 1) + currMB->list_offset][0];
  b2me_search_params P;
  b2me_bipred_job J;
  b2me_bipred_result R;

  b2_ensure_ctx(currMB);
  g_in_bipred = 1; b2_check_config(currMB, mv_block); g_in_bipred = 0;
  b2_ensure_cur(currMB->p_Vid);
  memset(&J, 0, sizeof(J));
  J.ref1 = (int16_t)b2_ref_slot(ref_picture1);
  J.ref2 = (int16_t)b2_ref_slot(ref_picture2);
  if (g_slot[J.ref1].pic != ref_picture1) J.ref1 = (int16_t)b2_ref_slot(ref_picture1);   /* ref2's upload may have recycled ref1's slot */
  if (J.ref1 == J.ref2 && ref_picture1 != ref_picture2) b2_fail("reference slots exhausted");
  b2_params(currMB->p_Inp, &P, lambda[F_PEL], lambda[H_PEL], lambda[Q_PEL], min_mcost);
  P.do_subpel = do_subpel ? 1 : 0; P.subpel_full = do_subpel == 2;
  J.min_mcost = (int64_t)min_mcost;
  J.pos_x = mv_block->pos_x; J.pos_y = mv_block->pos_y; J.blocktype = mv_block->blocktype;
  J.search_range = (int16_t)search_range_pel;
  J.pred1[0] = pred_mv1->mv_x; J.pred1[1] = pred_mv1->mv_y; J.pred2[0] = pred_mv2->mv_x; J.pred2[1] = pred_mv2->mv_y;
  J.mv1[0] = mv1->mv_x; J.mv1[1] = mv1->mv_y; J.mv2[0] = mv2->mv_x; J.mv2[1] = mv2->mv_y;
  J.weight1 = mv_block->weight1; J.weight2 = mv_block->weight2; J.offset_bi = mv_block->offsetBi;
  if (b2me_bipred_search(g_ctx, 1, &J, &P, mv_block->apply_weights ? 1 : 0, currSlice->luma_log_weight_denom, mv_block->test8x8, &R) != B2ME_OK)
    b2_fail("b2me_bipred_search failed");
  g_calls_bi++;
  if (do_subpel) { mv1->mv_x = R.mv_sub[0]; mv1->mv_y = R.mv_sub[1]; return (distblk)R.cost_sub; }
  mv1->mv_x = R.mv_int[0]; mv1->mv_y = R.mv_int[1];
  return (distblk)R.cost_int;
}

distblk full_search_bipred_motion_estimation(Macroblock *currMB, int list, MotionVector *pred_mv1, MotionVector *pred_mv2,
                                             MotionVector *mv1, MotionVector *mv2, MEBlock *mv_block, int search_range,
                                             distblk min_mcost, int lambda_factor)
{
  int lam[3];
  lam[F_PEL] = lam[H_PEL] = lam[Q_PEL] = lambda_factor;
  return b2_bipred(currMB, list, pred_mv1, pred_mv2, mv1, mv2, mv_block, search_range >> 2, min_mcost, lam, 0);
}

distblk sub_pel_bipred_motion_estimation(Macroblock *currMB, MEBlock *mv_block, int list, MotionVector *pred_mv1, MotionVector *pred_mv2,
                                         MotionVector *mv1, MotionVector *mv2, distblk min_mcost, int *lambda)
{
  if (mv_block->search_pos2 != 9 || mv_block->search_pos4 != 9) b2_fail("SubPelSearch position counts other than 9/9 are not supported");
  b2_check_test8x8(currMB, mv_block);
  return b2_bipred(currMB, list, pred_mv1, pred_mv2, mv1, mv2, mv_block, -1, min_mcost, lambda, 1);
}

distblk full_sub_pel_bipred_motion_estimation(Macroblock *currMB, MEBlock *mv_block, int list, MotionVector *pred_mv1, MotionVector *pred_mv2,
                                              MotionVector *mv1, MotionVector *mv2, distblk min_mcost, int *lambda)
{
  return b2_bipred(currMB, list, pred_mv1, pred_mv2, mv1, mv2, mv_block, -1, min_mcost, lambda, 2);
}

#ifdef B2ME_SHIM_DISTORTION
/* ---- every symbol of JM's me_distortion.o (JM/lencod/inc/me_distortion.h:21-91), that object left out of the link (lencod_b2d).
 * The computeSAD family at its own boundary -- the distortion pointers mv_block->computePredFPel / HPel / QPel,
 * computeBiPred1[] / computeBiPred2[] of EVERY search mode (EPZS, UMHex, ...) -- evaluates on the GPU, one candidate per call: a
 * functional demonstration of the boundary, not a fast path (a search that wants throughput hands whole predictor sets to
 * b2me_distortion_candidates or runs on the device altogether, b2me_epzs_search).  The weighted variants read planes that were
 * uploaded weighted (b2me_set_ref_weights); the bi-predictive ones go through b2me_bipred_distortion_candidates; the
 * mode decision's block distortions (distortion4x4 / 8x8 SAD / SSE / SATD, HadamardSAD4x4 / 8x8) through b2me_distortion_blocks.
 * The early exits of the reference return a value above the caller's bound; the full distortion returned here is above it too,
 * so every comparison falls the same way.  select_distortion and calcDifference are the reference's glue (function-pointer
 * selection, a subtraction loop) and are restated here because their object is gone. ---- */
#include "me_distortion.h"
static distblk b2_distortion(StorablePicture *ref1, MEBlock *mv_block, MotionVector *cand, int metric, int wp)
{
  b2me_candidate c;
  int64_t out = 0;
  b2_ensure_ctx2(mv_block->p_Vid, mv_block->p_Vid->p_Inp);
  if (mv_block->ChromaMEEnable) b2_fail("ChromaMEEnable is not supported");
  if (mv_block->p_Vid->bitdepth_luma != 8) b2_fail("only 8-bit luma is supported");
  b2_ensure_cur(mv_block->p_Vid);
  c.pos_x = mv_block->pos_x; c.pos_y = mv_block->pos_y; c.blocktype = mv_block->blocktype;
  c.ref = (int16_t)(wp ? b2_ref_slot_wp(ref1, 1, mv_block->weight_luma, mv_block->offset_luma, mv_block->p_Slice->luma_log_weight_denom) : b2_ref_slot(ref1));
  c.mv[0] = (int16_t)(cand->mv_x - mv_block->pos_x_padded); c.mv[1] = (int16_t)(cand->mv_y - mv_block->pos_y_padded);
  if (b2me_distortion_candidates(g_ctx, metric, mv_block->test8x8, 1, &c, &out) != B2ME_OK) b2_fail("b2me_distortion_candidates failed");
  g_calls_dist++;
  return (distblk)out;
}
distblk computeSAD(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{ (void)min_mcost; return b2_distortion(ref1, mv_block, cand, 0, 0); }
distblk computeSSE(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{ (void)min_mcost; return b2_distortion(ref1, mv_block, cand, 1, 0); }
distblk computeSATD(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{ (void)min_mcost; return b2_distortion(ref1, mv_block, cand, 2, 0); }
distblk computeSADWP(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{ (void)min_mcost; return b2_distortion(ref1, mv_block, cand, 0, 1); }
distblk computeSSEWP(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{ (void)min_mcost; return b2_distortion(ref1, mv_block, cand, 1, 1); }
distblk computeSATDWP(StorablePicture *ref1, MEBlock *mv_block, distblk min_mcost, MotionVector *cand)
{ (void)min_mcost; return b2_distortion(ref1, mv_block, cand, 2, 1); }

static distblk b2_bidistortion(StorablePicture *ref1, StorablePicture *ref2, MEBlock *mv_block, MotionVector *cand1, MotionVector *cand2, int metric, int wp)
{
  b2me_bipred_job J;
  int64_t out = 0;
  b2_ensure_ctx2(mv_block->p_Vid, mv_block->p_Vid->p_Inp);
  if (mv_block->ChromaMEEnable) b2_fail("ChromaMEEnable is not supported");
  b2_ensure_cur(mv_block->p_Vid);
  memset(&J, 0, sizeof(J));
  J.pos_x = mv_block->pos_x; J.pos_y = mv_block->pos_y; J.blocktype = mv_block->blocktype;
  J.ref1 = (int16_t)b2_ref_slot(ref1); J.ref2 = (int16_t)b2_ref_slot(ref2);
  if (g_slot[J.ref1].pic != ref1 || g_slot[J.ref1].wp) J.ref1 = (int16_t)b2_ref_slot(ref1);
  if (J.ref1 == J.ref2 && ref1 != ref2) b2_fail("reference slots exhausted");
  J.mv1[0] = (int16_t)(cand1->mv_x - mv_block->pos_x_padded); J.mv1[1] = (int16_t)(cand1->mv_y - mv_block->pos_y_padded);
  J.mv2[0] = (int16_t)(cand2->mv_x - mv_block->pos_x_padded); J.mv2[1] = (int16_t)(cand2->mv_y - mv_block->pos_y_padded);
  J.weight1 = mv_block->weight1; J.weight2 = mv_block->weight2; J.offset_bi = mv_block->offsetBi;
  if (b2me_bipred_distortion_candidates(g_ctx, metric, mv_block->test8x8, wp, wp ? mv_block->p_Slice->luma_log_weight_denom : 0, 1, &J, &out) != B2ME_OK)
    b2_fail("b2me_bipred_distortion_candidates failed");
  g_calls_bidist++;
  return (distblk)out;
}
distblk computeBiPredSAD1(StorablePicture *r1, StorablePicture *r2, MEBlock *b, distblk m, MotionVector *c1, MotionVector *c2) { (void)m; return b2_bidistortion(r1, r2, b, c1, c2, 0, 0); }
distblk computeBiPredSAD2(StorablePicture *r1, StorablePicture *r2, MEBlock *b, distblk m, MotionVector *c1, MotionVector *c2) { (void)m; return b2_bidistortion(r1, r2, b, c1, c2, 0, 1); }
distblk computeBiPredSSE1(StorablePicture *r1, StorablePicture *r2, MEBlock *b, distblk m, MotionVector *c1, MotionVector *c2) { (void)m; return b2_bidistortion(r1, r2, b, c1, c2, 1, 0); }
distblk computeBiPredSSE2(StorablePicture *r1, StorablePicture *r2, MEBlock *b, distblk m, MotionVector *c1, MotionVector *c2) { (void)m; return b2_bidistortion(r1, r2, b, c1, c2, 1, 1); }
distblk computeBiPredSATD1(StorablePicture *r1, StorablePicture *r2, MEBlock *b, distblk m, MotionVector *c1, MotionVector *c2) { (void)m; return b2_bidistortion(r1, r2, b, c1, c2, 2, 0); }
distblk computeBiPredSATD2(StorablePicture *r1, StorablePicture *r2, MEBlock *b, distblk m, MotionVector *c1, MotionVector *c2) { (void)m; return b2_bidistortion(r1, r2, b, c1, c2, 2, 1); }

/* the mode decision's distortions of one difference block (me_distortion.c:38-134) and the two Hadamards (:175-341) */
static int64_t b2_blockdist(short *diff, int kind, int n)
{
  int64_t out = 0;
  static int dev = -1;
  if (dev < 0) { const char *e = getenv("B2ME_DEVICE"); dev = e ? atoi(e) : 0; }
  if (b2me_distortion_blocks(dev, kind, n, 1, diff, &out) != B2ME_OK) b2_fail("b2me_distortion_blocks failed");
  g_calls_blk++;
  return out;
}
distblk distortion4x4SAD(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 0, 4); }
distblk distortion4x4SSE(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 1, 4); }
distblk distortion4x4SATD(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 2, 4); }
distblk distortion8x8SAD(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 0, 8); }
distblk distortion8x8SADthres(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 0, 8); }
distblk distortion8x8SSE(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 1, 8); }
distblk distortion8x8SATD(short *diff, distblk m) { (void)m; return (distblk)b2_blockdist(diff, 2, 8); }
int HadamardSAD4x4(short *diff) { return (int)(b2_blockdist(diff, 2, 4) >> 5); }      /* dist_scale = << 5 */
int HadamardSAD8x8(short *diff) { return (int)(b2_blockdist(diff, 2, 8) >> 5); }
void select_distortion(VideoParameters *p_Vid, InputParameters *p_Inp)
{
  const int m = p_Inp->ModeDecisionMetric;
  p_Vid->distortion4x4 = m == ERROR_SAD ? distortion4x4SAD : (m == ERROR_SSE ? distortion4x4SSE : distortion4x4SATD);
  p_Vid->distortion8x8 = m == ERROR_SAD ? distortion8x8SAD : (m == ERROR_SSE ? distortion8x8SSE : distortion8x8SATD);
}
void calcDifference(imgpel **origImg, int ox, int oy, imgpel **predImg, int px, int py, int width, int height, short *diff)
{
  int i, j;
  for (j = 0; j < height; j++) for (i = 0; i < width; i++) *diff++ = (short)(origImg[oy + j][ox + i] - predImg[py + j][px + i]);
}
#endif

#ifdef B2ME_SHIM_FULLFAST
/* ---- SearchMode 0 (fast full search): every symbol of JM's me_fullfast.o (JM/lencod/inc/me_fullfast.h:29-35), that object
 * left out of the link.  setup_fast_full_search -- the 4x4-SAD tables of one (macroblock, list, reference) around the rounded
 * 16x16 predictor and their tree sums -- is ONE GPU call (b2me_sad_table); the 41 calls of fast_full_search_motion_estimation
 * that follow scan that table with the caller's own predictor, the max_mvd guard included, exactly as
 * JM/lencod/src/me_fullfast.c:618-689 does.  The reference's BlockSAD arrays become one uint16 [41][max_pos] table per
 * (list, reference). ---- */
#include "me_fullfast.h"
#include "mv_prediction.h"

typedef struct {
  uint16_t *tab[2][B2_MAX_SLOTS];      /* [list][ref] -> [41][max_pos] */
  int done[2][B2_MAX_SLOTS];
  int range[2][B2_MAX_SLOTS];          /* max_search_range[list][ref] (pel) */
  MotionVector center[2][B2_MAX_SLOTS];
  int nref, max_pos;
} B2FF;
static B2FF g_ff;
static long g_calls_ff, g_setups_ff;

void initialize_fast_full_search(VideoParameters *p_Vid, InputParameters *p_Inp)
{
  int list, i, sr = p_Inp->search_range[p_Vid->view_id];
  memset(&g_ff, 0, sizeof(g_ff));
  g_ff.nref = imin(B2_MAX_SLOTS, p_Vid->max_num_references);
  g_ff.max_pos = (2 * sr + 1) * (2 * sr + 1);
  for (list = 0; list < 2; list++)
    for (i = 0; i < g_ff.nref; i++) {
      g_ff.tab[list][i] = (uint16_t *)calloc((size_t)B2ME_NPART * g_ff.max_pos, sizeof(uint16_t));
      if (!g_ff.tab[list][i]) no_mem_exit("b2me shim: fast full search tables");
      g_ff.range[list][i] = (p_Inp->full_search == 2 || i == 0) ? sr : sr / 2;       /* me_fullfast.c:113-128 */
    }
}
void clear_fast_full_search(VideoParameters *p_Vid)
{
  int list, i;
  (void)p_Vid;
  for (list = 0; list < 2; list++) for (i = 0; i < g_ff.nref; i++) { free(g_ff.tab[list][i]); g_ff.tab[list][i] = NULL; }
  if (getenv("B2ME_SHIM_VERBOSE")) fprintf(stderr, "b2me shim: %ld fast-full-search set-ups (GPU), %ld table scans\n", g_setups_ff, g_calls_ff);
}
void reset_fast_full_search(VideoParameters *p_Vid) { (void)p_Vid; memset(g_ff.done, 0, sizeof(g_ff.done)); }
void update_full_search_large_blocks(MEFullFast *p, int list, int refindex, int max_pos) { (void)p; (void)list; (void)refindex; (void)max_pos; }

void setup_fast_full_search(Macroblock *currMB, MEBlock *mv_block, int list)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  Slice *currSlice = currMB->p_Slice;
  short ref = mv_block->ref_idx;
  int range = g_ff.range[list][ref], rq = range << 2, slot;
  MotionVector pmv, *c = &g_ff.center[list][ref];
  PixelPos block[4];
  int16_t cm[2];

  b2_ensure_ctx(currMB);
  b2_check_config(currMB, mv_block);
  if (ref >= g_ff.nref) b2_fail("reference index beyond the fast-full-search tables");
  b2_ensure_cur(p_Vid);
  slot = b2_ref_slot(currSlice->listX[list + currMB->list_offset][ref]);
  /* search centre: the rounded predictor of the 16x16 block, clipped (me_fullfast.c:309-330) */
  get_neighbors(currMB, block, 0, 0, 16);
  currMB->GetMVPredictor(currMB, block, &pmv, ref, p_Vid->enc_picture->mv_info, list, 0, 0, 16, 16);
  c->mv_x = (short)(((pmv.mv_x + 2) >> 2) * 4);
  c->mv_y = (short)(((pmv.mv_y + 2) >> 2) * 4);
  if (!p_Inp->rdopt) { c->mv_x = (short)iClip3(-rq, rq, c->mv_x); c->mv_y = (short)iClip3(-rq, rq, c->mv_y); }
  c->mv_x = (short)iClip3(p_Vid->MaxHmvR[4] + rq, p_Vid->MaxHmvR[5] - rq, c->mv_x);
  c->mv_y = (short)iClip3(p_Vid->MaxVmvR[4] + rq, p_Vid->MaxVmvR[5] - rq, c->mv_y);
  cm[0] = c->mv_x; cm[1] = c->mv_y;
  if (b2me_sad_table(g_ctx, currMB->pix_x >> 4, currMB->opix_y >> 4, slot, cm, range, g_ff.tab[list][ref]) != B2ME_OK) b2_fail("b2me_sad_table failed");
  g_ff.done[list][ref] = 1; g_setups_ff++;
}

distblk fast_full_search_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  int search_range = imax(mv_block->searchRange.max_x, mv_block->searchRange.max_y) >> 2;
  int max_pos = (2 * search_range + 1) * (2 * search_range + 1), best_pos = 0, pos, part, q;
  int list = mv_block->list;
  short ref = mv_block->ref_idx;
  MotionVector cand = {0, 0}, *offset;
  int max_mvd = p_Vid->max_mvd - 1, tab_pos;
  const uint16_t *block_sad;
  static const int first[8] = {0, 0, 1, 3, 5, 9, 17, 25};
  static const int bw[8] = {0, 16, 16, 8, 8, 8, 4, 4}, bh[8] = {0, 16, 8, 16, 8, 4, 8, 4};

  if (!g_ff.done[list][ref]) currMB->p_SetupFastFullPelSearch(currMB, mv_block, list);
  offset = &g_ff.center[list][ref];
  tab_pos = (2 * g_ff.range[list][ref] + 1) * (2 * g_ff.range[list][ref] + 1);
  if (max_pos > tab_pos) b2_fail("fast full search: the block's range exceeds the table of its reference (the reference reads unwritten memory there; use RestrictSearchRange 2)");
  /* this library's partition number of (blocktype, block_x, block_y): raster order inside the blocktype */
  q = mv_block->blocktype;
  part = first[q] + (mv_block->block_y * 4 / bh[q]) * (16 / bw[q]) + mv_block->block_x * 4 / bw[q];
  block_sad = g_ff.tab[list][ref] + (size_t)part * tab_pos;
  g_calls_ff++;
  /* (the non-RDO (0,0) pre-check of :652-659 is unreachable: b2_check_config refuses RDOptimization 0) */
  for (pos = 0; pos < max_pos; pos++) {
    distblk mcost = dist_scale((distblk)block_sad[pos]);
    cand = add_MVs(p_Vid->spiral_qpel_search[pos], offset);
    if (mcost < min_mcost && GetMaxMVD(&cand, pred_mv) < max_mvd) {
      mcost += mv_cost(p_Vid, lambda_factor, &cand, pred_mv);
      if (mcost < min_mcost) { min_mcost = mcost; best_pos = pos; }
    }
  }
  mv_block->mv[list] = add_MVs(p_Vid->spiral_qpel_search[best_pos], offset);
  return min_mcost;
}
#endif

#ifdef B2ME_SHIM_EPZS
/* ---- SearchMode 3 (EPZS): the integer-pel stage on the GPU.  EPZS_motion_estimation (JM/lencod/src/me_epzs.c:54-407) and
 * EPZS_subMB_motion_estimation (:417-750) are made weak in the reference's me_epzs.o (objcopy on the object compiled from the
 * unmodified source) and defined here.  What stays on the host is the encoder's STATE, gathered with the reference's own helper
 * functions of me_epzs_common.o (declared in me_epzs_common.h:101-107): the predictor vectors, the stop criterion, prevSad,
 * the pattern tables (serialised from the live EPZSStructure objects).  The reference generates some predictor groups only
 * when the median candidate's cost exceeds a multiple of the stop criterion; here every group is generated, with the host part
 * of its condition, and the device applies the cost part (include/b2me.h b2me_epzs_job).  One launch per call (the median check,
 * the predictor scan and the whole pattern refinement) instead of one per search point. ---- */
#include "b2me_jm_epzs_job.h"

static distblk b2_epzs(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, int lambda_factor, int submb)
{
  Slice *currSlice = currMB->p_Slice;
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  EPZSParameters *p_EPZS = currSlice->p_EPZS;
  const int blocktype = mv_block->blocktype, list = mv_block->list, cur_list = list + currMB->list_offset;
  const short ref = mv_block->ref_idx;
  MotionVector *mv = &mv_block->mv[list];
  StorablePicture *ref_picture = currSlice->listX[cur_list][ref];
  distblk *prevSad = &p_EPZS->distortion[cur_list][blocktype - 1][mv_block->pos_x2];
  MotionVector *p_motion = NULL;
  b2me_epzs_job J;
  b2me_epzs_result R;
  int16_t pv[2 * B2_EPZS_MAXPRED];
  int ntot;

  b2_ensure_ctx(currMB);
  b2_check_config(currMB, mv_block);
  if (mv_block->apply_weights) b2_fail("weighted-prediction ME is not supported by the EPZS shim");
  b2_ensure_cur(p_Vid);
  if (p_Inp->EPZSSpatialMem) {
#if EPZSREF
    p_motion = &p_EPZS->p_motion[cur_list][ref][blocktype - 1][mv_block->block_y][mv_block->pos_x2];
#else
    p_motion = &p_EPZS->p_motion[cur_list][blocktype - 1][mv_block->block_y][mv_block->pos_x2];
#endif
  }
  ntot = b2_epzs_build_job(currMB, pred_mv, mv_block, lambda_factor, submb, b2_ref_slot(ref_picture), &J, pv);
  ++p_EPZS->BlkCount;
  if (p_EPZS->BlkCount == 0) ++p_EPZS->BlkCount;
  if (b2me_epzs_search(g_ctx, 1, &J, ntot, pv, g_npat, g_pat, &R) != B2ME_OK) b2_fail("b2me_epzs_search failed");
  g_calls_epzs++; g_points_epzs += R.npoints;
  if (!R.early && ((ref == 0) || (*prevSad > (distblk)R.cost))) *prevSad = (distblk)R.cost;
#if EPZSREF
  if (p_Inp->EPZSSpatialMem)
#else
  if (p_Inp->EPZSSpatialMem && ref == 0)
#endif
  { p_motion->mv_x = R.mv[0]; p_motion->mv_y = R.mv[1]; }
  mv->mv_x = R.mv[0]; mv->mv_y = R.mv[1];
  return (distblk)R.cost;
}
distblk EPZS_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{ (void)min_mcost; return b2_epzs(currMB, pred_mv, mv_block, lambda_factor, 0); }
distblk EPZS_subMB_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, distblk min_mcost, int lambda_factor)
{ (void)min_mcost; return b2_epzs(currMB, pred_mv, mv_block, lambda_factor, 1); }
#endif
