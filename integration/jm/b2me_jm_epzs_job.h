/* b2me_jm_epzs_job.h -- builds the b2me_epzs_job of one EPZS_motion_estimation / EPZS_subMB_motion_estimation call from the
 * encoder's state, with the reference's own helper functions (me_epzs_common.h:101-107).  Included by the drop-in shim
 * (b2me_jm_shim.c, -DB2ME_SHIM_EPZS) and by the oracle's boundary logger (oracle/jm_wrap_epzs.c), so that the job the GPU gets
 * and the job the golden capture records are built by the same code.  The includer provides b2_fail(const char *). */
#ifndef B2ME_JM_EPZS_JOB_H
#define B2ME_JM_EPZS_JOB_H
#include "me_epzs.h"
#include "me_epzs_common.h"

#define B2_EPZS_MAXPAT 8
static EPZSStructure *g_pat_ptr[B2_EPZS_MAXPAT];
static b2me_epzs_pattern g_pat[B2_EPZS_MAXPAT];
static int g_npat;

static int b2_epzs_pattern(EPZSStructure *s)            /* index of the serialised pattern (and of its successors) */
{
  int i, me;
  for (i = 0; i < g_npat; i++) if (g_pat_ptr[i] == s) return i;
  if (g_npat >= B2_EPZS_MAXPAT || s->searchPoints > 12) b2_fail("EPZS pattern table overflow");
  me = g_npat++;
  g_pat_ptr[me] = s;
  g_pat[me].npoints = s->searchPoints; g_pat[me].stop_search = s->stopSearch; g_pat[me].next_last = s->nextLast;
  for (i = 0; i < s->searchPoints; i++) {
    g_pat[me].pt[i].dx = s->point[i].motion.mv_x; g_pat[me].pt[i].dy = s->point[i].motion.mv_y;
    g_pat[me].pt[i].start_nmbr = (int16_t)s->point[i].start_nmbr; g_pat[me].pt[i].next_points = (int16_t)s->point[i].next_points;
  }
  g_pat[me].next_pattern = me;                           /* patterns may point at themselves */
  g_pat[me].next_pattern = b2_epzs_pattern(s->nextpattern);
  return me;
}

#define B2_EPZS_MAXPRED 256
/* ref_slot: the reference slot the caller's context holds listX[cur_list][ref] in.  pv: [B2_EPZS_MAXPRED][2].  Returns the number of predictors. */
static int b2_epzs_build_job(Macroblock *currMB, MotionVector *pred_mv, MEBlock *mv_block, int lambda_factor, int submb, int ref_slot,
                             b2me_epzs_job *Jp, int16_t *pv)
{
  Slice *currSlice = currMB->p_Slice;
  VideoParameters *p_Vid = currMB->p_Vid;
  InputParameters *p_Inp = currMB->p_Inp;
  EPZSParameters *p_EPZS = currSlice->p_EPZS;
  PicMotionParams **motion = p_Vid->enc_picture->mv_info;
  const int blocktype = mv_block->blocktype, list = mv_block->list, cur_list = list + currMB->list_offset;
  const short ref = mv_block->ref_idx;
  MotionVector *mv = &mv_block->mv[list];
  SearchWindow *searchRange = &mv_block->searchRange;
  StorablePicture *ref_picture = currSlice->listX[cur_list][ref];
  const distblk lambda_dist = weighted_cost(lambda_factor, submb ? 3 : 2);
  distblk *prevSad = &p_EPZS->distortion[cur_list][blocktype - 1][mv_block->pos_x2];
  SPoint *pt = p_EPZS->predictor->point;
  b2me_epzs_job J;
  int prednum = 5, n0, n1, i, ntot;
  short invalid_refs;

  memset(&J, 0, sizeof(J));
  J.pos_x = mv_block->pos_x; J.pos_y = mv_block->pos_y; J.blocktype = (int16_t)blocktype; J.ref = (int16_t)ref_slot;
  J.mv[0] = mv->mv_x; J.mv[1] = mv->mv_y; J.pred[0] = pred_mv->mv_x; J.pred[1] = pred_mv->mv_y;
  J.range[0] = (int16_t)searchRange->max_x; J.range[1] = (int16_t)searchRange->max_y;
  J.mv_range = submb ? 12 : 10;
  J.lambda_factor = lambda_factor;
  J.medthres = p_EPZS->medthres[blocktype];
  J.stop0 = J.medthres + lambda_dist;
  J.stop = EPZSDetermineStopCriterion(p_EPZS, prevSad, mv_block, lambda_dist);
  J.prev_sad = *prevSad;
  J.flags = (int16_t)(((ref > 0 && currSlice->structure == FRAME) ? B2ME_EPZS_REFGT0_FRAME : 0) | (submb ? B2ME_EPZS_EARLY34 : 0) |
                      (p_Inp->EPZSPattern != 0 ? B2ME_EPZS_ADAPT : 0));
  if (p_Inp->EPZSDual > 0 && (currSlice->slice_type == P_SLICE || (!submb && blocktype < 5))) J.flags |= B2ME_EPZS_DUAL;
  /* ---- predictor groups, in the reference's order (me_epzs.c:160-208 / :523-540) ---- */
  invalid_refs = EPZS_spatial_predictors(p_EPZS, mv_block, list, currMB->list_offset, ref, motion);
  if (p_Inp->EPZSSpatialMem) EPZS_spatial_memory_predictors(p_EPZS, mv_block, cur_list, &prednum, ref_picture->size_x >> 2);
  n0 = prednum;
  if (!submb && p_Inp->EPZSTemporal && blocktype < 5) {
    int first = prednum;
    EPZS_temporal_predictors(currMB, ref_picture, p_EPZS, mv_block, &first, 1, 0);       /* condition false: the co-located vector only */
    EPZS_temporal_predictors(currMB, ref_picture, p_EPZS, mv_block, &prednum, 0, 1);     /* condition true: + the neighbours (ref < 2) */
    n0 = first;
  }
  J.npred[0] = (int16_t)n0; J.npred[1] = (int16_t)(prednum - n0); J.cond_host[0] = 1; J.cond_host[1] = 1;
  n1 = prednum;
  if (!submb) {
    const int frame = currSlice->structure == FRAME && !currMB->list_offset;
    J.fixed_edge = (p_Inp->EPZSFixed == 3 && (currMB->mb_x == 0 || currMB->mb_y == 0)) ? 1 : 0;
    J.cond_host[2] = (((ref < 2 && blocktype < 4) || (ref < 1 && blocktype == 4) || (!frame && ref < 3)) &&
                      (p_Inp->EPZSFixed > 1 || (p_Inp->EPZSFixed && currSlice->slice_type == P_SLICE))) ? 1 : 0;
    if (J.fixed_edge || J.cond_host[2])
      EPZSWindowPredictors(mv, p_EPZS->predictor, &prednum,
                           (blocktype < 5) && (invalid_refs > 2) && (ref < 1 + (!frame)) ? p_EPZS->window_predictor_ext : p_EPZS->window_predictor);
  }
  J.npred[2] = (int16_t)(prednum - n1);
  n1 = prednum;
  if (currMB->mbAddrX != 0 && p_Inp->EPZSBlockType) {
    J.cond_host[3] = (int16_t)(1 | (ref == 0 ? 2 : 0));
    if (submb) EPZSBlockTypePredictors(currSlice, mv_block, pt, &prednum);
    else EPZSBlockTypePredictorsMB(currSlice, mv_block, pt, &prednum);
  }
  J.npred[3] = (int16_t)(prednum - n1);
  ntot = prednum;
  if (ntot > B2_EPZS_MAXPRED) b2_fail("EPZS predictor list overflow");
  for (i = 0; i < ntot; i++) { pv[2 * i] = pt[i].motion.mv_x; pv[2 * i + 1] = pt[i].motion.mv_y; }
  J.pred_first = 0;
  /* ---- patterns: the configured one, small diamond, square, the else branch of :281-284, the dual pattern ---- */
  J.pat_init = (int16_t)b2_epzs_pattern(p_EPZS->searchPattern);
  J.pat_sd = (int16_t)b2_epzs_pattern(p_Vid->sdiamond);
  J.pat_sq = (int16_t)b2_epzs_pattern(p_Vid->square);
  J.pat_else = submb ? J.pat_sq : ((blocktype > 4 || (ref > 0 && blocktype != 1)) ? J.pat_sq : J.pat_init);
  J.pat_dual = (int16_t)b2_epzs_pattern(p_EPZS->searchPatternD);
  *Jp = J;
  return ntot;
}
#endif
