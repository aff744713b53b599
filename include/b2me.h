/* b2me.h -- C ABI of libb2me.so: the B200-native (sm_100a) block-matching hot path of the
 * joint fractal + H.264/AVC encoder in HuddsinYuan/--h.264-by-zhaodongyu.
 *
 * Plain C: raw pointers, sizes and POD structs only (no JM / version1 / torch types).
 * Every entry point returns 0 on success or a negative B2ME_E* code; nothing aborts.
 * Pointers are HOST pointers unless the function name ends in _dev (device pointers, and a
 * cudaStream_t passed as void*; those calls are asynchronous on that stream).
 *
 * Reference interfaces replaced (aliases: JM/ = 4.对比程序/jm18.5/JM, V1/ =
 * 2.论文程序/ZhangLing_Yu_version1/H264Fractal, both under the reference root):
 *   b2me_search_frame / b2me_block_search
 *        <- full_search_motion_estimation   JM/lencod/src/me_fullsearch.c:39-103
 *           sub_pel_motion_estimation       JM/lencod/src/me_fullsearch.c:186-289
 *           computeSAD / computeSATD        JM/lencod/src/me_distortion.c:349-426 / 745-825
 *           (the per-call sequencing of BlockMotionSearch, JM/lencod/src/mv_search.c:960-976)
 *   b2me_bipred_search
 *        <- full_search_bipred_motion_estimation / sub_pel_bipred_motion_estimation  JM/lencod/src/me_fullsearch.c:112-176 / 300-399
 *           computeBiPred{SAD,SSE,SATD}{1,2}   JM/lencod/src/me_distortion.c:525-737, 943-1182, 1353-1549
 *   b2me_set_ref (sub-pel plane build)
 *        <- getSubImagesLuma                JM/lencod/src/img_luma.c:611-680
 *   b2me_tq4x4
 *        <- forward4x4 / inverse4x4         JM/lcommon/src/transform.c:20-118
 *           quant_4x4_normal                JM/lencod/src/quant4x4_normal.c:39-115
 *           residual_transform_quant_luma_4x4  JM/lencod/src/block.c:660-724
 *           dct_luma (version1)             V1/src/block.c:836-1045
 *   b2fr_*  (fractal)
 *        <- compute_domain_Sum/compute_range_Sum  V1/src/compute.c:277,686
 *           compute_rdSum / compute_rms           V1/src/compute.c:192,6
 *           full_search / bound_chk               V1/src/block_enc.c:1933,2894
 */
#ifndef B2ME_H
#define B2ME_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2ME_OK            0
#define B2ME_EINVAL       -1   /* bad argument (null pointer, size not multiple of 16, ...) */
#define B2ME_ECUDA        -2   /* CUDA runtime error; see b2me_last_error() */
#define B2ME_ENOMEM       -3
#define B2ME_EUNSUPPORTED -4   /* valid in the reference but not implemented on this path */

#define B2ME_NPART        41   /* 1+2+2+4+8+8+16 partitions of a macroblock (JM block_size[], macroblock.h:58) */
#define B2ME_DISTBLK_MAX  (((int64_t)0x7fffffff) << 5)   /* JM defines.h:135 */

/* Partition numbering used by every array below (p = 0..40), blocktype = JM blocktype 1..7:
 *   p 0       16x16 (bt 1)
 *   p 1..2    16x8  (bt 2)  top, bottom
 *   p 3..4    8x16  (bt 3)  left, right
 *   p 5..8    8x8   (bt 4)  raster
 *   p 9..16   8x4   (bt 5)  raster (2 per row, 4 rows)
 *   p 17..24  4x8   (bt 6)  raster (4 per row, 2 rows)
 *   p 25..40  4x4   (bt 7)  raster                                            */

typedef struct b2me_ctx b2me_ctx;

typedef struct b2me_search_params {
  int32_t lambda_factor[3];  /* F_PEL, H_PEL, Q_PEL lambda_mf (LAMBDA_FACTOR, JM defines.h:130) */
  int32_t restrict_mode;     /* p_Inp->full_search (RestrictSearchRange): 0,1,2 (mv_search.c:70-92) */
  int32_t metric_h;          /* MEDistortionHPel: 0 SAD, 1 SSE, 2 SATD (computeSAD / computeSSE / computeSATD) */
  int32_t metric_q;          /* MEDistortionQPel: 0 SAD, 1 SSE, 2 SATD */
  int32_t do_subpel;         /* !DisableSubpelME */
  int32_t subpel_full;       /* 0: sub_pel_motion_estimation (9 half + 8 quarter); 1: full_sub_pel_motion_estimation,
                                the 81 quarter-pel positions (SubPelME when EPZSSubPelME == 2, me_fullsearch.c:409-469) */
  int64_t min_mcost;         /* initial bound handed to the integer search (B2ME_DISTBLK_MAX) */
} b2me_search_params;

/* ---- context ------------------------------------------------------------------------ */
/* width/height: coded luma size, multiples of 16.  search_range: SearchRange in pel (<=64). */
int  b2me_create(b2me_ctx **out, int device, int width, int height, int nrefs, int search_range);
void b2me_destroy(b2me_ctx *ctx);
const char *b2me_last_error(b2me_ctx *ctx);          /* thread-unsafe, like the reference */
int  b2me_version(void);

/* ---- pictures ----------------------------------------------------------------------- */
/* Current (original) luma, 8-bit.  JM stores imgpel=uint16 holding 8-bit samples; the shim
 * narrows on upload (SURVEY Q-J3). */
int b2me_set_cur(b2me_ctx *ctx, const uint8_t *luma, int stride);
int b2me_set_cur_dev(b2me_ctx *ctx, const uint8_t *luma_dev, int stride, void *stream);
/* Reference picture ref_idx of list 0: uploads the reconstructed luma and builds the 16
 * quarter-pel planes on the device (getSubImagesLuma). */
int b2me_set_ref(b2me_ctx *ctx, int ref_idx, const uint8_t *luma, int stride);
int b2me_set_ref_dev(b2me_ctx *ctx, int ref_idx, const uint8_t *luma_dev, int stride, void *stream);
/* MB-row bands (one picture across several GPUs): rebuild the planes of luma rows [row_first, row_first + row_count)
 * only -- the rows this GPU's band reads (its own rows plus the halo it received) -- from a full-geometry picture
 * buffer; the pad above / below goes with the first / last picture row, everything else keeps its old content.
 * A slot with weights applied (b2me_set_ref_weights) is stored weighted as a whole: a partial range returns
 * B2ME_EUNSUPPORTED for it. */
int b2me_set_ref_rows_dev(b2me_ctx *ctx, int ref_idx, const uint8_t *luma_dev, int stride, int row_first, int row_count, void *stream);
/* Explicit weighted prediction for the single-list search (UseWeightedReferenceME: computeSADWP / SATDWP / SSEWP,
 * JM/lencod/src/me_distortion.c:434-517, 833-935, 1262-1345; weights from PrepareMEParams, mv_search.c:183-188):
 * luma weight / offset of reference ref_idx and the slice's luma_log_weight_denom.  Takes effect at the NEXT
 * b2me_set_ref[_dev] of that slot (the planes the distortions read are stored weighted); apply = 0 turns it off. */
int b2me_set_ref_weights(b2me_ctx *ctx, int ref_idx, int apply, int weight, int offset, int log_weight_denom);
/* Read back sub-pel plane [yy][xx] (padded (H+40) x (W+64), tightly packed) -- parity tests. */
int b2me_get_subplane(b2me_ctx *ctx, int ref_idx, int yy, int xx, uint8_t *out);

/* ---- motion search ------------------------------------------------------------------ */
/* Whole-frame batch: for every MB (raster), ref and partition p run the reference's integer
 * full search and, if params->do_subpel, its half/quarter-pel refinement.
 *   pred, center : [nmb][nrefs][41][2] int16, quarter-pel MV predictor and search centre
 *                  (centre = relative MV, multiple of 4; JM derives it at mv_search.c:931-932)
 *   mv_int, mv_sub : same shape, best integer / final quarter-pel MV
 *   cost_int, cost_sub : [nmb][nrefs][41] int64 motion costs (SAD or SATD <<5 + lambda*bits)
 * mv_sub/cost_sub may be NULL when do_subpel == 0.  With do_subpel != 0, mv_int and cost_int may BOTH be NULL (host-pointer
 * call only): BlockMotionSearch hands only the refined vector and cost on (mv_search.c:960-976), and the integer
 * stage's 16 MB per 1080p x 4 refs picture then stay on the device. */
int b2me_search_frame(b2me_ctx *ctx, const int16_t *pred, const int16_t *center,
                      const b2me_search_params *params,
                      int16_t *mv_int, int64_t *cost_int, int16_t *mv_sub, int64_t *cost_sub);
int b2me_search_frame_dev(b2me_ctx *ctx, const int16_t *pred_dev, const int16_t *center_dev,
                          const b2me_search_params *params,
                          int16_t *mv_int_dev, int64_t *cost_int_dev,
                          int16_t *mv_sub_dev, int64_t *cost_sub_dev, void *stream);
/* Streams and the _dev entry points.  A context owns ONE work-item counter and ONE error flag: at most one search of a
 * context may be in flight at a time, whatever the stream (use one context per concurrent stream -- contexts are cheap next
 * to their plane sets -- as bench.py's end-to-end leg does).  Uploads (b2me_set_cur_dev / b2me_set_ref[_rows]_dev) and the
 * searches that read them must be issued on the same stream or be ordered by the caller; a host-pointer b2me_set_ref is
 * ordered before later work on any stream by the library.  The _dev searches never synchronise, so their device-side
 * input check (a search centre that is not a multiple of 4) is not reported by the call itself: b2me_check_errors
 * synchronises `stream`, returns B2ME_EINVAL if a search since the last check saw such a centre, and clears the flag. */
int b2me_check_errors(b2me_ctx *ctx, void *stream);
/* MB sub-range variant (MB-row bands / bounded samples): MBs [mb_first, mb_first+mb_count);
 * arrays are still indexed by absolute MB number. */
int b2me_search_mbs_dev(b2me_ctx *ctx, int mb_first, int mb_count,
                        const int16_t *pred_dev, const int16_t *center_dev,
                        const b2me_search_params *params,
                        int16_t *mv_int_dev, int64_t *cost_int_dev,
                        int16_t *mv_sub_dev, int64_t *cost_sub_dev, void *stream);

/* Compact whole-frame search (the end-to-end path of bench.py): what the reference's BlockMotionSearch receives and what its mode
 * decision consumes, nothing else crossing the bus.
 *   pred_mb   [nmb][nrefs][2] int16: ONE quarter-pel predictor per (MB, ref), shared by its 41 partitions; the integer search
 *             centre is derived on the device, ((p + 2) >> 2) * 4 (JM_INT_DIVIDE, mv_search.c:931-932)
 *   best_ref  [nmb][21] int8, best_cost [nmb][21] int32: list_prediction_cost (mode_decision.c:275-300) per (mode, block) entry as
 *             b2me_select_refs_dev defines them (costs above INT32_MAX -- a bound that was never beaten -- saturate)
 *   best_mv   [nmb][41][2] int16: the refined vector of every partition for the reference its entry chose
 * Host pointers (pinned for speed); one synchronisation; 4.3 + 2.2 MB per 1080p x 4 refs picture instead of 14.9 + 16.1. */
int b2me_search_frame_best(b2me_ctx *ctx, const int16_t *pred_mb, const b2me_search_params *params, int ref_lambda,
                           int8_t *best_ref, int32_t *best_cost, int16_t *best_mv);

/* Drop-in for ONE call of full_search_motion_estimation (+ sub_pel_motion_estimation): block
 * at luma (pos_x,pos_y), JM blocktype 1..7, reference ref_idx.  search_range_pel is the
 * block's own range (min(max_x,max_y)>>2 after get_search_range). */
int b2me_block_search(b2me_ctx *ctx, int pos_x, int pos_y, int blocktype, int ref_idx,
                      const int16_t pred_mv[2], const int16_t center_mv[2],
                      const b2me_search_params *params, int search_range_pel,
                      int16_t mv_int[2], int64_t *cost_int, int16_t mv_sub[2], int64_t *cost_sub);

/* Drop-in service of JM's FAST full search (SearchMode 0): the SAD tables setup_fast_full_search builds for one macroblock and
 * reference (JM/lencod/src/me_fullfast.c:269-608 luma unweighted branch, update_full_search_large_blocks :195-260): for every
 * position pos of the spiral around center_mv (relative MV, multiple of 4; the reference centres on the rounded 16x16 predictor)
 * and every partition p:  table[p * npos + pos] = SAD of partition p at center + spiral[pos],  npos = (2 * search_range_pel + 1)^2,
 * uint16.  The 41 table scans of fast_full_search_motion_estimation (:618-689), whose motion costs need predictors that exist only
 * one partition after the other, stay with the caller (integration/jm/b2me_jm_shim.c).  Host pointer, synchronous. */
int b2me_sad_table(b2me_ctx *ctx, int mb_x, int mb_y, int ref_idx, const int16_t center_mv[2], int search_range_pel, uint16_t *table);

/* Drop-in for ONE call of sub_pel_motion_estimation: mv_in = the block's current MV (relative,
 * quarter-pel), min_mcost = the bound exactly as the caller hands it over (BlockMotionSearch passes
 * DISTBLK_MAX when the metric changes between levels, mv_search.c:971-974). */
int b2me_block_subpel(b2me_ctx *ctx, int pos_x, int pos_y, int blocktype, int ref_idx,
                      const int16_t pred_mv[2], const int16_t mv_in[2], const b2me_search_params *params,
                      int64_t min_mcost, int16_t mv_out[2], int64_t *cost_out);

/* ---- bi-predictive block search (B slices) ------------------------------------------------- */
/* One job = one call of full_search_bipred_motion_estimation (JM/lencod/src/me_fullsearch.c:112-176) followed, when
 * params->do_subpel, by sub_pel_bipred_motion_estimation (:300-399), sequenced as BiPredBlockMotionSearch does
 * (mv_search.c:1100-1126).  The distortions are computeBiPredSAD1 / SSE1 / SATD1 (me_distortion.c:525, 1353, 943) or,
 * with apply_weights, computeBiPredSAD2 / SSE2 / SATD2 (:628, 1450, 1048); F_PEL metric = SAD, H_PEL / Q_PEL =
 * params->metric_h / metric_q.  ref1 is the reference slot of the SEARCHED list (listX[list][ref]), ref2 the slot that
 * holds listX[list ^ 1][0], whose block stays at mv2.  Slots are the ones b2me_set_ref fills (upload them unweighted:
 * the bi-predictive weights are applied to the pair of samples, not to a plane). */
typedef struct b2me_bipred_job {
  int64_t min_mcost;          /* incoming bound (DISTBLK_MAX in the first refinement iteration) */
  int16_t pos_x, pos_y;       /* luma position of the block */
  int16_t blocktype;          /* 1..7 */
  int16_t ref1, ref2;         /* reference slots */
  int16_t search_range;       /* pel: (BiPredMESearchRange) >> iteration_no, <= the context's search_range;
                                 -1: sub_pel_bipred_motion_estimation ALONE (mv1 any quarter-pel, min_mcost used as handed over) */
  int16_t pred1[2], pred2[2]; /* predictors of the two lists, quarter-pel */
  int16_t mv1[2];             /* search centre of the searched list (relative MV, multiple of 4) */
  int16_t mv2[2];             /* the other list's vector (relative MV, any quarter-pel) */
  int16_t weight1, weight2, offset_bi;   /* MEBlock weight1 / weight2 / offsetBi (PrepareBiPredMEParams, mv_search.c:196-262);
                                            read only with apply_weights */
  int16_t reserved;
} b2me_bipred_job;            /* 48 bytes */
typedef struct b2me_bipred_result {
  int64_t cost_int, cost_sub; /* returned motion costs of the two calls (cost_sub = cost_int when !do_subpel) */
  int16_t mv_int[2], mv_sub[2];   /* mv1 after the integer search / after the sub-pel refinement */
} b2me_bipred_result;         /* 24 bytes */
/* test8x8: MEBlock test8x8 (8x8 Hadamard in SATD).  apply_weights && test8x8 with a SATD metric returns
 * B2ME_EUNSUPPORTED: that branch of the reference reads past its source row (me_distortion.c:1167, SURVEY Q-J5). */
int b2me_bipred_search(b2me_ctx *ctx, int njobs, const b2me_bipred_job *jobs, const b2me_search_params *params,
                       int apply_weights, int luma_log_weight_denom, int test8x8, b2me_bipred_result *out);
int b2me_bipred_search_dev(b2me_ctx *ctx, int njobs, const b2me_bipred_job *jobs_dev, const b2me_search_params *params,
                           int apply_weights, int luma_log_weight_denom, int test8x8, b2me_bipred_result *out_dev, void *stream);

/* The bi-predictive distortions at their own boundary (mv_block->computeBiPred1[] / computeBiPred2[]): computeBiPredSAD1 / SSE1 /
 * SATD1 (unweighted average) and computeBiPredSAD2 / SSE2 / SATD2 (apply_weights: weight1, weight2, offset_bi of the record)
 * (JM/lencod/src/me_distortion.c:525-737, 943-1182, 1353-1549) for n records: position, block type, the two reference slots,
 * mv1 = candidate of the first picture, mv2 = candidate of the second (quarter-pel, relative); the other fields are ignored.
 * out = distortion << 5.  metric 0 SAD, 1 SSE, 2 SATD. */
int b2me_bipred_distortion_candidates(b2me_ctx *ctx, int metric, int test8x8, int apply_weights, int luma_log_weight_denom, int n,
                                      const b2me_bipred_job *cands, int64_t *out);

/* ---- motion cost of the bi-predictive direction in the mode decision ------------------------------- */
/* BIDPartitionCost (JM/lencod/src/mv_search.c:1159-1250; called from list_prediction_cost's BI_PRED branch, mode_decision.c:
 * 765, 774): for one partition of a B macroblock, weighted_cost(lambda_factor, mvd_bits) + the mode decision's distortion
 * (p_Vid->distortion4x4 / distortion8x8: select_distortion, me_distortion.c:148-166) of the residual against luma_prediction
 * with p_dir == 2 (mc_prediction.c:144-236: every sub-block of the partition fetched with ITS vector from both lists through
 * UMVLine4X, then bi_prediction, or weighted_bi_prediction with the record's weights when apply_weights).
 *   region     : parttype = blocktype < 4 ? blocktype : 4; origin (bx0, by0)[parttype][block8x8] * 4 inside the macroblock,
 *                size block_size[parttype]; sub-blocks of block_size[blocktype] in raster order inside the region (1, 1, 1, 1, 2, 2, 4)
 *   mv_l0/l1   : all_mv[LIST][ref][blocktype][by][bx] of those sub-blocks (quarter-pel, relative), at most four
 *   mvd_bits   : what mv_bit_cost (mv_search.c:559-581) adds up for both lists -- the caller's motion-vector predictors
 *                (GetMVPredictor on the encoder's motion field) stay on the host
 *   metric     : ModeDecisionMetric 0 SAD, 1 SSE, 2 SATD; transform8x8: Transform8x8Mode != 0 (8x8 blocks for blocktype <= 4)
 * The twin BPredPartitionCost (mv_search.c:589-700; mode_decision.c:366, 372) is the same computation on the vectors of the
 * bi-predictive motion search (bipred_mv[list] instead of all_mv; luma_prediction_bi mc_prediction.c:244-284): same record.
 * out = the returned distblk.  One launch for n partitions (one warp each). */
typedef struct b2me_bid_job {
  int16_t mb_x, mb_y;         /* luma position of the macroblock (pix_x, opix_y) */
  int16_t blocktype;          /* 1..7 */
  int16_t block8x8;           /* 0..3 */
  int16_t ref_l0, ref_l1;     /* reference slots of the context holding listX[0][cur_ref[0]] / listX[1][cur_ref[1]] */
  int16_t mv_l0[4][2], mv_l1[4][2];
  int16_t weight_l0, weight_l1, offset_bi, reserved;   /* wbp_weight[0/1][ref0][ref1][0], (wp_offset0 + wp_offset1 + 1) >> 1; read only with apply_weights */
  int32_t mvd_bits, lambda_factor;
} b2me_bid_job;               /* 60 bytes */
int b2me_bid_partition_cost(b2me_ctx *ctx, int metric, int transform8x8, int apply_weights, int luma_log_weight_denom, int n,
                            const b2me_bid_job *jobs, int64_t *out);

/* ---- distortion at explicit candidates (the computeSAD family at its own boundary) ---------------------------- */
/* computeSAD / computeSSE / computeSATD (JM/lencod/src/me_distortion.c:349-426, 1190-1255, 745-825; the WP variants
 * when the slot was uploaded with b2me_set_ref_weights) for n independent (block, reference slot, vector) triples:
 * what mv_block->computePredFPel / HPel / QPel (JM/lencod/inc/global.h:316-318) return for `cand = block position +
 * mv` with min_mcost = DISTBLK_MAX, i.e. dist_scale(distortion) = distortion << 5, no motion-vector cost.  For searches
 * whose control flow stays in the reference's C (EPZS_motion_estimation, me_epzs.c:54; UMHEX): one call per predictor
 * set or refinement pattern.  metric 0 SAD, 1 SSE, 2 SATD; test8x8: 8x8 Hadamard (blocktypes 1..4). */
typedef struct b2me_candidate {
  int16_t pos_x, pos_y;       /* luma position of the block */
  int16_t blocktype;          /* 1..7 */
  int16_t ref;                /* reference slot */
  int16_t mv[2];              /* vector relative to the block position, quarter-pel */
} b2me_candidate;             /* 12 bytes */
int b2me_distortion_candidates(b2me_ctx *ctx, int metric, int test8x8, int n, const b2me_candidate *cands, int64_t *out);
int b2me_distortion_candidates_dev(b2me_ctx *ctx, int metric, int test8x8, int n, const b2me_candidate *cands_dev,
                                   int64_t *out_dev, void *stream);

/* ---- EPZS integer-pel search on the device --------------------------------------------------------- */
/* One job = one call of EPZS_motion_estimation (JM/lencod/src/me_epzs.c:54-407) or EPZS_subMB_motion_estimation (:417-750):
 * the median check and its early exits, the predictor scan with best / second-best bookkeeping, the pattern refinement
 * (small diamond / square / configured pattern, pattern chaining through nextpattern, the second-best "dual" round) and the
 * prevSad exits run on the GPU, one warp per job, every distortion = computeSAD at an integer position (<< 5), every
 * vector cost = lambda_factor * (mvbits(dx) + mvbits(dy)).  What stays with the caller is its STATE: the predictor values
 * (spatial / memory / temporal / window / block-type predictors are functions of the neighbouring blocks' final vectors and of
 * per-line memories), the stop criterion of EPZSDetermineStopCriterion (me_epzs_common.c:1764-1780) and prevSad; the pattern
 * tables are handed over as data (the caller serialises its EPZSStructure objects, me_epzs_common.c:46-145).
 * Predictor groups, in scan order: 0 always; 1 (temporal neighbours, me_epzs_common.c:1551) when cond_host[1] and
 * min_mcost > stop; 2 (window predictors, me_epzs.c:141-150) when fixed_edge or (cond_host[2] and min_mcost > 3 * stop);
 * 3 (block-type predictors, :156-159) when cond_host[3] & 1 and (cond_host[3] & 2 [ref == 0] or min_mcost > 2 * stop) --
 * min_mcost = the median candidate's cost, as in the reference. */
typedef struct b2me_epzs_point { int16_t dx, dy, start_nmbr, next_points; } b2me_epzs_point;
typedef struct b2me_epzs_pattern {
  int32_t npoints, stop_search, next_last, next_pattern;   /* next_pattern: index into the patterns array */
  b2me_epzs_point pt[12];
} b2me_epzs_pattern;                                        /* 112 bytes */
#define B2ME_EPZS_REFGT0_FRAME 1    /* ref > 0 && structure == FRAME: the prevSad exits (me_epzs.c:118, 351) apply */
#define B2ME_EPZS_EARLY34      2    /* sub-macroblock variant: stop after the predictors when min_mcost < 3 * stop >> 2 (:586) */
#define B2ME_EPZS_ADAPT        4    /* EPZSPattern != 0: choose the refinement pattern from the cost (:268-284) */
#define B2ME_EPZS_DUAL         8    /* the second-best round is allowed (EPZSDual > 0 and the slice / block-type condition, :366-368) */
typedef struct b2me_epzs_job {
  int16_t pos_x, pos_y;       /* luma position of the block */
  int16_t blocktype;          /* 1..7 */
  int16_t ref;                /* reference slot */
  int16_t mv[2];              /* search centre at entry (quarter-pel) */
  int16_t pred[2];            /* motion vector predictor (quarter-pel) */
  int16_t range[2];           /* searchRange.max_x / max_y (quarter-pel) */
  int16_t mv_range;           /* 10 (macroblock variant) / 12 (sub-macroblock variant) */
  int16_t flags;              /* B2ME_EPZS_* */
  int32_t lambda_factor;
  int64_t stop0;              /* medthres[blocktype] + lambda_dist */
  int64_t stop;               /* EPZSDetermineStopCriterion */
  int64_t medthres;
  int64_t prev_sad;           /* *prevSad at entry */
  int32_t pred_first;         /* first predictor of this job in the predictor array */
  int16_t npred[4];           /* predictors per group */
  int16_t cond_host[4];
  int16_t fixed_edge;
  int16_t pat_init, pat_sd, pat_sq, pat_else, pat_dual;   /* pattern indices: configured, small diamond, square, the else branch of :281-284, searchPatternD */
  int16_t pad_;
} b2me_epzs_job;
typedef struct b2me_epzs_result {
  int64_t cost;               /* the function's return value */
  int16_t mv[2];              /* *mv at return */
  int16_t early;              /* 1: returned through an early exit (prevSad is not updated), 0: the final return */
  int16_t npoints;            /* search points evaluated (distortion calls) */
} b2me_epzs_result;
/* preds: [npreds_total][2] int16 (the raw predictor vectors; set_integer_mv is applied on the device). */
int b2me_epzs_search(b2me_ctx *ctx, int njobs, const b2me_epzs_job *jobs, int npreds_total, const int16_t *preds,
                     int npatterns, const b2me_epzs_pattern *patterns, b2me_epzs_result *out);
int b2me_epzs_search_dev(b2me_ctx *ctx, int njobs, const b2me_epzs_job *jobs_dev, const int16_t *preds_dev,
                         int npatterns, const b2me_epzs_pattern *patterns_dev, b2me_epzs_result *out_dev, void *stream);

/* ---- in-loop deblocking filter of a frame picture ------------------------------------------------------ */
/* DeblockFrame (JM/lencod/src/loopFilter.c:63-111, DeblockMb :196-377) with the non-MBAFF functions of
 * JM/lencod/src/loop_filter_normal.c (GetStrengthVer / Hor :52-283, EdgeLoopLuma / Chroma Ver / Hor :285-758), 8-bit 4:2:0 frame
 * pictures, every macroblock of one slice type P / B / I (not SP / SI), DFDisableIdc 0 or 1 per macroblock.  The reconstructed
 * planes are filtered IN PLACE on the device, so that a band's reconstruction can stay on the GPU for the halo exchange and the
 * plane build of the next picture.  What the filter reads from encoder state arrives as two arrays:
 *   mbs  [mbh * mbw]   : intra, qp, chroma qp (qpc[0], qpc[1]), luma_transform_size_8x8_flag, DFDisableIdc == 1,
 *                        DFAlphaC0Offset, DFBetaOffset, cbp_blk bits 0..15 (bit 4 * by + bx: the 4x4 luma block holds coefficients)
 *   blks [H / 4][W / 4]: both lists' vectors (quarter-pel) and reference PICTURE identity per 4x4 block (any id that is equal
 *                        exactly when the pictures are the same; -1: the list is unused) -- enc_picture->mv_info.
 * Macroblock order is the reference's (raster; the kernel runs the 2:1 wavefront of JM_PARALLEL_DEBLOCK, loopFilter.c:92-109,
 * one CTA per macroblock row, one warp for luma and one for chroma; boundary strengths by a picture-wide pre-pass): the result is
 * bit-identical to the serial filter. */
typedef struct b2dbk_mb {
  uint8_t intra, qp, qpc_u, qpc_v, transform8x8, disable;
  int8_t alpha_off, beta_off;
  uint16_t cbp_blk, pad_;
} b2dbk_mb;                     /* 12 bytes */
typedef struct b2dbk_blk { int16_t mv[2][2]; int16_t ref[2]; } b2dbk_blk;   /* 12 bytes; mv[list][x, y] */
int b2dbk_frame(int device, int W, int H, uint8_t *y, uint8_t *u, uint8_t *v, const b2dbk_mb *mbs, const b2dbk_blk *blks);   /* host planes, packed */
int b2dbk_frame_dev(int W, int H, uint8_t *y_dev, int y_pitch, uint8_t *u_dev, uint8_t *v_dev, int c_pitch,
                    const b2dbk_mb *mbs_dev, const b2dbk_blk *blks_dev, int *progress_dev /* unused (kept for the ABI; may be NULL): the rows' counters live in stream-ordered scratch */, void *stream);
const char *b2dbk_last_error(void);

/* ---- reference selection per (mode, block) (the first step of the mode decision) ------------------- */
/* list_prediction_cost for list 0 (JM/lencod/src/mode_decision.c:275-300, update_mcost :256-267, ref_cost mv_search.h:114,
 * refbits mv_search.c:377-385) from the search's cost array, for every macroblock of the picture:
 *   cost [nmb][nrefs][41] int64 (cost_sub or cost_int of b2me_search_frame_dev), ref_lambda = lambda_mf[Q_PEL] (RDO build)
 *   best_ref [nmb][21] int8, best_cost [nmb][21] int64 for the entries
 *     0: mode 1;  1..2: mode 2 blocks 0,1;  3..4: mode 3 blocks 0,1;  5 + 4*(m-4) + b: mode m = 4..7, 8x8 quadrant b
 *   (the motion cost of a quadrant in modes 5..7 is the sum over its sub-partitions, as PartitionMotionSearch accumulates it).
 * Device pointers; keeps the 41 x nrefs candidates on the device and hands 21 (reference, cost) pairs per macroblock on. */
int b2me_select_refs_dev(b2me_ctx *ctx, const int64_t *cost_dev, int ref_lambda, int8_t *best_ref_dev, int64_t *best_cost_dev, void *stream);
/* The same for either list of a B slice (list < BI_PRED, mode_decision.c:286-300): cost holds the searches of ONE list
 * ([nmb][nrefs][41], nrefs = the context's), list_size = listXsize[cur_list] <= nrefs bounds the loop and decides whether
 * reference bits are charged (ref_cost: none when the list holds a single picture). */
int b2me_select_refs_list_dev(b2me_ctx *ctx, const int64_t *cost_dev, int list_size, int ref_lambda, int8_t *best_ref_dev, int64_t *best_cost_dev, void *stream);

/* ---- motion-compensated prediction (keeps the vectors on the device between search and transform) ---- */
/* luma_prediction with p_dir == 0 (list 0), no weighting (JM/lencod/src/mc_prediction.c:144-236;
 * OneComponentLumaPrediction :117-136) for every macroblock of the picture, from a search result array.
 *   mb_mode [nmb] uint8: 1 16x16, 2 16x8, 3 8x16, 8 P8x8;   b8mode [nmb][4] uint8: 4 8x8, 5 8x4, 6 4x8, 7 4x4
 *   ref8 [nmb][4] int8: reference slot of each 8x8 quadrant;  mv [nmb][nrefs][41][2]: mv_sub (or mv_int) of the search
 *   orig_blk, pred_blk [nmb*16][16] uint8: the sixteen 4x4 blocks of every MB in raster order, each raster --
 *   exactly the `orig` / `pred` arguments of b2tq_4x4_dev, so search -> prediction -> transform/quant chain on the
 *   device.  All pointers are device pointers. */
int b2me_mc_luma_dev(b2me_ctx *ctx, const uint8_t *mb_mode, const uint8_t *b8mode, const int8_t *ref8, const int16_t *mv,
                     uint8_t *orig_blk, uint8_t *pred_blk, void *stream);

/* Luma AND chroma prediction of 4:2:0 macroblocks from list 0, list 1 or both (SURVEY 8f-1, B slices and chroma):
 *   luma_prediction (JM/lencod/src/mc_prediction.c:144-236) with p_dir 0 / 1 / 2 -- bi_prediction (:82-99) = (l0 + l1 + 1) >> 1 --
 *   and chroma_prediction (:469-566) with the bilinear eighth-sample interpolation of
 *   OneComponentChromaPrediction4x4_regenerate (:292-353).  Unweighted prediction.
 *   pdir [nmb][4] uint8: per 8x8 quadrant 0 list 0, 1 list 1, 2 bi-predictive;  ref8 [nmb][2][4] int8: reference SLOT per list and
 *   quadrant (both lists' pictures live in the context's slots);  mv_l0 / mv_l1 [nmb][nrefs][41][2]: the vectors of either list,
 *   indexed by slot (the two may be the same array);  chroma planes: b2me_set_cur_chroma / b2me_set_ref_chroma.
 *   orig_y / pred_y [nmb*16][16] as b2me_mc_luma_dev;  orig_c / pred_c [nmb][2][4][16]: the four 4x4 blocks (raster) of Cb, then Cr. */
int b2me_set_cur_chroma(b2me_ctx *ctx, const uint8_t *u, const uint8_t *v, int stride);
int b2me_set_cur_chroma_dev(b2me_ctx *ctx, const uint8_t *u_dev, const uint8_t *v_dev, int stride, void *stream);
int b2me_set_ref_chroma(b2me_ctx *ctx, int ref, const uint8_t *u, const uint8_t *v, int stride);
int b2me_set_ref_chroma_dev(b2me_ctx *ctx, int ref, const uint8_t *u_dev, const uint8_t *v_dev, int stride, void *stream);
int b2me_mc_mb_dev(b2me_ctx *ctx, const uint8_t *mb_mode, const uint8_t *b8mode, const uint8_t *pdir, const int8_t *ref8,
                   const int16_t *mv_l0, const int16_t *mv_l1, uint8_t *orig_y, uint8_t *pred_y, uint8_t *orig_c, uint8_t *pred_c, void *stream);

/* ---- mode-decision distortions on precomputed difference blocks ------------------------ */
/* distortion4x4/8x8{SAD,SSE,SATD} (JM/lencod/src/me_distortion.c:38-134; p_Vid->distortion4x4/8x8 as bound by
 * select_distortion :136-158): diff [nblk][n*n] int16 raster, n = 4 or 8, kind 0 SAD, 1 SSE, 2 SATD
 * (HadamardSAD4x4 / HadamardSAD8x8); out [nblk] = dist_scale(distortion) = distortion << 5. */
int b2me_distortion_blocks(int device, int kind, int n, int nblk, const int16_t *diff, int64_t *out);
int b2me_distortion_blocks_dev(int kind, int n, int nblk, const int16_t *diff_dev, int64_t *out_dev, void *stream);

/* ---- instrumentation ---------------------------------------------------------------- */
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t b2me_launch_count(b2me_ctx *ctx);
/* Integer-search counters accumulated since the last reset: out[0] = candidates that survived the
 * packed filter and were re-evaluated exactly, out[1] = window passes, out[2] = (MB,ref) items. */
int b2me_search_stats(b2me_ctx *ctx, int64_t out[3], int reset);
/* Device time (ms, CUDA events on the launching stream) accumulated per kernel family since the
 * last reset: which = 0 integer search, 1 sub-pel planes, 2 sub-pel refinement. Enables timing
 * when enable != 0 (adds two event records per launch). */
int b2me_kernel_timing(b2me_ctx *ctx, int enable);
int b2me_kernel_time_ms(b2me_ctx *ctx, int which, double *ms, int64_t *launches);
/* Micro-benchmark of an instruction mix on the whole chip; kind: 0 VABSDIFF4.ACC only,
 * 1 +IMAD 1:1, 2 +IADD3 1:1, 3 +LOP3 1:1, 4 IADD3 only, 5 IMAD only, 6 VIMNMX.U16x2 only,
 * 7 VABSDIFF4+LDS.  Returns giga warp-lane-ops per second of the FIRST op of the mix. */
int b2me_ubench(int device, int kind, int iters, double *gops);

/* ==== fractal range/domain block search (version1) ====================================== */
/* One context per picture size.  Planes are 8-bit 4:2:0 (chroma planes width/2 x height/2, tightly
 * packed); `con` follows version1: 1 = Y, 2 = U, 3 = V.  Plane sets: 0 = C (previous reconstructed
 * frame), 1..3 = H, M, N (the "fractional-pel" sets the shipped program allocates but never
 * fills, SURVEY Q-F3: they stay zero with zero sum tables unless the caller loads them).
 * Results are indexed [macroblock of the component plane][p], p = the 41-partition numbering
 * above (the 1+2+2+4+8+8+16 range blocks of a 16x16 macroblock). */
typedef struct b2fr_ctx b2fr_ctx;
int  b2fr_create(b2fr_ctx **out, int device, int width, int height, int search_range);
void b2fr_destroy(b2fr_ctx *ctx);
const char *b2fr_last_error(b2fr_ctx *ctx);
/* range (current) frame + its block-grid sums:  <- compute_range_Sum  V1/src/compute.c:686 */
int b2fr_set_range(b2fr_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v);
/* domain (reference) plane set; build_sums != 0 rebuilds its sliding sum tables
 *                                              <- compute_domain_Sum V1/src/compute.c:277 */
int b2fr_set_domain(b2fr_ctx *ctx, int plane_set, const uint8_t *y, const uint8_t *u, const uint8_t *v, int build_sums);
/* full_search of EVERY range block of component `con` against plane set `plane_set`:
 *   xy [nmb][41][2] int32 displacement (0,0 when the start candidate wins),
 *   scale_offset [nmb][41][2] double (alpha, beta after QUAN_A), rms [nmb][41] double (1e30 = rejected)
 *                                              <- full_search V1/src/block_enc.c:1933 over the grid */
int b2fr_search_plane(b2fr_ctx *ctx, int plane_set, int con, int32_t *xy, double *scale_offset, double *rms);
/* Drop-in for ONE call of  double full_search(block_x, block_y, block_size_x, block_size_y, con, TRANS_NODE*)
 * against the plane set changeReferenceFrame selected: xy is in/out (left untouched when the
 * (0,0) start candidate wins, SURVEY Q-F11), scale_offset = {trans->scale, trans->offset}.  The
 * first call after new pictures searches the whole (plane set, component) on the GPU; later calls
 * are table look-ups (SURVEY Q-F10). */
int b2fr_full_search(b2fr_ctx *ctx, int plane_set, int block_x, int block_y, int block_size_x, int block_size_y,
                     int con, int32_t xy[2], double scale_offset[2], double *rms);
/* F5 -- the partition cascade of EVERY macroblock of component `con`:
 *   <- encode_one_macroblock V1/src/block_enc.c:508-1051, encode_block_rect :1072, encode_block_8 :1337, encode_block_4 :1676
 * (num_regions == 1, full search, current view 'C').  Searches the four plane sets if that has not happened yet, then
 * decides on the device: 16x16 is kept unless chun (squared normalised cross-correlation with the co-located block of
 * set C) lies in [0.9, 1] and rms > tol_16^2 * 256; then four 8x8 blocks, each kept if rms <= tol_8^2 * 64, else 8x4
 * pair / 4x8 pair (each block <= tol_8^2 * 32), else four 4x4.  nodes [nmb][21]: the reference's TRANS_NODE tree in
 * pre-order (macroblock, 8x8 block 0, its four children, 8x8 block 1, ...), field for field what the reference leaves
 * there (including what earlier attempts on the same node left behind).  tol = {tol_16, tol_8, tol_4} (tol_4 is read
 * by the reference but decides nothing). */
typedef struct b2fr_node {
  int32_t block_type, partition, reference, x, y, reserved;   /* partition 0 whole, 1 upper/lower, 2 left/right, 3 quarters;
                                                                 reference 0..3 = plane set C, H, M, N */
  double scale, offset;
} b2fr_node;                  /* 40 bytes */
int b2fr_encode_plane(b2fr_ctx *ctx, int con, const double tol[3], b2fr_node *nodes);
/* F8 -- the fractal prediction of EVERY macroblock of component `con` from its TRANS_NODE tree:
 *   <- decode_one_macroblock V1/src/block_dec.c:20-283, decode_block_rect :285, decode_block_8 :760, decode_block_4 :978
 * Each leaf block becomes (unsigned char) bound(0.5 + scale * d + offset - scale * mean_d), d = the block of plane set
 * `reference` displaced by (x, y), mean_d = its sum / n.  nodes [nmb][21] as b2fr_encode_plane returns them, or NULL: use the
 * trees the last b2fr_encode_plane of this component left on the device (search -> cascade -> prediction without a round
 * trip).  rec: the component plane, tightly packed (rows / columns beyond the last whole macroblock are 0). */
int b2fr_decode_plane(b2fr_ctx *ctx, int con, const b2fr_node *nodes, uint8_t *rec);
/* parity read-back of the sum tables: domain table of block size bw x bh at every pixel offset
 * ([h][w] int32, 0 where the block does not fit), range 4x4 table on the block grid ([h/4][w/4]) */
int b2fr_get_domain_table(b2fr_ctx *ctx, int plane_set, int con, int bw, int bh, int squares, int32_t *out);
int b2fr_get_range_table(b2fr_ctx *ctx, int con, int squares, int32_t *out);
int64_t b2fr_launch_count(b2fr_ctx *ctx);

/* ==== fractal range x domain-POOL matching (BASELINE config 5; tensor cores) ==============================
 * Every 8x8 range block of the range plane (raster order, nr = (rw/8)*(rh/8)), in its 8 isometries, against a
 * pool of `pool_size` domain blocks (2x2-averaged 16x16 blocks of the domain plane on a uniform grid,
 * b2fp_pool_positions).  The cross terms are tcgen05.mma kind::i8 tiles; the per-pair fit is compute_rms's
 * (V1/src/compute.c:156-182, QUAN_A V1/inc/defines_enc.h:591-601) in exact integer arithmetic.  version1 itself
 * has no pool search (its full_search, V1/src/block_enc.c:1933, is b2fr_* above): the semantics are DEFINED by
 * oracle/b2_oracle_pool.c, whose header states them.
 *   best_dom [nr] int32  pool index of the best domain (-1: every pair rejected by the alpha limits)
 *   best_iso [nr] uint8  isometry 0..7 of the range block (0 id, 1 mirror x, 2 mirror y, 3 rot 180, 4 transpose,
 *                        5 rot 90 cw, 6 rot 90 ccw, 7 anti-transpose); ties: lowest isometry, then lowest index
 *   aq       [nr] int16  100 * scale after QUAN_A;   beta [nr] int16  offset after QUAN_A
 *   err_num  [nr] int64  640000 * collage error (rms of compute_rms) as an exact integer */
typedef struct b2fp_ctx b2fp_ctx;
int  b2fp_create(b2fp_ctx **out, int device, int range_w, int range_h, int domain_w, int domain_h, int pool_size);
void b2fp_destroy(b2fp_ctx *ctx);
const char *b2fp_last_error(b2fp_ctx *ctx);
int b2fp_pool_positions(b2fp_ctx *ctx, int32_t *xy /* [pool_size][2] top-left corners */);
/* uploads both planes and builds the operands: range rows x 8 isometries and the det-sorted pool, both in
 * UMMA core-matrix order */
int b2fp_set_planes(b2fp_ctx *ctx, const uint8_t *range_plane, int range_stride, const uint8_t *domain_plane, int domain_stride);
int b2fp_set_planes_dev(b2fp_ctx *ctx, const uint8_t *range_dev, int range_stride, const uint8_t *domain_dev, int domain_stride, void *stream);
int b2fp_search(b2fp_ctx *ctx, int32_t *best_dom, uint8_t *best_iso, int16_t *aq, int16_t *beta, int64_t *err_num);
int b2fp_search_dev(b2fp_ctx *ctx, int32_t *best_dom_dev, uint8_t *best_iso_dev, int16_t *aq_dev, int16_t *beta_dev,
                    int64_t *err_num_dev, void *stream);
/* instrumentation: device time of k_frac_pool launched by b2fp_search (CUDA events); a tensor-only pass of the
 * same kernel (no per-pair epilogue arithmetic) = the kind::i8 rate this shape can reach; filter counters
 * out[0] exact evaluations, out[1] 32-column chunks re-examined, out[2] chunks */
int b2fp_kernel_time_ms(b2fp_ctx *ctx, double *ms, int64_t *launches, int reset);
int b2fp_probe(b2fp_ctx *ctx, double *ms);
int b2fp_stats(b2fp_ctx *ctx, int64_t out[3], int reset);
int64_t b2fp_launch_count(b2fp_ctx *ctx);
/* Micro-benchmark: dense tcgen05.mma kind::i8 (u8 x u8 -> s32, M 128 x N 256 x K 32 per instruction) issued back to back on every
 * SM from resident shared-memory operands, no epilogue: the measured int8 rate of the chip in Tops (2 ops per MAC) -- the
 * roofline denominator bench.py uses for k_frac_pool (BASELINE.md section 4 asks for a measured figure). */
int b2fp_ubench_i8(int device, int iters, double *tops);

/* ==== residual transform + quantisation + reconstruction ================================= */
/* Parameter block = the reference's per-(plane, intra, qp) LevelQuantParams table
 * (JM/lencod/inc/global.h LevelQuantParams; q_matrix.c:566-590, q_offsets.c:163-188) plus the
 * switches residual_transform_quant_luma_4x4/_8x8 read from the slice / macroblock. */
typedef struct b2tq_params {
  int32_t qp;            /* qp_scaled of the plane (0..51); qp_per = qp / 6 */
  int32_t mode;          /* 0: JM 18.5 (skip all-zero residual blocks, inverse only when a level is nonzero)
                            1: version1 dct_luma (V1/src/block.c:836; 4x4 only, no level clip) */
  int32_t cavlc;         /* symbol_mode == CAVLC: clip |level| to 2063 (quant4x4_normal.c:86; 4x4, mode 0) */
  int32_t field_scan;    /* 0 zig-zag (SNGL_SCAN), 1 FIELD_SCAN */
  int32_t disthres;      /* COEFF_COST4x4/8x8[disthres] */
  int32_t reserved[3];
  int32_t scale[64];     /* ScaleComp    raster [j*n+i], n = 4 or 8 */
  int32_t offset[64];    /* OffsetComp */
  int32_t invscale[64];  /* InvScaleComp (= dequant << 4 in mode 0; dequant_coef in mode 1) */
} b2tq_params;
/* flat matrices + default rounding offsets of the reference; intra: 0 inter block, 1 intra block
 * of a P/B slice (offset 342/2048 like inter), 2 intra block of an I slice (682/2048); mode as above */
int b2tq_default_params(b2tq_params *p, int is8x8, int qp, int intra, int mode);
/* nblk independent blocks, tightly packed raster blocks: orig/pred/recon [nblk][16|64] u8;
 * level [nblk][16|64] int16 and run [nblk][16|64] u8 = ACLevel/ACRun lists (zero-terminated,
 * zero-padded); coeff_cost [nblk] int32 (the increment the reference adds to *coeff_cost);
 * nonzero [nblk] u8 (the return value).  residual = orig - pred is formed on the device.
 *   <- residual_transform_quant_luma_4x4  JM/lencod/src/block.c:660-724
 *   <- residual_transform_quant_luma_8x8  JM/lencod/src/transform8x8.c:522-602 */
int b2tq_4x4(int device, const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred,
             int16_t *level, uint8_t *run, uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero);
int b2tq_8x8(int device, const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred,
             int16_t *level, uint8_t *run, uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero);
int b2tq_4x4_dev(const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred,
                 int16_t *level, uint8_t *run, uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero, void *stream);
int b2tq_8x8_dev(const b2tq_params *p, int nblk, const uint8_t *orig, const uint8_t *pred,
                 int16_t *level, uint8_t *run, uint8_t *recon, int32_t *coeff_cost, uint8_t *nonzero, void *stream);
/* Intra16x16 luma macroblocks:  <- residual_transform_quant_luma_16x16  JM/lencod/src/block.c:207-345
 *    hadamard4x4 / ihadamard4x4 of the sixteen DC coefficients   JM/lcommon/src/transform.c:121-214
 *    quant_dc4x4_normal / quant_ac4x4_normal                     JM/lencod/src/quant4x4_normal.c:200-270, 117-190
 * p = the INTRA 4x4 table of the plane (b2tq_default_params intra 1 or 2), mode 0.  orig / pred / recon [nmb][256] raster 16x16;
 * dc_level [nmb][16] int16 + dc_run [nmb][16] u8 = cofDC (zero-terminated, zero-padded); ac_level / ac_run [nmb][16][16]: the
 * ACLevel / ACRun lists of the sixteen 4x4 blocks in RASTER block order (the reference files them under cofAC[b8][b4]);
 * ac_coef [nmb] = the return value (15 when any AC level is nonzero). */
int b2tq_16x16(int device, const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
               int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *ac_coef);
int b2tq_16x16_dev(const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                   int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *ac_coef, void *stream);
/* Chroma of 4:2:0 macroblocks, one plane at a time:  <- residual_transform_quant_chroma_4x4  JM/lencod/src/block.c:953-1200
 *    hadamard2x2 / ihadamard2x2 of the four DC coefficients       JM/lcommon/src/transform.c:302-331
 *    quant_dc2x2_normal                                            JM/lencod/src/quantChroma_normal.c:37-96
 *    quant_ac4x4_normal + the _CHROMA_COEFF_COST_ rule             JM/lencod/src/quant4x4_normal.c:117-190, block.c:1137-1168
 * p->qp = the CHROMA qp (qpc[uv]) and p's tables the chroma plane's quantiser (4x4).  (hadamard4x2 belongs to 4:2:2: out of scope.)
 * orig / pred / recon [nmb][64] raster 8x8; dc_level [nmb][4] int16 + dc_run [nmb][4] u8 = cofDC (zero-terminated, zero-padded);
 * ac_level / ac_run [nmb][4][16]: the four 4x4 blocks in raster order; cr_cbp [nmb]: 0 no coefficient, 1 DC only, 2 AC. */
int b2tq_chroma(int device, const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *cr_cbp);
int b2tq_chroma_dev(const b2tq_params *p, int nmb, const uint8_t *orig, const uint8_t *pred, int16_t *dc_level, uint8_t *dc_run,
                    int16_t *ac_level, uint8_t *ac_run, uint8_t *recon, uint8_t *cr_cbp, void *stream);
const char *b2tq_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B2ME_H */
