/* b2me_jm_bid_job.h -- reference-side half of the BIDPartitionCost / BPredPartitionCost boundary (JM/lencod/src/mv_search.c:1159-1250, 589-700): builds the
 * b2me_bid_job record of include/b2me.h from the encoder's own state.  What stays on the host is what needs the encoder's
 * motion field: mvd_bits, added up exactly as mv_bit_cost does (mv_search.c:559-581: get_neighbors + currMB->GetMVPredictor per
 * sub-block and list, p_Vid->mvbits[]).  Included by a drop-in shim (the call becomes one record of a batched
 * b2me_bid_partition_cost) and by the test logger oracle/jm_wrap_bid.c, so both build the record with the same code.
 * slot_l0 / slot_l1: the context's reference slots that hold listX[LIST_0 + list_offset][cur_ref[0]] and
 * listX[LIST_1 + list_offset][cur_ref[1]] (the caller's picture cache decides them). */
#ifndef B2ME_JM_BID_JOB_H
#define B2ME_JM_BID_JOB_H
#include <string.h>
#include "global.h"
#include "mv_search.h"
#include "b2me.h"

static const short b2_bid_bs[8][2] = {{16,16}, {16,16}, {16,8}, {8,16}, {8,8}, {8,4}, {4,8}, {4,4}};   /* block_size[][] (lencod/src/mv_search.c:46) */
static const short b2_bid_bx0[5][4] = {{0,0,0,0}, {0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,2,0,2}};
static const short b2_bid_by0[5][4] = {{0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,0,0,0}, {0,0,2,2}};

/* returns 1 when weighted_bi_prediction applies (luma_prediction's apply_weights for p_dir == 2).
 * mv_array: currSlice->all_mv for BIDPartitionCost; currSlice->bipred_mv[list] for its twin BPredPartitionCost
 * (mv_search.c:589-700: the same cost on the vectors of the bi-predictive motion search, luma_prediction_bi mc_prediction.c:244-284) */
static int b2_bipart_build_job(Macroblock *currMB, MotionVector *****mv_array, int blocktype, int block8x8, char cur_ref[2], int lambda_factor,
                               int slot_l0, int slot_l1, b2me_bid_job *J);
static int b2_bid_build_job(Macroblock *currMB, int blocktype, int block8x8, char cur_ref[2], int lambda_factor,
                            int slot_l0, int slot_l1, b2me_bid_job *J)
{ return b2_bipart_build_job(currMB, currMB->p_Slice->all_mv, blocktype, block8x8, cur_ref, lambda_factor, slot_l0, slot_l1, J); }
static int b2_bpred_build_job(Macroblock *currMB, int blocktype, int block8x8, short ref_l0, short ref_l1, int lambda_factor, int list,
                              int slot_l0, int slot_l1, b2me_bid_job *J)
{ char cur_ref[2]; cur_ref[0] = (char)ref_l0; cur_ref[1] = (char)ref_l1;
  return b2_bipart_build_job(currMB, currMB->p_Slice->bipred_mv[list], blocktype, block8x8, cur_ref, lambda_factor, slot_l0, slot_l1, J); }
static int b2_bipart_build_job(Macroblock *currMB, MotionVector *****mv_array, int blocktype, int block8x8, char cur_ref[2], int lambda_factor,
                               int slot_l0, int slot_l1, b2me_bid_job *J)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  Slice *currSlice = currMB->p_Slice;
  const int parttype = blocktype < 4 ? blocktype : 4;
  const int step_h0 = b2_bid_bs[parttype][0], step_v0 = b2_bid_bs[parttype][1];
  const int step_h = b2_bid_bs[blocktype][0], step_v = b2_bid_bs[blocktype][1];
  const int bx = b2_bid_bx0[parttype][block8x8] << 2, by = b2_bid_by0[parttype][block8x8] << 2;
  int list, v, h, n, bits = 0;
  memset(J, 0, sizeof(*J));
  J->mb_x = (int16_t)currMB->pix_x; J->mb_y = (int16_t)currMB->opix_y;
  J->blocktype = (int16_t)blocktype; J->block8x8 = (int16_t)block8x8;
  J->ref_l0 = (int16_t)slot_l0; J->ref_l1 = (int16_t)slot_l1;
  for (list = 0; list < 2; list++) {
    MotionVector **all_mv = mv_array[list][(int)cur_ref[list]][blocktype];
    n = 0;
    for (v = by; v < by + step_v0; v += step_v)
      for (h = bx; h < bx + step_h0; h += step_h, n++) {
        PixelPos block[4];
        MotionVector pmv;
        get_neighbors(currMB, block, h, v, step_h);
        currMB->GetMVPredictor(currMB, block, &pmv, cur_ref[list], p_Vid->enc_picture->mv_info, list, h, v, step_h, step_v);
        bits += p_Vid->mvbits[all_mv[v >> 2][h >> 2].mv_x - pmv.mv_x] + p_Vid->mvbits[all_mv[v >> 2][h >> 2].mv_y - pmv.mv_y];
        if (list == 0) { J->mv_l0[n][0] = all_mv[v >> 2][h >> 2].mv_x; J->mv_l0[n][1] = all_mv[v >> 2][h >> 2].mv_y; }
        else           { J->mv_l1[n][0] = all_mv[v >> 2][h >> 2].mv_x; J->mv_l1[n][1] = all_mv[v >> 2][h >> 2].mv_y; }
      }
  }
  J->mvd_bits = bits; J->lambda_factor = lambda_factor;
  if (currSlice->weighted_prediction == 1 || currSlice->weighted_prediction == 2) {
    J->weight_l0 = (int16_t)currSlice->wbp_weight[0][(int)cur_ref[0]][(int)cur_ref[1]][0];
    J->weight_l1 = (int16_t)currSlice->wbp_weight[1][(int)cur_ref[0]][(int)cur_ref[1]][0];
    J->offset_bi = (int16_t)((currSlice->wp_offset[0][(int)cur_ref[0]][0] + currSlice->wp_offset[1][(int)cur_ref[1]][0] + 1) >> 1);
    return 1;
  }
  return 0;
}
#endif
