/* jm_wrap_bid.c -- TEST INFRASTRUCTURE ONLY.  Boundary logger for the stock JM 18.5 encoder's BIDPartitionCost
 * (JM/lencod/src/mv_search.c:1159-1250) and its twin BPredPartitionCost (:589-700): linked with  -Wl,--wrap=BIDPartitionCost,--wrap=BPredPartitionCost  it records, for the calls the mode decision
 * makes (mode_decision.c), the b2me_bid_job record a drop-in shim would hand to the GPU -- built by the SAME code,
 * integration/jm/b2me_jm_bid_job.h -- and what the unmodified function returned, plus the luma pictures it read, into the
 * binary file named by $B2_WRAP_LOG.  The encoder's behaviour is unchanged: the real function runs.
 * Records (little endian):
 *   'C': int32 0x43, poc, W, H; W*H bytes: the current picture (pCurImg) -- whenever the coded picture changes; resets the slots
 *   'R': int32 0x52, slot, W, H; W*H bytes: integer luma of a reference picture -- first time the picture is met under this poc
 *   'B': int32 0x42, metric (ModeDecisionMetric), transform8x8, apply_weights | twin << 1 (twin: the call was BPredPartitionCost), luma_log_weight_denom; sizeof(b2me_bid_job) bytes; int64 cost
 * Every $B2_WRAP_STRIDE-th call is kept (default 1). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mbuffer.h"
#include "mv_search.h"
#include "b2me.h"
#include "b2me_jm_bid_job.h"

static FILE *g_log;
static StorablePicture *g_slot[32];
static int g_nslot, g_poc = 0x7fffffff, g_calls, g_stride;
static FILE *logf_(void)
{
  if (!g_log) { const char *n = getenv("B2_WRAP_LOG"), *s = getenv("B2_WRAP_STRIDE"); g_log = fopen(n ? n : "wrap_bid.log", "wb"); g_stride = s ? atoi(s) : 1; if (g_stride < 1) g_stride = 1; }
  return g_log;
}
static void w32(int v) { fwrite(&v, 4, 1, logf_()); }
static void plane(imgpel **img, int W, int H)
{
  int x, y;
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) { unsigned char c = (unsigned char)img[y][x]; fwrite(&c, 1, 1, logf_()); }
}
static int slot_of(VideoParameters *p_Vid, StorablePicture *p)
{
  int k;
  for (k = 0; k < g_nslot; k++) if (g_slot[k] == p) return k;
  if (g_nslot >= 32) error("jm_wrap_bid: more than 32 reference pictures under one coded picture", 500);
  g_slot[g_nslot] = p;
  w32(0x52); w32(g_nslot); w32(p_Vid->width); w32(p_Vid->height);
  plane(p->p_curr_img_sub[0][0], p_Vid->width, p_Vid->height);
  return g_nslot++;
}

static void log_call(Macroblock *currMB, StorablePicture *p0, StorablePicture *p1, const b2me_bid_job *Jin, int wp, int twin, distblk c)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  Slice *currSlice = currMB->p_Slice;
  b2me_bid_job J = *Jin;
  if (p_Vid->enc_picture->poc != g_poc) {
    g_poc = p_Vid->enc_picture->poc; g_nslot = 0;
    w32(0x43); w32(g_poc); w32(p_Vid->width); w32(p_Vid->height);
    plane(p_Vid->pCurImg, p_Vid->width, p_Vid->height);
  }
  J.ref_l0 = (int16_t)slot_of(p_Vid, p0); J.ref_l1 = (int16_t)slot_of(p_Vid, p1);
  w32(0x42); w32(currSlice->p_Inp->ModeDecisionMetric); w32(currSlice->p_Inp->Transform8x8Mode != 0); w32(wp | (twin << 1)); w32(currSlice->luma_log_weight_denom);
  fwrite(&J, sizeof(J), 1, logf_());
  { long long v = (long long)c; fwrite(&v, 8, 1, logf_()); }
}

/* the twin: the same cost on the vectors of the bi-predictive motion search (mode_decision.c:366, 372) */
distblk __real_BPredPartitionCost(Macroblock *currMB, int blocktype, int block8x8, short ref_l0, short ref_l1, int lambda_factor, int list);
distblk __wrap_BPredPartitionCost(Macroblock *currMB, int blocktype, int block8x8, short ref_l0, short ref_l1, int lambda_factor, int list)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  Slice *currSlice = currMB->p_Slice;
  const distblk c = __real_BPredPartitionCost(currMB, blocktype, block8x8, ref_l0, ref_l1, lambda_factor, list);
  b2me_bid_job J;
  int wp;
  logf_();
  if (p_Vid->mb_aff_frame_flag || p_Vid->structure != FRAME || (g_calls++ % g_stride)) return c;
  wp = b2_bpred_build_job(currMB, blocktype, block8x8, ref_l0, ref_l1, lambda_factor, list, 0, 0, &J);
  log_call(currMB, currSlice->listX[LIST_0 + currMB->list_offset][ref_l0], currSlice->listX[LIST_1 + currMB->list_offset][ref_l1], &J, wp, 1, c);
  return c;
}

distblk __real_BIDPartitionCost(Macroblock *currMB, int blocktype, int block8x8, char cur_ref[2], int lambda_factor);
distblk __wrap_BIDPartitionCost(Macroblock *currMB, int blocktype, int block8x8, char cur_ref[2], int lambda_factor)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  Slice *currSlice = currMB->p_Slice;
  const distblk c = __real_BIDPartitionCost(currMB, blocktype, block8x8, cur_ref, lambda_factor);
  b2me_bid_job J;
  int wp;
  logf_();
  if (p_Vid->mb_aff_frame_flag || p_Vid->structure != FRAME || (g_calls++ % g_stride)) return c;
  wp = b2_bid_build_job(currMB, blocktype, block8x8, cur_ref, lambda_factor, 0, 0, &J);
  log_call(currMB, currSlice->listX[LIST_0 + currMB->list_offset][(int)cur_ref[LIST_0]], currSlice->listX[LIST_1 + currMB->list_offset][(int)cur_ref[LIST_1]], &J, wp, 0, c);
  return c;
}
