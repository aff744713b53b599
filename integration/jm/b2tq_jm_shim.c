/* b2tq_jm_shim.c -- the reference-side binding of the residual transform + quantisation: JM 18.5's
 *     int residual_transform_quant_luma_4x4(Macroblock *currMB, ColorPlane pl, int block_x, int block_y, int *coeff_cost, int intra)
 * (JM/lencod/src/block.c:660-724: forward4x4 -> quant_4x4_normal -> inverse4x4 -> sample_reconstruct) served by b2tq_4x4 of
 * libb2me.so (k_tq4x4, hand-written sm_100a CUDA behind include/b2me.h).
 *
 * The function lives in block.o next to its callers and is reached through currMB->residual_transform_quant_luma_4x4
 * (set at block.c:2407/2428), so the object's own definition is weakened (objcopy --weaken-symbol, oracle/Makefile.jm:
 * the source is untouched) and this strong definition takes every reference, the function-pointer assignment included.
 * One block per call is the reference's own granularity -- a functional drop-in of the boundary; throughput comes from the
 * batched calls (b2tq_4x4_dev over all blocks of a picture, chained behind b2me_mc_luma_dev).
 * What the caller sees afterwards is what block.c:660-724 leaves: ACLevel / ACRun of cofAC[b8][b4], the reconstructed
 * samples in enc_picture, *coeff_cost, currMB->subblock_x/y, the return value.  (mb_rres, the reconstructed residual, is
 * scratch of the reference function in this configuration and is not produced.)
 * Configurations outside the kernel's coverage stop the encoder through JM's error() -- no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mbuffer.h"
#include "block.h"
#include "quant4x4.h"
#include "quant_params.h"
#include "b2me.h"

static long g_tq_calls;
static int g_tq_dev = -1;

static void b2tq_fail(const char *what)
{
  char msg[400];
  snprintf(msg, sizeof(msg), "b2tq shim: %s (%s)", what, b2tq_last_error());
  error(msg, 500);
}
static void b2tq_report(void)
{
  if (getenv("B2ME_SHIM_VERBOSE")) fprintf(stderr, "b2tq shim: %ld transform/quant calls\n", g_tq_calls);
}

int residual_transform_quant_luma_4x4(Macroblock *currMB, ColorPlane pl, int block_x, int block_y, int *coeff_cost, int intra)
{
  const int pos_x = block_x >> BLOCK_SHIFT, pos_y = block_y >> BLOCK_SHIFT;
  const int b8 = 2 * (pos_y >> 1) + (pos_x >> 1) + (pl << 2), b4 = 2 * (pos_y & 0x01) + (pos_x & 0x01);
  Slice *currSlice = currMB->p_Slice;
  VideoParameters *p_Vid = currSlice->p_Vid;
  imgpel **img_enc = p_Vid->enc_picture->p_curr_img;
  imgpel **mb_pred = currSlice->mb_pred[pl];
  int **mb_ores = currSlice->mb_ores[pl];
  const int qp = currMB->qp_scaled[pl];
  LevelQuantParams **q = p_Vid->p_Quant->q_params_4x4[pl][intra][qp];
  int *ACLevel = currSlice->cofAC[b8][b4][0], *ACRun = currSlice->cofAC[b8][b4][1];
  b2tq_params P;
  uint8_t orig[16], pred[16], recon[16], run[16], nz = 0;
  int16_t level[16];
  int32_t cost = 0;
  int i, j, any = 0;

  if (g_tq_dev < 0) { const char *e = getenv("B2ME_DEVICE"); g_tq_dev = e ? atoi(e) : 0; atexit(b2tq_report); }
  if (pl != PLANE_Y || p_Vid->yuv_format == YUV444) b2tq_fail("only the luma plane of 4:2:0 / 4:0:0 is supported");
  if (p_Vid->bitdepth_luma != 8) b2tq_fail("only 8-bit luma is supported");
  if (p_Vid->AdaptiveRounding) b2tq_fail("AdaptiveRounding is not supported");
  if (currSlice->quant_4x4 != quant_4x4_normal) b2tq_fail("only quant_4x4_normal (no RDOQ, no adaptive rounding) is supported");
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++) {
      const int p = mb_pred[block_y + j][block_x + i], r = mb_ores[block_y + j][block_x + i];
      pred[4 * j + i] = (uint8_t)p; orig[4 * j + i] = (uint8_t)(p + r);     /* mb_ores = original - prediction */
      any |= r;
    }
  memset(&P, 0, sizeof(P));
  P.qp = qp; P.mode = 0; P.cavlc = currSlice->symbol_mode == CAVLC; P.field_scan = currMB->is_field_mode ? 1 : 0;
  P.disthres = currSlice->disthres;
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++) {
      P.scale[4 * j + i] = q[j][i].ScaleComp; P.offset[4 * j + i] = q[j][i].OffsetComp; P.invscale[4 * j + i] = q[j][i].InvScaleComp;
    }
  if (b2tq_4x4(g_tq_dev, &P, 1, orig, pred, level, run, recon, &cost, &nz) != B2ME_OK) b2tq_fail("b2tq_4x4 failed");
  g_tq_calls++;
  if (any) {
    currMB->subblock_x = ((b8 & 0x1) == 0) ? (((b4 & 0x1) == 0) ? 0 : 4) : (((b4 & 0x1) == 0) ? 8 : 12);
    currMB->subblock_y = (b8 < 2) ? ((b4 < 2) ? 0 : 4) : ((b4 < 2) ? 8 : 12);
    for (i = 0; i < 16 && level[i] != 0; i++) { ACLevel[i] = level[i]; ACRun[i] = run[i]; }
    ACLevel[i] = 0;
    *coeff_cost += cost;
  } else {
    ACLevel[0] = 0;
  }
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++) img_enc[currMB->pix_y + block_y + j][currMB->pix_x + block_x + i] = recon[4 * j + i];
  return nz;
}
