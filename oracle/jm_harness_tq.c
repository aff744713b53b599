/* jm_harness_tq.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED JM 18.5 residual transform + quantisation path:
 *   residual_transform_quant_luma_4x4   JM/lencod/src/block.c:660-724
 *   residual_transform_quant_luma_8x8   JM/lencod/src/transform8x8.c:522-602
 *     -> forward4x4 / inverse4x4 / forward8x8 / inverse8x8   JM/lcommon/src/transform.c
 *     -> quant_4x4_normal / quant_8x8_normal                 JM/lencod/src/quant4x4_normal.c:39, quant8x8_normal.c:43
 *     -> sample_reconstruct                                  JM/lcommon/src/blk_prediction.c:48
 * with the quantiser tables built by the reference's own init_qmatrix / CalculateQuant4x4Param /
 * CalculateQuant8x8Param (q_matrix.c:539,591,729) and init_qoffset / CalculateOffset4x4Param /
 * CalculateOffset8x8Param (q_offsets.c:385,487,570).
 */
#include <string.h>
#include <stdlib.h>
#include "global.h"
#include "memalloc.h"
#include "mbuffer.h"
#include "block.h"
#include "transform8x8.h"
#include "q_matrix.h"
#include "q_offsets.h"
#include "quant4x4.h"
#include "quant8x8.h"

typedef struct {
  VideoParameters *p_Vid;
  InputParameters *p_Inp;
  Slice *slice;
  Macroblock mb;
  StorablePicture *enc;
} JMQ;

/* slice_type: P_SLICE (0) / I_SLICE (2); symbol_mode: 0 CAVLC, 1 CABAC */
void *jmq_create(int slice_type, int symbol_mode)
{
  JMQ *h = (JMQ *)calloc(1, sizeof(JMQ));
  VideoParameters *p_Vid = (VideoParameters *)calloc(1, sizeof(VideoParameters));
  InputParameters *p_Inp = (InputParameters *)calloc(1, sizeof(InputParameters));
  Slice *s = (Slice *)calloc(1, sizeof(Slice));
  h->p_Vid = p_Vid; h->p_Inp = p_Inp; h->slice = s;
  p_Vid->p_Inp = p_Inp;
  p_Inp->output.bit_depth[0] = p_Inp->output.bit_depth[1] = 8;
  p_Vid->bitdepth_luma = p_Vid->bitdepth_chroma = 8;
  p_Vid->max_imgpel_value = 255;
  p_Vid->yuv_format = YUV420;
  p_Vid->type = slice_type;
  p_Vid->nal_reference_idc = NALU_PRIORITY_HIGHEST;
  p_Vid->p_Quant = (QuantParameters *)calloc(1, sizeof(QuantParameters));
  p_Vid->p_QScale = (ScaleParameters *)calloc(1, sizeof(ScaleParameters));
  p_Vid->active_sps = (seq_parameter_set_rbsp_t *)calloc(1, sizeof(seq_parameter_set_rbsp_t));
  p_Vid->active_pps = (pic_parameter_set_rbsp_t *)calloc(1, sizeof(pic_parameter_set_rbsp_t));
  init_qmatrix(p_Vid, p_Inp);
  init_qoffset(p_Vid);
  CalculateQuant4x4Param(p_Vid);
  CalculateQuant8x8Param(p_Vid);
  CalculateOffset4x4Param(p_Vid);
  CalculateOffset8x8Param(p_Vid);
  h->enc = alloc_storable_picture(p_Vid, FRAME, 16, 16, 8, 8);
  h->enc->p_curr_img = h->enc->imgY;
  p_Vid->enc_picture = h->enc;
  s->p_Vid = p_Vid; s->p_Inp = p_Inp; s->slice_type = slice_type; s->symbol_mode = (char)symbol_mode;
  get_mem3Dpel(&s->mb_pred, 3, 16, 16);
  get_mem3Dint(&s->mb_ores, 3, 16, 16);
  get_mem3Dint(&s->mb_rres, 3, 16, 16);
  get_mem2Dint(&s->tblk16x16, 16, 16);
  get_mem4Dint(&s->cofAC, 12, 4, 2, 65);
  init_quant_4x4(s);
  init_quant_8x8(s);
  h->mb.p_Vid = p_Vid; h->mb.p_Inp = p_Inp; h->mb.p_Slice = s;
  return h;
}

/* nblk 4x4 blocks; block k is placed at (4*(k&3), 4*((k>>2)&3)) of the macroblock.
 * out: level[nblk][17], run[nblk][17] (ACLevel/ACRun as the reference leaves them; entries after
 * the terminating level 0 are set to 0), recon[nblk][16], cost[nblk], nz[nblk] */
void jmq_tq4x4(void *hh, int qp, int intra, int nblk, const unsigned char *orig, const unsigned char *pred,
               int *level, int *run, unsigned char *recon, int *cost, int *nz)
{
  JMQ *h = (JMQ *)hh; Slice *s = h->slice; int k, i, j;
  h->mb.qp_scaled[0] = qp;
  for (k = 0; k < nblk; k++) {
    int bx = 4 * (k & 3), by = 4 * ((k >> 2) & 3), c = 0, n;
    int pos_x = bx >> 2, pos_y = by >> 2, b8 = 2 * (pos_y >> 1) + (pos_x >> 1), b4 = 2 * (pos_y & 1) + (pos_x & 1);
    for (j = 0; j < 4; j++)
      for (i = 0; i < 4; i++) {
        s->mb_pred[0][by + j][bx + i] = pred[k * 16 + j * 4 + i];
        s->mb_ores[0][by + j][bx + i] = (int)orig[k * 16 + j * 4 + i] - (int)pred[k * 16 + j * 4 + i];
      }
    memset(s->cofAC[b8][b4][0], 0, 65 * sizeof(int)); memset(s->cofAC[b8][b4][1], 0, 65 * sizeof(int));
    nz[k] = residual_transform_quant_luma_4x4(&h->mb, PLANE_Y, bx, by, &c, intra);
    cost[k] = c;
    for (n = 0; n < 17; n++) { level[k * 17 + n] = s->cofAC[b8][b4][0][n]; run[k * 17 + n] = s->cofAC[b8][b4][1][n]; }
    for (n = 0; n < 16 && level[k * 17 + n] != 0; n++) ;
    for (; n < 17; n++) { level[k * 17 + n] = 0; run[k * 17 + n] = 0; }
    for (j = 0; j < 4; j++)
      for (i = 0; i < 4; i++) recon[k * 16 + j * 4 + i] = (unsigned char)h->enc->imgY[by + j][bx + i];
  }
}

/* nblk 8x8 blocks; block k is b8 = k&3.  level/run [nblk][65], recon [nblk][64] */
void jmq_tq8x8(void *hh, int qp, int intra, int nblk, const unsigned char *orig, const unsigned char *pred,
               int *level, int *run, unsigned char *recon, int *cost, int *nz)
{
  JMQ *h = (JMQ *)hh; Slice *s = h->slice; int k, i, j;
  h->mb.qp_scaled[0] = qp;
  for (k = 0; k < nblk; k++) {
    int b8 = k & 3, bx = 8 * (b8 & 1), by = 8 * (b8 >> 1), c = 0, n;
    for (j = 0; j < 8; j++)
      for (i = 0; i < 8; i++) {
        s->mb_pred[0][by + j][bx + i] = pred[k * 64 + j * 8 + i];
        s->mb_ores[0][by + j][bx + i] = (int)orig[k * 64 + j * 8 + i] - (int)pred[k * 64 + j * 8 + i];
      }
    memset(s->cofAC[b8][0][0], 0, 65 * sizeof(int)); memset(s->cofAC[b8][0][1], 0, 65 * sizeof(int));
    nz[k] = residual_transform_quant_luma_8x8(&h->mb, PLANE_Y, b8, &c, intra);
    cost[k] = c;
    for (n = 0; n < 65; n++) { level[k * 65 + n] = s->cofAC[b8][0][0][n]; run[k * 65 + n] = s->cofAC[b8][0][1][n]; }
    for (n = 0; n < 64 && level[k * 65 + n] != 0; n++) ;
    for (; n < 65; n++) { level[k * 65 + n] = 0; run[k * 65 + n] = 0; }
    for (j = 0; j < 8; j++)
      for (i = 0; i < 8; i++) recon[k * 64 + j * 8 + i] = (unsigned char)h->enc->imgY[by + j][bx + i];
  }
}

/* the LevelQuantParams the reference derived, for the product's parameter block:
 * out[3][16] (or [3][64]) = ScaleComp, OffsetComp, InvScaleComp, raster [j][i] */
void jmq_params(void *hh, int is8x8, int qp, int intra, int *out)
{
  JMQ *h = (JMQ *)hh; int n = is8x8 ? 8 : 4, i, j;
  LevelQuantParams **q = is8x8 ? h->p_Vid->p_Quant->q_params_8x8[0][intra][qp] : h->p_Vid->p_Quant->q_params_4x4[0][intra][qp];
  for (j = 0; j < n; j++)
    for (i = 0; i < n; i++) {
      out[0 * n * n + j * n + i] = q[j][i].ScaleComp;
      out[1 * n * n + j * n + i] = q[j][i].OffsetComp;
      out[2 * n * n + j * n + i] = q[j][i].InvScaleComp;
    }
}

/* ---- Intra16x16 luma: residual_transform_quant_luma_16x16 (JM/lencod/src/block.c:207-345) -> forward4x4 x 16, hadamard4x4 /
 * ihadamard4x4 (JM/lcommon/src/transform.c:121-214), quant_dc4x4_normal / quant_ac4x4_normal (quant4x4_normal.c:200, 117),
 * inverse4x4, sample_reconstruct.  nmb macroblocks; orig / pred / recon [nmb][256] raster 16x16;
 * dc_level / dc_run [nmb][17]; ac_level / ac_run [nmb][16 blocks, raster][16]; ac_coef [nmb] = the return value. */
void jmq_tq16x16(void *hh, int qp, int nmb, const unsigned char *orig, const unsigned char *pred,
                 int *dc_level, int *dc_run, int *ac_level, int *ac_run, unsigned char *recon, int *ac_coef)
{
  JMQ *h = (JMQ *)hh; Slice *s = h->slice; VideoParameters *p_Vid = h->p_Vid; int k, i, j, n, b;
  static imgpel **cur;
  if (!s->cofDC) { get_mem3Dint(&s->cofDC, 3, 2, 18); get_mem2Dint(&s->tblk4x4, 4, 4); get_mem4Dpel(&s->mpr_16x16, 3, 5, 16, 16); get_mem2Dpel(&cur, 16, 16); }
  p_Vid->pCurImg = cur;
  h->mb.qp_scaled[0] = qp; h->mb.i16mode = 2; h->mb.opix_y = 0; h->mb.pix_x = 0; h->mb.pix_y = 0; h->mb.is_field_mode = 0;
  for (k = 0; k < nmb; k++) {
    for (j = 0; j < 16; j++)
      for (i = 0; i < 16; i++) { cur[j][i] = orig[k * 256 + j * 16 + i]; s->mpr_16x16[0][2][j][i] = pred[k * 256 + j * 16 + i]; }
    for (b = 0; b < 4; b++) for (n = 0; n < 4; n++) { memset(s->cofAC[b][n][0], 0, 65 * sizeof(int)); memset(s->cofAC[b][n][1], 0, 65 * sizeof(int)); }
    memset(s->cofDC[0][0], 0, 18 * sizeof(int)); memset(s->cofDC[0][1], 0, 18 * sizeof(int));
    ac_coef[k] = residual_transform_quant_luma_16x16(&h->mb, PLANE_Y);
    for (n = 0; n < 17; n++) { dc_level[k * 17 + n] = s->cofDC[0][0][n]; dc_run[k * 17 + n] = s->cofDC[0][1][n]; }
    for (n = 0; n < 16 && dc_level[k * 17 + n] != 0; n++) ;
    for (; n < 17; n++) { dc_level[k * 17 + n] = 0; dc_run[k * 17 + n] = 0; }
    for (b = 0; b < 16; b++) {
      int jj = b >> 2, ii = b & 3, b8 = 2 * (jj >> 1) + (ii >> 1), b4 = 2 * (jj & 1) + (ii & 1);
      int *L = ac_level + (k * 16 + b) * 16, *R = ac_run + (k * 16 + b) * 16;
      for (n = 0; n < 16; n++) { L[n] = s->cofAC[b8][b4][0][n]; R[n] = s->cofAC[b8][b4][1][n]; }
      for (n = 0; n < 15 && L[n] != 0; n++) ;
      for (; n < 16; n++) { L[n] = 0; R[n] = 0; }
    }
    for (j = 0; j < 16; j++)
      for (i = 0; i < 16; i++) recon[k * 256 + j * 16 + i] = (unsigned char)h->enc->imgY[j][i];
  }
}

/* ---- chroma (4:2:0): residual_transform_quant_chroma_4x4 (JM/lencod/src/block.c:953-1200) -> forward4x4 x 4, hadamard2x2 /
 * ihadamard2x2 (lcommon/src/transform.c:302-331), quant_dc2x2_normal (quantChroma_normal.c:37-96), quant_ac4x4_normal
 * (quant4x4_normal.c:117-190) with the _CHROMA_COEFF_COST_ rule, inverse4x4, sample_reconstruct.
 * nmb chroma blocks of one plane; orig / pred / recon [nmb][64] raster 8x8; dc_level / dc_run [nmb][5]; ac_level / ac_run
 * [nmb][4 blocks, raster][16]; cr_cbp [nmb] = the return value (called with cr_cbp = 0); plane = uv + 1 selects the quantiser set. */
#include "quantChroma.h"
void jmq_tq_chroma(void *hh, int qpc, int intra, int uv, int nmb, const unsigned char *orig, const unsigned char *pred,
                   int *dc_level, int *dc_run, int *ac_level, int *ac_run, unsigned char *recon, int *cr_cbp)
{
  JMQ *h = (JMQ *)hh; Slice *s = h->slice; VideoParameters *p_Vid = h->p_Vid; int k, i, j, n, b;
  if (!s->cofDC) { get_mem3Dint(&s->cofDC, 3, 2, 18); get_mem2Dint(&s->tblk4x4, 4, 4); }
  if (!h->enc->imgUV) get_mem3Dpel(&h->enc->imgUV, 2, 8, 8);
  p_Vid->mb_cr_size_x = p_Vid->mb_cr_size_y = 8; p_Vid->num_blk8x8_uv = 2; p_Vid->yuv_format = YUV420; p_Vid->AdaptiveRounding = 0;
  s->bitdepth_chroma_qp_scale = 0;
  init_quant_Chroma(s);
  h->mb.qpc[uv] = qpc; h->mb.mb_type = intra ? I16MB : P16x16; h->mb.pix_c_x = 0; h->mb.pix_c_y = 0; h->mb.is_field_mode = 0;
  h->mb.luma_transform_size_8x8_flag = 0;
  p_Vid->max_pel_value_comp[1] = p_Vid->max_pel_value_comp[2] = 255;
  for (k = 0; k < nmb; k++) {
    for (j = 0; j < 8; j++)
      for (i = 0; i < 8; i++) {
        s->mb_pred[uv + 1][j][i] = pred[k * 64 + j * 8 + i];
        s->mb_ores[uv + 1][j][i] = (int)orig[k * 64 + j * 8 + i] - (int)pred[k * 64 + j * 8 + i];
      }
    for (b = 0; b < 4; b++) { memset(s->cofAC[4 + uv][b][0], 0, 65 * sizeof(int)); memset(s->cofAC[4 + uv][b][1], 0, 65 * sizeof(int)); }
    memset(s->cofDC[uv + 1][0], 0, 18 * sizeof(int)); memset(s->cofDC[uv + 1][1], 0, 18 * sizeof(int));
    h->mb.cbp_blk = 0;
    cr_cbp[k] = residual_transform_quant_chroma_4x4(&h->mb, uv, 0);
    for (n = 0; n < 5; n++) { dc_level[k * 5 + n] = s->cofDC[uv + 1][0][n]; dc_run[k * 5 + n] = s->cofDC[uv + 1][1][n]; }
    for (n = 0; n < 4 && dc_level[k * 5 + n] != 0; n++) ;
    for (; n < 5; n++) { dc_level[k * 5 + n] = 0; dc_run[k * 5 + n] = 0; }
    for (b = 0; b < 4; b++) {                      /* cofAC[4 + uv][b4]: b4 = 2 * (y / 4) + (x / 4) in 4:2:0 (hor_offset / ver_offset) */
      int *L = ac_level + (k * 4 + b) * 16, *R = ac_run + (k * 4 + b) * 16;
      for (n = 0; n < 16; n++) { L[n] = s->cofAC[4 + uv][b][0][n]; R[n] = s->cofAC[4 + uv][b][1][n]; }
      for (n = 0; n < 15 && L[n] != 0; n++) ;
      for (; n < 16; n++) { L[n] = 0; R[n] = 0; }
    }
    for (j = 0; j < 8; j++)
      for (i = 0; i < 8; i++) recon[k * 64 + j * 8 + i] = (unsigned char)h->enc->imgUV[uv][j][i];
  }
}
/* LevelQuantParams of a chroma plane (plane = uv + 1): out[3][16] = ScaleComp, OffsetComp, InvScaleComp */
void jmq_params_chroma(void *hh, int plane, int qp, int intra, int *out)
{
  JMQ *h = (JMQ *)hh; int i, j;
  LevelQuantParams **q = h->p_Vid->p_Quant->q_params_4x4[plane][intra][qp];
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++) { out[j * 4 + i] = q[j][i].ScaleComp; out[16 + j * 4 + i] = q[j][i].OffsetComp; out[32 + j * 4 + i] = q[j][i].InvScaleComp; }
}
