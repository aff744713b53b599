/* jm_wrap_dbk.c -- TEST INFRASTRUCTURE ONLY.  Boundary logger for the stock JM 18.5 encoder's deblocking filter:
 * linked with  -Wl,--wrap=DeblockFrame  it records, for every coded picture, what DeblockFrame (JM/lencod/src/loopFilter.c:63)
 * reads -- the unfiltered reconstruction, the per-macroblock and per-4x4-block records of include/b2me.h (b2dbk_mb, b2dbk_blk)
 * filled from mb_data[] / enc_picture->mv_info -- and the planes the unmodified function leaves, into $B2_WRAP_LOG.
 * Record: int32 0x44, W, H, slice_type; Y (W*H), U, V (W*H/4 each) before; mbs; blks; Y, U, V after. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mbuffer.h"
#include "b2me.h"

static FILE *g_log;
static FILE *logf_(void)
{
  if (!g_log) { const char *n = getenv("B2_WRAP_LOG"); g_log = fopen(n ? n : "wrap_dbk.log", "wb"); }
  return g_log;
}
static void w32(int v) { fwrite(&v, 4, 1, logf_()); }
static void planes(VideoParameters *p_Vid, imgpel **imgY, imgpel ***imgUV)
{
  int x, y, c, W = p_Vid->width, H = p_Vid->height;
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) { unsigned char b = (unsigned char)imgY[y][x]; fwrite(&b, 1, 1, logf_()); }
  for (c = 0; c < 2; c++) for (y = 0; y < H / 2; y++) for (x = 0; x < W / 2; x++) { unsigned char b = (unsigned char)imgUV[c][y][x]; fwrite(&b, 1, 1, logf_()); }
}

void __real_DeblockFrame(VideoParameters *p_Vid, imgpel **imgY, imgpel ***imgUV);
void __wrap_DeblockFrame(VideoParameters *p_Vid, imgpel **imgY, imgpel ***imgUV)
{
  const int W = p_Vid->width, H = p_Vid->height;
  StorablePicture *ids[64]; int nid = 0;
  unsigned i; int bx, by, l, k;
  if (!imgUV || p_Vid->mb_aff_frame_flag || p_Vid->structure != FRAME || p_Vid->yuv_format != YUV420) { __real_DeblockFrame(p_Vid, imgY, imgUV); return; }
  w32(0x44); w32(W); w32(H); w32(p_Vid->type);
  planes(p_Vid, imgY, imgUV);
  for (i = 0; i < p_Vid->PicSizeInMbs; i++) {
    Macroblock *m = &p_Vid->mb_data[i];
    b2dbk_mb r;
    memset(&r, 0, sizeof(r));
    r.intra = (m->mb_type == I4MB || m->mb_type == I8MB || m->mb_type == I16MB || m->mb_type == IPCM) ? 1 : 0;
    r.qp = (uint8_t)m->qp; r.qpc_u = (uint8_t)m->qpc[0]; r.qpc_v = (uint8_t)m->qpc[1];
    r.transform8x8 = m->luma_transform_size_8x8_flag; r.disable = m->DFDisableIdc == 1;
    r.alpha_off = m->DFAlphaC0Offset; r.beta_off = m->DFBetaOffset; r.cbp_blk = (uint16_t)(m->cbp_blk & 0xffff);
    fwrite(&r, sizeof(r), 1, logf_());
  }
  for (by = 0; by < H / 4; by++)
    for (bx = 0; bx < W / 4; bx++) {
      PicMotionParams *p = &p_Vid->enc_picture->mv_info[by][bx];
      b2dbk_blk r;
      for (l = 0; l < 2; l++) {
        r.mv[l][0] = p->mv[l].mv_x; r.mv[l][1] = p->mv[l].mv_y; r.ref[l] = -1;
        if (p->ref_idx[l] != -1) {                         /* picture identity, as GetStrengthVer compares ref_pic pointers */
          for (k = 0; k < nid && ids[k] != p->ref_pic[l]; k++) ;
          if (k == nid && nid < 64) ids[nid++] = p->ref_pic[l];
          r.ref[l] = (int16_t)k;
        }
      }
      fwrite(&r, sizeof(r), 1, logf_());
    }
  __real_DeblockFrame(p_Vid, imgY, imgUV);
  planes(p_Vid, imgY, imgUV);
}
