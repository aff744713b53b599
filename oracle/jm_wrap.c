/* jm_wrap.c -- TEST INFRASTRUCTURE ONLY.  Boundary logger for the stock JM 18.5 encoder:
 * linked with  -Wl,--wrap=full_search_motion_estimation,--wrap=sub_pel_motion_estimation
 * it records the inputs and outputs of every call of the two hot-path entry points
 * (JM/lencod/src/me_fullsearch.c:39 and :186) plus the luma planes they read, into the binary
 * file named by $B2_WRAP_LOG.  The encoder's behaviour is unchanged (the real functions run).
 * Record layout (little endian int32 words):
 *   'F' records: tag 0x46, poc, ref, W, H then W*H bytes cur luma, W*H bytes ref luma ([0][0] plane)
 *                -- emitted the first time a (frame, ref) pair is seen
 *   'I' records: tag 0x49, poc, pos_x, pos_y, blocktype, ref, pred_x, pred_y, cen_x, cen_y,
 *                sr_pel, lambda, min_in(lo,hi), out_x, out_y, cost(lo,hi)
 *   'S' records: tag 0x53, poc, pos_x, pos_y, blocktype, ref, pred_x, pred_y, in_x, in_y,
 *                lamH, lamQ, min_in(lo,hi), out_x, out_y, cost(lo,hi)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mbuffer.h"
#include "me_fullsearch.h"

static FILE *g_log;
static int g_seen_poc[64][16];
static int g_nseen;

static FILE *logf_(void)
{
  if (!g_log) { const char *n = getenv("B2_WRAP_LOG"); g_log = fopen(n ? n : "wrap.log", "wb"); }
  return g_log;
}
static void w32(int v) { fwrite(&v, 4, 1, logf_()); }
static void w64(long long v) { fwrite(&v, 8, 1, logf_()); }

static void dump_frame(Macroblock *currMB, MEBlock *b)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  Slice *s = currMB->p_Slice;
  int poc = p_Vid->enc_picture->poc, ref = b->ref_idx, i, x, y;
  StorablePicture *rp = s->listX[b->list + currMB->list_offset][ref];
  int W = rp->size_x, H = rp->size_y;
  for (i = 0; i < g_nseen; i++) if (g_seen_poc[i][0] == poc && g_seen_poc[i][1] == ref) return;
  if (g_nseen < 64) { g_seen_poc[g_nseen][0] = poc; g_seen_poc[g_nseen][1] = ref; g_nseen++; }
  w32(0x46); w32(poc); w32(ref); w32(W); w32(H);
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) { unsigned char c = (unsigned char)p_Vid->pCurImg[y][x]; fwrite(&c, 1, 1, logf_()); }
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) { unsigned char c = (unsigned char)rp->p_curr_img_sub[0][0][y][x]; fwrite(&c, 1, 1, logf_()); }
}

distblk __real_full_search_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int);
distblk __wrap_full_search_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *b, distblk min_mcost, int lambda)
{
  MotionVector cen = b->mv[(short)b->list];
  distblk c;
  dump_frame(currMB, b);
  c = __real_full_search_motion_estimation(currMB, pred_mv, b, min_mcost, lambda);
  w32(0x49); w32(currMB->p_Vid->enc_picture->poc); w32(b->pos_x); w32(b->pos_y); w32(b->blocktype); w32(b->ref_idx);
  w32(pred_mv->mv_x); w32(pred_mv->mv_y); w32(cen.mv_x); w32(cen.mv_y);
  w32(imin(b->searchRange.max_x, b->searchRange.max_y) >> 2); w32(lambda); w64((long long)min_mcost);
  w32(b->mv[(short)b->list].mv_x); w32(b->mv[(short)b->list].mv_y); w64((long long)c);
  return c;
}

distblk __real_sub_pel_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int *);
distblk __wrap_sub_pel_motion_estimation(Macroblock *currMB, MotionVector *pred, MEBlock *b, distblk min_mcost, int *lambda)
{
  MotionVector in = b->mv[(short)b->list];
  distblk c = __real_sub_pel_motion_estimation(currMB, pred, b, min_mcost, lambda);
  w32(0x53); w32(currMB->p_Vid->enc_picture->poc); w32(b->pos_x); w32(b->pos_y); w32(b->blocktype); w32(b->ref_idx);
  w32(pred->mv_x); w32(pred->mv_y); w32(in.mv_x); w32(in.mv_y);
  w32(lambda[H_PEL]); w32(lambda[Q_PEL]); w64((long long)min_mcost);
  w32(b->mv[(short)b->list].mv_x); w32(b->mv[(short)b->list].mv_y); w64((long long)c);
  fflush(logf_());
  return c;
}
