/* jm_wrap_epzs.c -- TEST INFRASTRUCTURE ONLY.  Boundary logger for the stock JM 18.5 encoder's EPZS integer search:
 * linked with  -Wl,--wrap=EPZS_motion_estimation,--wrap=EPZS_subMB_motion_estimation  it records, for every call of the two
 * functions (JM/lencod/src/me_epzs.c:54 and :417), the job the drop-in shim would hand to the GPU -- built by the SAME code,
 * integration/jm/b2me_jm_epzs_job.h -- and what the unmodified function returned (cost, vector), plus the luma planes it
 * read, into the binary file named by $B2_WRAP_LOG.  The encoder's behaviour is unchanged: the real functions run.
 * Records (little endian):
 *   'F': int32 0x46, poc, ref, W, H; W*H bytes current luma; W*H bytes reference luma -- first time a (poc, ref) pair is seen
 *   'P': int32 0x50, npat; npat * sizeof(b2me_epzs_pattern) bytes -- whenever the pattern table grew
 *   'E': int32 0x45, poc, ref, submb, npred; sizeof(b2me_epzs_job) bytes; npred*2 int16; int64 cost; int32 mv_x, mv_y
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mbuffer.h"
#include "mv_search.h"
#include "b2me.h"

static void b2_fail(const char *what) { error((char *)what, 500); }
#include "b2me_jm_epzs_job.h"

static FILE *g_log;
static int g_seen[256][2], g_nseen, g_npat_logged;
static FILE *logf_(void)
{
  if (!g_log) { const char *n = getenv("B2_WRAP_LOG"); g_log = fopen(n ? n : "wrap_epzs.log", "wb"); }
  return g_log;
}
static void w32(int v) { fwrite(&v, 4, 1, logf_()); }
static void w64(long long v) { fwrite(&v, 8, 1, logf_()); }

static void dump_frame(Macroblock *currMB, MEBlock *b)
{
  VideoParameters *p_Vid = currMB->p_Vid;
  int poc = p_Vid->enc_picture->poc, ref = b->ref_idx, i, x, y;
  StorablePicture *rp = currMB->p_Slice->listX[b->list + currMB->list_offset][ref];
  int W = rp->size_x, H = rp->size_y;
  for (i = 0; i < g_nseen; i++) if (g_seen[i][0] == poc && g_seen[i][1] == ref) return;
  if (g_nseen < 256) { g_seen[g_nseen][0] = poc; g_seen[g_nseen][1] = ref; g_nseen++; }
  w32(0x46); w32(poc); w32(ref); w32(W); w32(H);
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) { unsigned char c = (unsigned char)p_Vid->pCurImg[y][x]; fwrite(&c, 1, 1, logf_()); }
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) { unsigned char c = (unsigned char)rp->p_curr_img_sub[0][0][y][x]; fwrite(&c, 1, 1, logf_()); }
}

distblk __real_EPZS_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int);
distblk __real_EPZS_subMB_motion_estimation(Macroblock *, MotionVector *, MEBlock *, distblk, int);
static distblk wrap(Macroblock *currMB, MotionVector *pred_mv, MEBlock *b, distblk min_mcost, int lambda, int submb)
{
  b2me_epzs_job J;
  int16_t pv[2 * B2_EPZS_MAXPRED];
  int n;
  distblk c;
  if (b->list != 0 || currMB->list_offset) return submb ? __real_EPZS_subMB_motion_estimation(currMB, pred_mv, b, min_mcost, lambda)
                                                        : __real_EPZS_motion_estimation(currMB, pred_mv, b, min_mcost, lambda);
  dump_frame(currMB, b);
  n = b2_epzs_build_job(currMB, pred_mv, b, lambda, submb, b->ref_idx, &J, pv);
  if (g_npat != g_npat_logged) { w32(0x50); w32(g_npat); fwrite(g_pat, sizeof(b2me_epzs_pattern), g_npat, logf_()); g_npat_logged = g_npat; }
  c = submb ? __real_EPZS_subMB_motion_estimation(currMB, pred_mv, b, min_mcost, lambda) : __real_EPZS_motion_estimation(currMB, pred_mv, b, min_mcost, lambda);
  w32(0x45); w32(currMB->p_Vid->enc_picture->poc); w32(b->ref_idx); w32(submb); w32(n);
  fwrite(&J, sizeof(J), 1, logf_()); fwrite(pv, 4, n, logf_());
  w64((long long)c); w32(b->mv[0].mv_x); w32(b->mv[0].mv_y);
  return c;
}
distblk __wrap_EPZS_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *b, distblk min_mcost, int lambda)
{ return wrap(currMB, pred_mv, b, min_mcost, lambda, 0); }
distblk __wrap_EPZS_subMB_motion_estimation(Macroblock *currMB, MotionVector *pred_mv, MEBlock *b, distblk min_mcost, int lambda)
{ return wrap(currMB, pred_mv, b, min_mcost, lambda, 1); }
