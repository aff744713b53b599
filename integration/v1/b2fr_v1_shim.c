/* b2fr_v1_shim.c -- the reference-side binding for version1 (2.论文程序/ZhangLing_Yu_version1/H264Fractal): its
 * range/domain block search served by libb2me.so (hand-written sm_100a CUDA behind include/b2me.h).
 *
 * Link-level replacement of ONE function: `double full_search(int block_x, int block_y, int block_size_x,
 * int block_size_y, int con, TRANS_NODE *trans)` (V1/inc/block_enc.h:13), which V1/src/block_enc.c:1933 defines next to
 * its callers.  A maintainer guards that definition with `#ifndef B2FR`; this repository leaves the source untouched and
 * weakens the symbol in the compiled object instead (objcopy --weaken-symbol=full_search, oracle/Makefile.v1), so that the
 * definition below wins at link time and every call inside block_enc.o lands here.
 *
 * Everything else of version1 runs unchanged: encode_one_macroblock and its cascade call full_search exactly as before
 * and read the same globals.  State: one b2fr_ctx for the picture geometry (input->imagewidth/imageheight/search_range);
 * b2fr_v1_new_frame() goes where V1/src/code.c:256-270 calls compute_domain_Sum() / compute_range_Sum(): it uploads the
 * range frame and the domain plane sets.  The plane set a call refers to is the one changeReferenceFrame(char) selected,
 * recognised by the *_temp pointers it swaps (V1/src/block_enc.c:80-214).  The first full_search after new pictures
 * searches every range block of the (plane set, component) on the GPU; later calls are look-ups.
 * No CPU fallback: any failure stops the program.
 */
#include <stdio.h>
#include <stdlib.h>
#include "windows.h"
#include "global.h"
#include "i_global.h"
#include "block_enc.h"
#include "b2me.h"

static b2fr_ctx *g_fr;
static long g_calls;

static void b2fr_fail(const char *what)
{
  fprintf(stderr, "b2fr shim: %s (%s)\n", what, b2fr_last_error(g_fr));
  exit(1);
}

/* have[s] != 0: plane set s (0 C, 1 H, 2 M, 3 N) holds a picture and its sum tables were built (compute_domain_Sum);
 * sets the program never fills stay zero with zero tables on both sides (SURVEY Q-F3) */
void b2fr_v1_new_frame(const int have[4])
{
  byte **y[4]; byte ***uv[4]; int s;
  y[0] = imgY_ref; y[1] = imgY_ref_h; y[2] = imgY_ref_m; y[3] = imgY_ref_n;
  uv[0] = imgUV_ref; uv[1] = imgUV_ref_h; uv[2] = imgUV_ref_m; uv[3] = imgUV_ref_n;
  if (!g_fr) {
    const char *e = getenv("B2ME_DEVICE");
    if (b2fr_create(&g_fr, e ? atoi(e) : 0, input->imagewidth, input->imageheight, input->search_range) != B2ME_OK) b2fr_fail("b2fr_create failed");
  }
  for (s = 0; s < 4; s++)      /* rows are contiguous (get_mem2D, V1/src/memalloc.c) */
    if (have[s] && b2fr_set_domain(g_fr, s, y[s][0], uv[s][0][0], uv[s][1][0], 1) != B2ME_OK) b2fr_fail("b2fr_set_domain failed");
  if (b2fr_set_range(g_fr, imgY_org[0], imgUV_org[0][0], imgUV_org[1][0]) != B2ME_OK) b2fr_fail("b2fr_set_range failed");
}

long b2fr_v1_calls(void) { return g_calls; }

double full_search(int block_x, int block_y, int block_size_x, int block_size_y, int con, TRANS_NODE *trans)
{
  int32_t xy[2]; double so[2], rms;
  const int set = imgY_ref_temp == imgY_ref ? 0 : imgY_ref_temp == imgY_ref_h ? 1 : imgY_ref_temp == imgY_ref_m ? 2 :
                  imgY_ref_temp == imgY_ref_n ? 3 : -1;
  if (!g_fr) { fprintf(stderr, "b2fr shim: full_search before b2fr_v1_new_frame()\n"); exit(1); }
  if (set < 0) b2fr_fail("changeReferenceFrame selected a plane set the shim does not know");
  xy[0] = trans->x; xy[1] = trans->y;
  if (b2fr_full_search(g_fr, set, block_x, block_y, block_size_x, block_size_y, con, xy, so, &rms) != B2ME_OK) b2fr_fail("b2fr_full_search failed");
  trans->x = xy[0]; trans->y = xy[1];              /* untouched when the (0,0) start candidate wins (Q-F11) */
  trans->scale = so[0]; trans->offset = so[1];
  no = block_size_x * block_size_y;                /* the global compute_rms leaves behind; the cascade's tol * tol * no reads it (Q-F13) */
  g_calls++;
  return rms;
}
