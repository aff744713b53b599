/* v1_harness_tq.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Drives the UNMODIFIED version1 dct_luma (V1/src/block.c:836-1045: 4x4 forward transform, quantisation with
 * qp_const = (1 << q_bits) / 3, dequantisation, inverse transform, reconstruction) for batches of 4x4 blocks, in the
 * block layout of b2tq_4x4 (include/b2me.h).  block.c is compiled by oracle/Makefile.v1 from a scratch copy in which ONE
 * line is renamed (the file-local `static const int dequant_coef` collides with the extern declaration of
 * V1/inc/block.h under gcc); FIELD_SCAN / SNGL_SCAN / QP_SCALE_CR / dequant_coef are the reference's own
 * tables (the .rodata of its macroblock.o) and sign() its own function (block_dec.o, -ffunction-sections).
 */
#include "windows.h"
#include "global.h"
#include "i_global.h"
#include "image.h"
#include "mbuffer.h"
#include <stdint.h>

StorablePicture *enc_picture;
int dct_luma(int block_x, int block_y, int *coeff_cost, int old_intra_mode);
/* referenced by other functions of block.c, never reached from dct_luma */
void getNeighbour(int curr_mb_nr, int xN, int yN, int luma, PixelPos *pix) { (void)curr_mb_nr; (void)xN; (void)yN; (void)luma; (void)pix; abort(); }
void levrun_linfo_c2x2(int level, int run, int *len, int *info) { (void)level; (void)run; (void)len; (void)info; abort(); }
void levrun_linfo_inter(int level, int run, int *len, int *info) { (void)level; (void)run; (void)len; (void)info; abort(); }

static InputParameters h_input;
static ImageParameters h_img;
static int ready;

static void setup(void)
{
  int i, j, k;
  if (ready) return;
  input = &h_input; img = &h_img;
  memset(&h_input, 0, sizeof(h_input)); memset(&h_img, 0, sizeof(h_img));
  img->mb_data = (Macroblock *)calloc(1, sizeof(Macroblock));
  img->current_mb_nr = 0;
  img->cofAC = (int ****)calloc(4, sizeof(int ***));           /* [b8][b4][level/run][18], as get_mem_ACcoeff lays it out */
  for (i = 0; i < 4; i++) {
    img->cofAC[i] = (int ***)calloc(4, sizeof(int **));
    for (j = 0; j < 4; j++) {
      img->cofAC[i][j] = (int **)calloc(2, sizeof(int *));
      for (k = 0; k < 2; k++) img->cofAC[i][j][k] = (int *)calloc(18, sizeof(int));
    }
  }
  enc_picture = (StorablePicture *)calloc(1, sizeof(StorablePicture));
  enc_picture->imgY = (byte **)calloc(16, sizeof(byte *));
  for (i = 0; i < 16; i++) enc_picture->imgY[i] = (byte *)calloc(16, sizeof(byte));
  ready = 1;
}

/* nblk blocks; orig / pred / recon [nblk][16] raster u8; level / run [nblk][16] int32 (zero-terminated lists, zero-padded);
 * cost [nblk] = the increment dct_luma adds to *coeff_cost; nz [nblk] = its return value.  slice_type: img->type. */
void v1tq_dct_luma(int qp, int slice_type, int nblk, const uint8_t *orig, const uint8_t *pred,
                   int32_t *level, int32_t *run, uint8_t *recon, int32_t *cost, int32_t *nz)
{
  int b, x, y, k;
  setup();
  img->type = slice_type;
  img->mb_data[0].qp = qp;
  img->pix_x = img->pix_y = 0;
  UV_dct = 0;
  for (b = 0; b < nblk; b++) {
    int cc = 0;
    for (y = 0; y < 4; y++)
      for (x = 0; x < 4; x++) {
        img->mpr[x][y] = pred[b * 16 + y * 4 + x];                                     /* mpr[x][y], m7[x][y] (block.c:872, 994) */
        img->m7[x][y] = (int)orig[b * 16 + y * 4 + x] - (int)pred[b * 16 + y * 4 + x];
      }
    for (k = 0; k < 18; k++) { img->cofAC[0][0][0][k] = 0; img->cofAC[0][0][1][k] = 0; }
    nz[b] = dct_luma(0, 0, &cc, 0);
    cost[b] = cc;
    for (k = 0; k < 16; k++) { level[b * 16 + k] = img->cofAC[0][0][0][k]; run[b * 16 + k] = img->cofAC[0][0][1][k]; }
    for (y = 0; y < 4; y++)
      for (x = 0; x < 4; x++) recon[b * 16 + y * 4 + x] = (uint8_t)enc_picture->imgY[y][x];
  }
}
