// b2_ctx.h -- internal context of libb2me (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/b2me.h"

struct b2me_ctx {
  int device;
  int W, H, Wp, Hp, mbw, mbh, nmb, nrefs, R;
  size_t plane_size;            // Wp*Hp
  uint8_t *d_cur;               // [H][W]
  uint8_t *d_curc;              // [2][H/2][W/2] current chroma (b2me_set_cur_chroma; allocated on first use)
  uint8_t *d_refc;              // [nrefs][2][H/2][W/2] reference chroma (b2me_set_ref_chroma)
  uint8_t *d_planes;            // [nrefs][16][Hp][Wp]  index [yy*4+xx]
  uint8_t *d_stage;             // [H][W] staging for host uploads
  uint8_t *d_spl;               // [nrefs][16][Hq][Wq] search planes: integer picture with an edge-replicated pad of spad, 16 byte shifts
  int Wq, Hq, spad;
  CUtensorMap tmap_spl;         // 3-D tensor map over d_spl (x, y, ref), box = one window copy
  CUtensorMap *d_tmap_spl;      // its device copy (global memory, 64-byte aligned)
  // device mirrors for the host-pointer batch API
  int16_t *d_pred, *d_center, *d_mv_int, *d_mv_sub;
  long long *d_cost_int, *d_cost_sub;
  // pinned host staging for the single-block drop-in path
  int16_t *h_io16;              // [4][41][2]
  long long *h_io64;            // [2][41]
  int16_t *d_io16;
  long long *d_io64;
  int *d_errflag;
  uint16_t *d_sadtab; size_t sadtab_bytes;   // b2me_sad_table's device buffer
  // compact frame search (b2me_search_frame_best), allocated at the first call
  int16_t *d_pred_mb, *d_best_mv; int8_t *d_best_ref; long long *d_best_cost; int32_t *d_best_cost32;
  int *d_work;                  // k_sad_fs item counter
  int wp_apply[16], wp_weight[16], wp_offset[16], wp_denom[16];   // explicit weighted prediction per reference slot
  unsigned long long *d_stats;
  cudaStream_t stream;          // internal stream for host-pointer calls
  cudaStream_t stream_h2d, stream_d2h;   // copy streams of the banded host-pointer search
  cudaEvent_t ev_band[8];       // [0..3] predictors of band b are up, [4..7] band b is searched
  cudaEvent_t ev0, ev1;
  cudaEvent_t ev_planes; int planes_pending;   // a host-pointer b2me_set_ref left its plane build running on `stream`
  cudaEvent_t ev_copy;          // host-pointer uploads return when the copy (not the plane build behind it) is done
  int timing;
  double t_ms[4];
  int64_t t_n[4];
  int64_t launches;
  int sm_count;
  int fs_smem_bytes;
  char err[512];
};

namespace b2 {

struct FsArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *spl; int Wq, Hq, spad;       // search planes [nrefs][16][Hq][Wq]: plane s = integer picture (edge-replicated pad of spad) shifted left by s bytes
  int W, H, mbw, nrefs;
  int R;                 // half window (pel)
  int restrict_mode;     // get_search_range mode; -1: use sr_override for every partition
  int sr_override;
  int lambda_f;
  long long min_mcost;
  const int16_t *pred, *center;
  int16_t *mv_int; long long *cost_int;
  int mb_first, ref_first, refs_per_mb;   // item -> (mb, ref)
  int nitems;
  int abs_index;         // 1: arrays indexed ((mb*nrefs+ref)*41+p), 0: (item*41+p)
  unsigned long long part_mask;           // active partitions
  int *errflag;
  int *work_counter;     // device int, zero at launch: next (MB, ref) item to hand out
  int flags;             // bit 0: stage every window with the clamped (non-TMA) path (debug, B2ME_FS_NOTMA=1)
  int claim;             // tasks claimed per visit of a buffer's counter (amortises the claim / completion protocol)
  int one;               // 1 (a run-time constant the compiler cannot fold, see sad_fs.cu addmin2)
  unsigned long long *stats;   // optional: [0] exact re-evaluations, [1] window passes, [2] items
};

struct FsGeom {                  // window geometry of a search range (sad_fs.cu fs_geom)
  int pitch;                     // bytes per window row: == 32 or 96 (mod 128)
  int rows;                      // physical rows per copy (logical rows + 3)
  int copy_bytes;                // bytes per copy (multiple of 128)
  int slot_bytes;                // 4 copies = one window buffer
  int total;                     // dynamic shared memory (two buffers)
  int threads;                   // CTA size
};

struct SubArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, Hp, mbw, nrefs;
  int lambda_h, lambda_q, metric_h, metric_q;
  int start_hp, start_qp;
  int use_bound;         // 1: cost_int is the initial bound of the half-pel stage, 0: DISTBLK_MAX
  int full81;            // 1: full_sub_pel_motion_estimation (81 quarter-pel positions) instead of the half + quarter stages
  const int16_t *pred;
  const int16_t *mv_int; const long long *cost_int;
  int16_t *mv_sub; long long *cost_sub;
  int mb_first, ref_first, refs_per_mb, nitems, abs_index;
  unsigned long long part_mask;
};

struct McArgs {
  const uint8_t *cur; int cur_pitch;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, mbw, nmb, nrefs;
  const uint8_t *mb_mode, *b8mode; const int8_t *ref8; const int16_t *mv;
  uint8_t *orig_blk, *pred_blk;
};
cudaError_t launch_mc_luma(const McArgs &a, cudaStream_t s);
struct McMbArgs {                // luma + chroma, either list or both (k_mc_mb)
  const uint8_t *cur; int cur_pitch; const uint8_t *curc; const uint8_t *refc;
  const uint8_t *planes; size_t plane_size;
  int W, H, Wp, mbw, nmb, nrefs;
  const uint8_t *mb_mode, *b8mode, *pdir; const int8_t *ref8; const int16_t *mv0, *mv1;
  uint8_t *orig_y, *pred_y, *orig_c, *pred_c;
};
cudaError_t launch_mc_mb(const McMbArgs &a, cudaStream_t s);
cudaError_t launch_expand_pred(int n, const int16_t *pred_mb, int16_t *pred, int16_t *center, cudaStream_t s);
cudaError_t launch_select_gather(int nmb, int nrefs, const long long *cost, int ref_lambda, const int16_t *mv, int8_t *best_ref, long long *best_cost,
                                 int16_t *best_mv, int32_t *cost32, cudaStream_t s);
cudaError_t launch_subpel_planes(const uint8_t *luma, int pitch, int W, int H, uint8_t *planes16, int row_lo, int row_hi, cudaStream_t s);
cudaError_t launch_sad_fs(const FsArgs &a, const CUtensorMap *tm, int sm_count, cudaStream_t s, int *smem_bytes_out);
FsGeom fs_geom_host(int R);
// Stream-ordered scratch (cudaMallocAsync) is used by the host-pointer entries and by b2dbk_frame_dev.  By default the device's
// memory pool hands freed memory back to the driver at every synchronisation, which turns the next cudaMallocAsync into a real
// allocation on the host -- milliseconds, inside whatever the caller is timing.  Keep it in the pool (once per device).
void b2_pool_retain(int device);
cudaError_t launch_apply_wp(uint8_t *buf, size_t bytes, int weight, int offset, int log_denom, cudaStream_t s);
cudaError_t launch_search_plane(const uint8_t *luma, int pitch, int W, int H, uint8_t *out, int Wq, int Hq, int spad, int row_lo, int row_hi, cudaStream_t s);
cudaError_t launch_subpel_refine(const SubArgs &a, cudaStream_t s);
cudaError_t ubench(int kind, int iters, double *gops);
cudaError_t launch_distortion(int kind, int n, int nblk, const int16_t *diff, long long *out, cudaStream_t s);

}  // namespace b2
