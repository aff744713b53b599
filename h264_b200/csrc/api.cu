// api.cu -- the C ABI of libb2me.so (include/b2me.h): context, picture upload, search entry
// points.  No CPU fallback exists: every entry point either runs the CUDA path or returns an
// error code.
#include <cstdlib>
#include <cstring>
#include <new>
#include "b2_common.cuh"
#include "b2_ctx.h"

using namespace b2;

static char g_err[512] = "no context";

extern "C" int b2me_version(void) { return 100; }

extern "C" const char *b2me_last_error(b2me_ctx *ctx) { return ctx ? ctx->err : g_err; }

namespace b2 {
void b2_pool_retain(int device)
{
  static unsigned char done[64] = {0};
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[device] = 1;
}
}  // namespace b2

extern "C" int b2me_create(b2me_ctx **out, int device, int width, int height, int nrefs, int search_range)
{
  if (!out || width <= 0 || height <= 0 || (width & 15) || (height & 15) || nrefs < 1 || nrefs > 16 ||
      search_range < 1 || search_range > 64) {
    snprintf(g_err, sizeof(g_err), "b2me_create: invalid argument");
    return B2ME_EINVAL;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    snprintf(g_err, sizeof(g_err), "b2me_create: no CUDA device %d (%s)", device, cudaGetErrorString(e));
    return B2ME_ECUDA;
  }
  b2me_ctx *c = new (std::nothrow) b2me_ctx();
  if (!c) return B2ME_ENOMEM;
  memset(c, 0, sizeof(*c));
  c->device = device; c->W = width; c->H = height; c->nrefs = nrefs; c->R = search_range;
  c->Wp = width + 2 * PADX; c->Hp = height + 2 * PADY;
  c->mbw = width / 16; c->mbh = height / 16; c->nmb = c->mbw * c->mbh;
  c->plane_size = (size_t)c->Wp * c->Hp;
  *out = c;
  B2_CUDA_CHECK(c, cudaSetDevice(device));
  b2_pool_retain(device);
  B2_CUDA_CHECK(c, cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  const size_t n = (size_t)c->nmb * nrefs * NPART;
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_cur, (size_t)width * height));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_stage, (size_t)width * height));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_planes, c->plane_size * 16 * nrefs));
  B2_CUDA_CHECK(c, cudaMemset(c->d_planes, 0, c->plane_size * 16 * nrefs));
  // search planes of the integer full search + their TMA descriptor (box = one window copy)
  c->spad = ((search_range + 32) + 15) & ~15;
  c->Wq = width + 2 * c->spad; c->Hq = height + 2 * c->spad;
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_spl, (size_t)c->Wq * c->Hq * 16 * nrefs));
  B2_CUDA_CHECK(c, cudaMemset(c->d_spl, 0, (size_t)c->Wq * c->Hq * 16 * nrefs));
  {
    const FsGeom G = fs_geom_host(search_range);
    const cuuint64_t dims[3] = {(cuuint64_t)c->Wq, (cuuint64_t)c->Hq, (cuuint64_t)nrefs * 16};
    const cuuint64_t strides[2] = {(cuuint64_t)c->Wq, (cuuint64_t)c->Wq * c->Hq};
    const cuuint32_t box[3] = {(cuuint32_t)G.pitch, (cuuint32_t)G.rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // resolved through the runtime so that libb2me.so carries no link-time dependency on libcuda.so.1
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B2_CUDA_CHECK(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled is not available in this driver");
      return B2ME_ECUDA;
    }
    CUresult cr = reinterpret_cast<encode_fn>(fn)(&c->tmap_spl, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->d_spl, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled failed (%d)", (int)cr);
      return B2ME_ECUDA;
    }
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_tmap_spl, sizeof(CUtensorMap)));
    B2_CUDA_CHECK(c, cudaMemcpy(c->d_tmap_spl, &c->tmap_spl, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  }
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_pred, n * 2 * sizeof(int16_t)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_center, n * 2 * sizeof(int16_t)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_mv_int, n * 2 * sizeof(int16_t)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_mv_sub, n * 2 * sizeof(int16_t)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_cost_int, n * sizeof(long long)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_cost_sub, n * sizeof(long long)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_io16, 4 * NPART * 2 * sizeof(int16_t)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_io64, 2 * NPART * sizeof(long long)));
  B2_CUDA_CHECK(c, cudaMallocHost(&c->h_io16, 4 * NPART * 2 * sizeof(int16_t)));
  B2_CUDA_CHECK(c, cudaMallocHost(&c->h_io64, 2 * NPART * sizeof(long long)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_errflag, sizeof(int)));
  B2_CUDA_CHECK(c, cudaMemset(c->d_errflag, 0, sizeof(int)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_work, sizeof(int)));
  B2_CUDA_CHECK(c, cudaMalloc(&c->d_stats, 16 * sizeof(unsigned long long)));
  B2_CUDA_CHECK(c, cudaMemset(c->d_stats, 0, 16 * sizeof(unsigned long long)));
  B2_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  B2_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking));
  B2_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < 8; i++) B2_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_band[i], cudaEventDisableTiming));
  B2_CUDA_CHECK(c, cudaEventCreate(&c->ev0));
  B2_CUDA_CHECK(c, cudaEventCreate(&c->ev1));
  B2_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
  B2_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_planes, cudaEventDisableTiming));
  return B2ME_OK;
}

extern "C" void b2me_destroy(b2me_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_cur); cudaFree(c->d_curc); cudaFree(c->d_refc); cudaFree(c->d_stage); cudaFree(c->d_planes); cudaFree(c->d_spl); cudaFree(c->d_tmap_spl);
  cudaFree(c->d_pred); cudaFree(c->d_center); cudaFree(c->d_mv_int); cudaFree(c->d_mv_sub);
  cudaFree(c->d_cost_int); cudaFree(c->d_cost_sub); cudaFree(c->d_io16); cudaFree(c->d_io64);
  cudaFree(c->d_sadtab); cudaFree(c->d_pred_mb); cudaFree(c->d_best_ref); cudaFree(c->d_best_cost); cudaFree(c->d_best_cost32); cudaFree(c->d_best_mv);
  cudaFreeHost(c->h_io16); cudaFreeHost(c->h_io64); cudaFree(c->d_errflag); cudaFree(c->d_work); cudaFree(c->d_stats);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream_h2d) cudaStreamDestroy(c->stream_h2d);
  if (c->stream_d2h) cudaStreamDestroy(c->stream_d2h);
  for (int i = 0; i < 8; i++) if (c->ev_band[i]) cudaEventDestroy(c->ev_band[i]);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->ev_planes) cudaEventDestroy(c->ev_planes);
  delete c;
}

// ---- timing helper: device time of one kernel family on the launching stream --------------
struct FamilyTimer {
  b2me_ctx *c; int which; cudaStream_t s;
  FamilyTimer(b2me_ctx *c_, int w, cudaStream_t s_) : c(c_), which(w), s(s_) { if (c->timing) cudaEventRecord(c->ev0, s); }
  void stop() {
    if (!c->timing) return;
    cudaEventRecord(c->ev1, s); cudaEventSynchronize(c->ev1);
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->t_ms[which] += ms; c->t_n[which]++;
  }
};

extern "C" int b2me_kernel_timing(b2me_ctx *c, int enable)
{
  if (!c) return B2ME_EINVAL;
  c->timing = enable;
  for (int i = 0; i < 4; i++) { c->t_ms[i] = 0; c->t_n[i] = 0; }
  return B2ME_OK;
}
extern "C" int b2me_kernel_time_ms(b2me_ctx *c, int which, double *ms, int64_t *launches)
{
  if (!c || which < 0 || which > 3) return B2ME_EINVAL;
  if (ms) *ms = c->t_ms[which];
  if (launches) *launches = c->t_n[which];
  return B2ME_OK;
}
extern "C" int64_t b2me_launch_count(b2me_ctx *c) { return c ? c->launches : 0; }
extern "C" int b2me_search_stats(b2me_ctx *c, int64_t out[3], int reset)
{
  if (!c || !out) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  B2_CUDA_CHECK(c, cudaDeviceSynchronize());
  unsigned long long h[16];
  B2_CUDA_CHECK(c, cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
  out[0] = (int64_t)h[0]; out[1] = (int64_t)h[1]; out[2] = (int64_t)h[2];
  if (getenv("B2ME_FS_PROFILE"))
    fprintf(stderr, "[b2me] k_sad_fs warp-cycles: task %llu exact %llu idle %llu total %llu | producer %llu (busy %llu); claim %llu decode %llu post %llu; CTA0 warp0: %llu cycles in %llu ns = %.0f MHz; cold entries %llu, %llu warp-cycles in them\n", h[3], h[4], h[6], h[7], h[5], h[8], h[11], h[12], h[13], h[9], h[10], h[10] ? 1e3 * (double)h[9] / (double)h[10] : 0.0, h[14], h[15]);
  if (reset) B2_CUDA_CHECK(c, cudaMemset(c->d_stats, 0, sizeof(h)));
  return B2ME_OK;
}

// ---- pictures ------------------------------------------------------------------------------
extern "C" int b2me_set_cur_dev(b2me_ctx *c, const uint8_t *luma_dev, int stride, void *stream)
{
  if (!c || !luma_dev || stride < c->W) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  B2_CUDA_CHECK(c, cudaMemcpy2DAsync(c->d_cur, c->W, luma_dev, stride, c->W, c->H, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return B2ME_OK;
}
extern "C" int b2me_set_cur(b2me_ctx *c, const uint8_t *luma, int stride)
{
  if (!c || !luma || stride < c->W) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  B2_CUDA_CHECK(c, cudaMemcpy2DAsync(c->d_cur, c->W, luma, stride, c->W, c->H, cudaMemcpyHostToDevice, c->stream));
  B2_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return B2ME_OK;
}
// Work on a caller's stream that reads the reference planes is ordered after a plane build that a host-pointer
// b2me_set_ref left running on the context's own stream.
static int after_uploads(b2me_ctx *c, cudaStream_t s)
{
  if (c->planes_pending && s != c->stream) B2_CUDA_CHECK(c, cudaStreamWaitEvent(s, c->ev_planes, 0));
  return B2ME_OK;
}

static int build_planes(b2me_ctx *c, int ref_idx, const uint8_t *luma_dev, int stride, cudaStream_t s, int row_lo = 0, int row_hi = 1 << 30)
{
  FamilyTimer t(c, 1, s);
  B2_CUDA_CHECK(c, launch_subpel_planes(luma_dev, stride, c->W, c->H, c->d_planes + (size_t)ref_idx * 16 * c->plane_size, row_lo, row_hi, s));
  B2_CUDA_CHECK(c, launch_search_plane(luma_dev, stride, c->W, c->H, c->d_spl + (size_t)ref_idx * 16 * c->Wq * c->Hq, c->Wq, c->Hq, c->spad, row_lo, row_hi, s));
  c->launches += 3;
  if (c->wp_apply[ref_idx]) {          // weighted reference: map the planes the distortions read (see k_apply_wp)
    B2_CUDA_CHECK(c, launch_apply_wp(c->d_planes + (size_t)ref_idx * 16 * c->plane_size, c->plane_size * 16,
                                     c->wp_weight[ref_idx], c->wp_offset[ref_idx], c->wp_denom[ref_idx], s));
    B2_CUDA_CHECK(c, launch_apply_wp(c->d_spl + (size_t)ref_idx * 16 * c->Wq * c->Hq, (size_t)c->Wq * c->Hq * 16,
                                     c->wp_weight[ref_idx], c->wp_offset[ref_idx], c->wp_denom[ref_idx], s));
    c->launches += 2;
  }
  t.stop();
  return B2ME_OK;
}
extern "C" int b2me_set_ref_weights(b2me_ctx *c, int ref_idx, int apply, int weight, int offset, int log_weight_denom)
{
  if (!c || ref_idx < 0 || ref_idx >= c->nrefs || log_weight_denom < 0 || log_weight_denom > 7 ||
      weight < -128 || weight > 127 || offset < -128 || offset > 127) return B2ME_EINVAL;
  c->wp_apply[ref_idx] = apply ? 1 : 0; c->wp_weight[ref_idx] = weight; c->wp_offset[ref_idx] = offset; c->wp_denom[ref_idx] = log_weight_denom;
  return B2ME_OK;
}
extern "C" int b2me_set_ref_dev(b2me_ctx *c, int ref_idx, const uint8_t *luma_dev, int stride, void *stream)
{
  if (!c || !luma_dev || stride < c->W || ref_idx < 0 || ref_idx >= c->nrefs) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r0 = after_uploads(c, (cudaStream_t)stream); if (r0) return r0; }
  return build_planes(c, ref_idx, luma_dev, stride, (cudaStream_t)stream);
}
extern "C" int b2me_set_ref_rows_dev(b2me_ctx *c, int ref_idx, const uint8_t *luma_dev, int stride, int row_first, int row_count, void *stream)
{
  if (!c || !luma_dev || stride < c->W || ref_idx < 0 || ref_idx >= c->nrefs || row_first < 0 || row_count < 0 || row_first + row_count > c->H) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (c->wp_apply[ref_idx] && !(row_first == 0 && row_count == c->H)) {
    // k_apply_wp maps the slot's whole plane set in place: a partial rebuild would weight the untouched rows a second time
    snprintf(c->err, sizeof(c->err), "b2me_set_ref_rows_dev: reference slot %d is weighted (b2me_set_ref_weights); upload it whole", ref_idx);
    return B2ME_EUNSUPPORTED;
  }
  { int r0 = after_uploads(c, (cudaStream_t)stream); if (r0) return r0; }
  return build_planes(c, ref_idx, luma_dev, stride, (cudaStream_t)stream, row_first, row_first + row_count);
}
extern "C" int b2me_set_ref(b2me_ctx *c, int ref_idx, const uint8_t *luma, int stride)
{
  if (!c || !luma || stride < c->W || ref_idx < 0 || ref_idx >= c->nrefs) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  B2_CUDA_CHECK(c, cudaMemcpy2DAsync(c->d_stage, c->W, luma, stride, c->W, c->H, cudaMemcpyHostToDevice, c->stream));
  B2_CUDA_CHECK(c, cudaEventRecord(c->ev_copy, c->stream));
  int r = build_planes(c, ref_idx, c->d_stage, c->W, c->stream);
  if (r) return r;
  B2_CUDA_CHECK(c, cudaEventRecord(c->ev_planes, c->stream));
  c->planes_pending = 1;
  // the caller's buffer is free once the copy is done; the plane kernels run on behind it (every later call of this
  // context runs on c->stream or waits for ev_planes, see after_uploads)
  B2_CUDA_CHECK(c, cudaEventSynchronize(c->ev_copy));
  return B2ME_OK;
}
extern "C" int b2me_get_subplane(b2me_ctx *c, int ref_idx, int yy, int xx, uint8_t *out)
{
  if (!c || !out || ref_idx < 0 || ref_idx >= c->nrefs || yy < 0 || yy > 3 || xx < 0 || xx > 3) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  B2_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  B2_CUDA_CHECK(c, cudaMemcpy(out, c->d_planes + ((size_t)ref_idx * 16 + yy * 4 + xx) * c->plane_size, c->plane_size, cudaMemcpyDeviceToHost));
  return B2ME_OK;
}

// ---- search --------------------------------------------------------------------------------
static int check_params(b2me_ctx *c, const b2me_search_params *p)
{
  if (!p) return B2ME_EINVAL;
  if (p->restrict_mode < 0 || p->restrict_mode > 2) return B2ME_EINVAL;
  if (p->do_subpel && (p->metric_h < 0 || p->metric_h > 2 || p->metric_q < 0 || p->metric_q > 2)) {
    snprintf(c->err, sizeof(c->err), "sub-pel metric must be 0 (SAD), 1 (SSE) or 2 (SATD)");
    return B2ME_EINVAL;
  }
  return B2ME_OK;
}

static int run_subpel(b2me_ctx *c, int mb_first, int mb_count, int ref_first, int refs_per_mb, int abs_index,
                      unsigned long long mask, const int16_t *pred, const b2me_search_params *p,
                      const int16_t *mv_int, const long long *cost_int, int16_t *mv_sub, long long *cost_sub,
                      int use_bound, cudaStream_t s)
{
  { int r0 = after_uploads(c, s); if (r0) return r0; }
  SubArgs q;
  q.cur = c->d_cur; q.cur_pitch = c->W; q.planes = c->d_planes; q.plane_size = c->plane_size;
  q.W = c->W; q.H = c->H; q.Wp = c->Wp; q.Hp = c->Hp; q.mbw = c->mbw; q.nrefs = c->nrefs;
  q.lambda_h = p->lambda_factor[1]; q.lambda_q = p->lambda_factor[2]; q.metric_h = p->metric_h; q.metric_q = p->metric_q;
  // p_Vid->start_me_refinement_hp/_qp (mv_search.c:445-446), ChromaME off, F_PEL metric = SAD
  q.start_hp = (0 != p->metric_h) ? 0 : 1;
  q.start_qp = (p->metric_h != p->metric_q) ? 0 : 1;
  // BlockMotionSearch resets the bound to DISTBLK_MAX when start_me_refinement_hp == 0 (mv_search.c:971-974);
  // the single-call drop-in receives the caller's bound instead
  q.use_bound = use_bound >= 0 ? use_bound : q.start_hp;
  q.full81 = p->subpel_full ? 1 : 0;
  q.pred = pred; q.mv_int = mv_int; q.cost_int = cost_int; q.mv_sub = mv_sub; q.cost_sub = cost_sub;
  q.mb_first = mb_first; q.ref_first = ref_first; q.refs_per_mb = refs_per_mb; q.nitems = mb_count * refs_per_mb;
  q.abs_index = abs_index; q.part_mask = mask;
  FamilyTimer t(c, 2, s);
  B2_CUDA_CHECK(c, launch_subpel_refine(q, s));
  c->launches++;
  t.stop();
  return B2ME_OK;
}

static int run_search(b2me_ctx *c, int mb_first, int mb_count, int ref_first, int refs_per_mb, int abs_index,
                      unsigned long long mask, int sr_override,
                      const int16_t *pred, const int16_t *center, const b2me_search_params *p,
                      int16_t *mv_int, long long *cost_int, int16_t *mv_sub, long long *cost_sub, cudaStream_t s)
{
  { int r0 = after_uploads(c, s); if (r0) return r0; }
  FsArgs f;
  f.cur = c->d_cur; f.cur_pitch = c->W; f.spl = c->d_spl; f.Wq = c->Wq; f.Hq = c->Hq; f.spad = c->spad;
  f.W = c->W; f.H = c->H; f.mbw = c->mbw; f.nrefs = c->nrefs;
  f.R = c->R; f.restrict_mode = sr_override >= 0 ? -1 : p->restrict_mode; f.sr_override = sr_override;
  f.lambda_f = p->lambda_factor[0]; f.min_mcost = p->min_mcost;
  f.pred = pred; f.center = center; f.mv_int = mv_int; f.cost_int = cost_int;
  f.mb_first = mb_first; f.ref_first = ref_first; f.refs_per_mb = refs_per_mb; f.nitems = mb_count * refs_per_mb;
  f.abs_index = abs_index; f.part_mask = mask; f.errflag = c->d_errflag; f.stats = c->d_stats; f.one = 1;
  { const char *e = getenv("B2ME_FS_NOTMA"); f.flags = (e && e[0] == '1') ? 1 : 0;
    const char *r = getenv("B2ME_FS_REP"); if (r) f.flags |= atoi(r) << 8;
    const char *z = getenv("B2ME_FS_SLEEP"); if (z && z[0] == '1') f.flags |= 2;
    const char *sd = getenv("B2ME_FS_NOSEED"); if (sd && sd[0] == '1') f.flags |= 4;
    const char *nb = getenv("B2ME_FS_NOB"); if (nb && nb[0] == '1') f.flags |= 8;      /* timing probe: drops the type-B tasks (WRONG results) */
    const char *k = getenv("B2ME_FS_CLAIM"); f.claim = k ? atoi(k) : (fs_geom_host(f.R).pitch == 96 ? 1 : 3); if (f.claim < 1) f.claim = 1; }   // measured: 18 tasks per unit and 4 workers want single claims (the tail), 68 tasks and 12 workers claims of 3
  f.work_counter = c->d_work;
  B2_CUDA_CHECK(c, cudaMemsetAsync(c->d_work, 0, sizeof(int), s));
  {
    FamilyTimer t(c, 0, s);
    B2_CUDA_CHECK(c, launch_sad_fs(f, c->d_tmap_spl, c->sm_count, s, &c->fs_smem_bytes));
    c->launches++;
    t.stop();
  }
  if (p->do_subpel)
    return run_subpel(c, mb_first, mb_count, ref_first, refs_per_mb, abs_index, mask, pred, p, mv_int, cost_int, mv_sub, cost_sub, -1, s);
  return B2ME_OK;
}

extern "C" int b2me_search_mbs_dev(b2me_ctx *c, int mb_first, int mb_count, const int16_t *pred, const int16_t *center,
                                   const b2me_search_params *p, int16_t *mv_int, int64_t *cost_int,
                                   int16_t *mv_sub, int64_t *cost_sub, void *stream)
{
  if (!c || !pred || !center || !mv_int || !cost_int) return B2ME_EINVAL;
  int r = check_params(c, p);
  if (r) return r;
  if (p->do_subpel && (!mv_sub || !cost_sub)) return B2ME_EINVAL;
  if (mb_first < 0 || mb_count < 0 || mb_first + mb_count > c->nmb) return B2ME_EINVAL;
  if (mb_count == 0) return B2ME_OK;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  return run_search(c, mb_first, mb_count, 0, c->nrefs, 1, (1ull << NPART) - 1, -1, pred, center, p, mv_int,
                    (long long *)cost_int, mv_sub, (long long *)cost_sub, (cudaStream_t)stream);
}

extern "C" int b2me_search_frame_dev(b2me_ctx *c, const int16_t *pred, const int16_t *center, const b2me_search_params *p,
                                     int16_t *mv_int, int64_t *cost_int, int16_t *mv_sub, int64_t *cost_sub, void *stream)
{
  if (!c) return B2ME_EINVAL;
  return b2me_search_mbs_dev(c, 0, c->nmb, pred, center, p, mv_int, cost_int, mv_sub, cost_sub, stream);
}

static int check_errflag(b2me_ctx *c, cudaStream_t s);
// The _dev searches never synchronise, so the device-side input check (a search centre that is not integer-pel) is read here.
extern "C" int b2me_check_errors(b2me_ctx *c, void *stream)
{
  if (!c) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  return check_errflag(c, (cudaStream_t)stream);
}

static int check_errflag(b2me_ctx *c, cudaStream_t s)
{
  int flag = 0;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaStreamSynchronize(s));
  if (flag) {
    cudaMemsetAsync(c->d_errflag, 0, sizeof(int), s);
    snprintf(c->err, sizeof(c->err), "search centre is not integer-pel (quarter-pel centres are not a full-search input)");
    return B2ME_EINVAL;
  }
  return B2ME_OK;
}

extern "C" int b2me_search_frame(b2me_ctx *c, const int16_t *pred, const int16_t *center, const b2me_search_params *p,
                                 int16_t *mv_int, int64_t *cost_int, int16_t *mv_sub, int64_t *cost_sub)
{
  if (!c || !pred || !center) return B2ME_EINVAL;
  int r = check_params(c, p);
  if (r) return r;
  if (p->do_subpel && (!mv_sub || !cost_sub)) return B2ME_EINVAL;
  if ((!mv_int) != (!cost_int) || (!mv_int && !p->do_subpel)) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  // The picture is processed in bands of MB rows so that the host<->device copies overlap the search: predictors
  // of band i+1 go up (H2D stream) and results of band i-1 come down (D2H stream) while band i is searched.
  int nband = c->mbh >= 16 ? 2 : 1;               // measured (tools/e2e_probe.py, 1080p x 4 refs): 1 band 2.07 ms, 2: 2.02, 3: 2.08, 4: 2.20
  { const char *e = getenv("B2ME_BANDS"); if (e && atoi(e) >= 1 && atoi(e) <= 4 && c->mbh >= 16) nband = atoi(e); }   // development probe
  cudaStream_t s = c->stream;
  for (int b = 0; b < nband; b++) {
    const int row0 = (int)((long long)c->mbh * b / nband), row1 = (int)((long long)c->mbh * (b + 1) / nband);
    const int mb0 = row0 * c->mbw, cnt = (row1 - row0) * c->mbw;
    const size_t o2 = (size_t)mb0 * c->nrefs * NPART * 2, n2 = (size_t)cnt * c->nrefs * NPART * 2;
    B2_CUDA_CHECK(c, cudaMemcpyAsync(c->d_pred + o2, pred + o2, n2 * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream_h2d));
    B2_CUDA_CHECK(c, cudaMemcpyAsync(c->d_center + o2, center + o2, n2 * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream_h2d));
    B2_CUDA_CHECK(c, cudaEventRecord(c->ev_band[b], c->stream_h2d));
  }
  for (int b = 0; b < nband; b++) {
    const int row0 = (int)((long long)c->mbh * b / nband), row1 = (int)((long long)c->mbh * (b + 1) / nband);
    const int mb0 = row0 * c->mbw, cnt = (row1 - row0) * c->mbw;
    const size_t o1 = (size_t)mb0 * c->nrefs * NPART, n1 = (size_t)cnt * c->nrefs * NPART;
    B2_CUDA_CHECK(c, cudaStreamWaitEvent(s, c->ev_band[b], 0));
    r = run_search(c, mb0, cnt, 0, c->nrefs, 1, (1ull << NPART) - 1, -1, c->d_pred, c->d_center, p,
                   c->d_mv_int, c->d_cost_int, c->d_mv_sub, c->d_cost_sub, s);
    if (r) return r;
    B2_CUDA_CHECK(c, cudaEventRecord(c->ev_band[4 + b], s));
    B2_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream_d2h, c->ev_band[4 + b], 0));
    if (mv_int) {
      B2_CUDA_CHECK(c, cudaMemcpyAsync(mv_int + o1 * 2, c->d_mv_int + o1 * 2, n1 * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream_d2h));
      B2_CUDA_CHECK(c, cudaMemcpyAsync(cost_int + o1, c->d_cost_int + o1, n1 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream_d2h));
    }
    if (p->do_subpel) {
      B2_CUDA_CHECK(c, cudaMemcpyAsync(mv_sub + o1 * 2, c->d_mv_sub + o1 * 2, n1 * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream_d2h));
      B2_CUDA_CHECK(c, cudaMemcpyAsync(cost_sub + o1, c->d_cost_sub + o1, n1 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream_d2h));
    }
  }
  B2_CUDA_CHECK(c, cudaStreamSynchronize(c->stream_d2h));
  return check_errflag(c, s);
}

// Compact whole-frame search: ONE predictor per (MB, ref) up, the mode decision's view down (see include/b2me.h).  Everything in
// between stays on the device; one synchronisation.
extern "C" int b2me_search_frame_best(b2me_ctx *c, const int16_t *pred_mb, const b2me_search_params *p, int ref_lambda,
                                      int8_t *best_ref, int32_t *best_cost, int16_t *best_mv)
{
  if (!c || !pred_mb || !best_ref || !best_cost || !best_mv) return B2ME_EINVAL;
  int r = check_params(c, p);
  if (r) return r;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  const size_t nitem = (size_t)c->nmb * c->nrefs;
  if (!c->d_pred_mb) {
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_pred_mb, nitem * 2 * sizeof(int16_t)));
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_best_ref, (size_t)c->nmb * 21));
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_best_cost, (size_t)c->nmb * 21 * sizeof(long long)));
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_best_cost32, (size_t)c->nmb * 21 * sizeof(int32_t)));
    B2_CUDA_CHECK(c, cudaMalloc(&c->d_best_mv, (size_t)c->nmb * NPART * 2 * sizeof(int16_t)));
  }
  cudaStream_t s = c->stream;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(c->d_pred_mb, pred_mb, nitem * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, s));
  B2_CUDA_CHECK(c, launch_expand_pred((int)(nitem * NPART), c->d_pred_mb, c->d_pred, c->d_center, s));
  c->launches++;
  r = run_search(c, 0, c->nmb, 0, c->nrefs, 1, (1ull << NPART) - 1, -1, c->d_pred, c->d_center, p,
                 c->d_mv_int, c->d_cost_int, c->d_mv_sub, c->d_cost_sub, s);
  if (r) return r;
  B2_CUDA_CHECK(c, launch_select_gather(c->nmb, c->nrefs, p->do_subpel ? c->d_cost_sub : c->d_cost_int, ref_lambda,
                                        p->do_subpel ? c->d_mv_sub : c->d_mv_int, c->d_best_ref, c->d_best_cost, c->d_best_mv, c->d_best_cost32, s));
  c->launches += 2;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(best_ref, c->d_best_ref, (size_t)c->nmb * 21, cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(best_cost, c->d_best_cost32, (size_t)c->nmb * 21 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(best_mv, c->d_best_mv, (size_t)c->nmb * NPART * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
  return check_errflag(c, s);
}

extern "C" int b2me_block_search(b2me_ctx *c, int pos_x, int pos_y, int blocktype, int ref_idx,
                                 const int16_t pred_mv[2], const int16_t center_mv[2], const b2me_search_params *p,
                                 int search_range_pel, int16_t mv_int[2], int64_t *cost_int, int16_t mv_sub[2], int64_t *cost_sub)
{
  if (!c || !pred_mv || !center_mv || !mv_int || !cost_int) return B2ME_EINVAL;
  int r = check_params(c, p);
  if (r) return r;
  if (p->do_subpel && (!mv_sub || !cost_sub)) return B2ME_EINVAL;
  if (blocktype < 1 || blocktype > 7 || ref_idx < 0 || ref_idx >= c->nrefs || search_range_pel < 0 || search_range_pel > c->R ||
      pos_x < 0 || pos_y < 0 || pos_x >= c->W || pos_y >= c->H) return B2ME_EINVAL;
  // locate the partition of this blocktype at (pos_x & 15, pos_y & 15)
  int part = -1;
  for (int q = part_first(blocktype); q < NPART; q++) {
    PartGeom g = part_geom(q);
    if (g.bt != blocktype) break;
    if (g.ox == (pos_x & 15) && g.oy == (pos_y & 15)) { part = q; break; }
  }
  if (part < 0) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  int16_t *h16 = c->h_io16; long long *h64 = c->h_io64;
  h16[part * 2] = pred_mv[0]; h16[part * 2 + 1] = pred_mv[1];
  h16[(NPART + part) * 2] = center_mv[0]; h16[(NPART + part) * 2 + 1] = center_mv[1];
  B2_CUDA_CHECK(c, cudaMemcpyAsync(c->d_io16, h16, 2 * NPART * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, s));
  const int mb = (pos_y >> 4) * c->mbw + (pos_x >> 4);
  r = run_search(c, mb, 1, ref_idx, 1, 0, 1ull << part, search_range_pel, c->d_io16, c->d_io16 + NPART * 2, p,
                 c->d_io16 + 2 * NPART * 2, c->d_io64, c->d_io16 + 3 * NPART * 2, c->d_io64 + NPART, s);
  if (r) return r;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(h16 + 2 * NPART * 2, c->d_io16 + 2 * NPART * 2, 2 * NPART * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(h64, c->d_io64, 2 * NPART * sizeof(long long), cudaMemcpyDeviceToHost, s));
  r = check_errflag(c, s);
  if (r) return r;
  mv_int[0] = h16[(2 * NPART + part) * 2]; mv_int[1] = h16[(2 * NPART + part) * 2 + 1];
  *cost_int = h64[part];
  if (p->do_subpel) {
    mv_sub[0] = h16[(3 * NPART + part) * 2]; mv_sub[1] = h16[(3 * NPART + part) * 2 + 1];
    *cost_sub = h64[NPART + part];
  }
  return B2ME_OK;
}

// Drop-in for ONE call of sub_pel_motion_estimation: start MV mv_in (quarter-pel, relative), bound
// min_mcost exactly as BlockMotionSearch hands it over (mv_search.c:967-976).
extern "C" int b2me_block_subpel(b2me_ctx *c, int pos_x, int pos_y, int blocktype, int ref_idx,
                                 const int16_t pred_mv[2], const int16_t mv_in[2], const b2me_search_params *p,
                                 int64_t min_mcost, int16_t mv_out[2], int64_t *cost_out)
{
  if (!c || !pred_mv || !mv_in || !mv_out || !cost_out) return B2ME_EINVAL;
  int r = check_params(c, p);
  if (r) return r;
  if (blocktype < 1 || blocktype > 7 || ref_idx < 0 || ref_idx >= c->nrefs || pos_x < 0 || pos_y < 0 || pos_x >= c->W || pos_y >= c->H) return B2ME_EINVAL;
  int part = -1;
  for (int q = part_first(blocktype); q < NPART; q++) {
    PartGeom g = part_geom(q);
    if (g.bt != blocktype) break;
    if (g.ox == (pos_x & 15) && g.oy == (pos_y & 15)) { part = q; break; }
  }
  if (part < 0) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  int16_t *h16 = c->h_io16; long long *h64 = c->h_io64;
  h16[part * 2] = pred_mv[0]; h16[part * 2 + 1] = pred_mv[1];
  h16[(2 * NPART + part) * 2] = mv_in[0]; h16[(2 * NPART + part) * 2 + 1] = mv_in[1];
  h64[part] = min_mcost;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(c->d_io16, h16, 3 * NPART * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, s));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(c->d_io64, h64, NPART * sizeof(long long), cudaMemcpyHostToDevice, s));
  const int mb = (pos_y >> 4) * c->mbw + (pos_x >> 4);
  r = run_subpel(c, mb, 1, ref_idx, 1, 0, 1ull << part, c->d_io16, p, c->d_io16 + 2 * NPART * 2, c->d_io64,
                 c->d_io16 + 3 * NPART * 2, c->d_io64 + NPART, 1, s);
  if (r) return r;
  B2_CUDA_CHECK(c, cudaMemcpyAsync(h16 + 3 * NPART * 2, c->d_io16 + 3 * NPART * 2, NPART * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaMemcpyAsync(h64 + NPART, c->d_io64 + NPART, NPART * sizeof(long long), cudaMemcpyDeviceToHost, s));
  B2_CUDA_CHECK(c, cudaStreamSynchronize(s));
  mv_out[0] = h16[(3 * NPART + part) * 2]; mv_out[1] = h16[(3 * NPART + part) * 2 + 1];
  *cost_out = h64[NPART + part];
  return B2ME_OK;
}

// luma_prediction (list 0) of the whole picture from a search result array; device pointers
extern "C" int b2me_mc_luma_dev(b2me_ctx *c, const uint8_t *mb_mode, const uint8_t *b8mode, const int8_t *ref8, const int16_t *mv,
                                uint8_t *orig_blk, uint8_t *pred_blk, void *stream)
{
  if (!c || !mb_mode || !b8mode || !ref8 || !mv || !orig_blk || !pred_blk) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r0 = after_uploads(c, (cudaStream_t)stream); if (r0) return r0; }
  McArgs m;
  m.cur = c->d_cur; m.cur_pitch = c->W; m.planes = c->d_planes; m.plane_size = c->plane_size;
  m.W = c->W; m.H = c->H; m.Wp = c->Wp; m.mbw = c->mbw; m.nmb = c->nmb; m.nrefs = c->nrefs;
  m.mb_mode = mb_mode; m.b8mode = b8mode; m.ref8 = ref8; m.mv = mv; m.orig_blk = orig_blk; m.pred_blk = pred_blk;
  B2_CUDA_CHECK(c, launch_mc_luma(m, (cudaStream_t)stream));
  c->launches++;
  return B2ME_OK;
}

// ---- chroma planes (4:2:0) for the prediction of whole macroblocks, b2me_mc_mb_dev ----
static int ensure_chroma(b2me_ctx *c)
{
  const size_t n = (size_t)(c->W / 2) * (c->H / 2) * 2;
  if (!c->d_curc) B2_CUDA_CHECK(c, cudaMalloc(&c->d_curc, n));
  if (!c->d_refc) B2_CUDA_CHECK(c, cudaMalloc(&c->d_refc, n * c->nrefs));
  return B2ME_OK;
}
static int set_chroma(b2me_ctx *c, uint8_t *dst, const uint8_t *u, const uint8_t *v, int stride, cudaMemcpyKind kind, cudaStream_t s)
{
  const int Wc = c->W / 2, Hc = c->H / 2;
  B2_CUDA_CHECK(c, cudaMemcpy2DAsync(dst, Wc, u, stride, Wc, Hc, kind, s));
  B2_CUDA_CHECK(c, cudaMemcpy2DAsync(dst + (size_t)Wc * Hc, Wc, v, stride, Wc, Hc, kind, s));
  return B2ME_OK;
}
extern "C" int b2me_set_cur_chroma_dev(b2me_ctx *c, const uint8_t *u_dev, const uint8_t *v_dev, int stride, void *stream)
{
  if (!c || !u_dev || !v_dev || stride < c->W / 2) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r = ensure_chroma(c); if (r) return r; }
  return set_chroma(c, c->d_curc, u_dev, v_dev, stride, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
extern "C" int b2me_set_ref_chroma_dev(b2me_ctx *c, int ref, const uint8_t *u_dev, const uint8_t *v_dev, int stride, void *stream)
{
  if (!c || !u_dev || !v_dev || ref < 0 || ref >= c->nrefs || stride < c->W / 2) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r = ensure_chroma(c); if (r) return r; }
  return set_chroma(c, c->d_refc + (size_t)ref * (c->W / 2) * (c->H / 2) * 2, u_dev, v_dev, stride, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
extern "C" int b2me_set_cur_chroma(b2me_ctx *c, const uint8_t *u, const uint8_t *v, int stride)
{
  if (!c || !u || !v || stride < c->W / 2) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r = ensure_chroma(c); if (r) return r; }
  { int r = set_chroma(c, c->d_curc, u, v, stride, cudaMemcpyHostToDevice, c->stream); if (r) return r; }
  B2_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return B2ME_OK;
}
extern "C" int b2me_set_ref_chroma(b2me_ctx *c, int ref, const uint8_t *u, const uint8_t *v, int stride)
{
  if (!c || !u || !v || ref < 0 || ref >= c->nrefs || stride < c->W / 2) return B2ME_EINVAL;
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r = ensure_chroma(c); if (r) return r; }
  { int r = set_chroma(c, c->d_refc + (size_t)ref * (c->W / 2) * (c->H / 2) * 2, u, v, stride, cudaMemcpyHostToDevice, c->stream); if (r) return r; }
  B2_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return B2ME_OK;
}
extern "C" int b2me_mc_mb_dev(b2me_ctx *c, const uint8_t *mb_mode, const uint8_t *b8mode, const uint8_t *pdir, const int8_t *ref8, const int16_t *mv_l0,
                              const int16_t *mv_l1, uint8_t *orig_y, uint8_t *pred_y, uint8_t *orig_c, uint8_t *pred_c, void *stream)
{
  if (!c || !mb_mode || !b8mode || !pdir || !ref8 || !mv_l0 || !mv_l1 || !orig_y || !pred_y || !orig_c || !pred_c) return B2ME_EINVAL;
  if (!c->d_curc || !c->d_refc) { snprintf(c->err, sizeof(c->err), "b2me_mc_mb_dev: no chroma planes (b2me_set_cur_chroma / b2me_set_ref_chroma)"); return B2ME_EINVAL; }
  B2_CUDA_CHECK(c, cudaSetDevice(c->device));
  { int r0 = after_uploads(c, (cudaStream_t)stream); if (r0) return r0; }
  McMbArgs m;
  m.cur = c->d_cur; m.cur_pitch = c->W; m.curc = c->d_curc; m.refc = c->d_refc; m.planes = c->d_planes; m.plane_size = c->plane_size;
  m.W = c->W; m.H = c->H; m.Wp = c->Wp; m.mbw = c->mbw; m.nmb = c->nmb; m.nrefs = c->nrefs;
  m.mb_mode = mb_mode; m.b8mode = b8mode; m.pdir = pdir; m.ref8 = ref8; m.mv0 = mv_l0; m.mv1 = mv_l1;
  m.orig_y = orig_y; m.pred_y = pred_y; m.orig_c = orig_c; m.pred_c = pred_c;
  B2_CUDA_CHECK(c, launch_mc_mb(m, (cudaStream_t)stream));
  c->launches++;
  return B2ME_OK;
}

// distortion4x4/8x8{SAD,SSE,SATD} of the mode decision (me_distortion.c:38-134) for nblk difference blocks
extern "C" int b2me_distortion_blocks_dev(int kind, int n, int nblk, const int16_t *diff_dev, int64_t *out_dev, void *stream)
{
  if (kind < 0 || kind > 2 || (n != 4 && n != 8) || nblk < 0 || !diff_dev || !out_dev) return B2ME_EINVAL;
  if (nblk == 0) return B2ME_OK;
  cudaError_t e = launch_distortion(kind, n, nblk, diff_dev, (long long *)out_dev, (cudaStream_t)stream);
  if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "b2me_distortion_blocks: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  return B2ME_OK;
}
extern "C" int b2me_distortion_blocks(int device, int kind, int n, int nblk, const int16_t *diff, int64_t *out)
{
  if (kind < 0 || kind > 2 || (n != 4 && n != 8) || nblk < 0 || !diff || !out) return B2ME_EINVAL;
  if (nblk == 0) return B2ME_OK;
  if (cudaSetDevice(device) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "b2me_distortion_blocks: no CUDA device %d", device); return B2ME_ECUDA; }
  int16_t *d = nullptr; long long *o = nullptr;
  cudaError_t e = cudaMalloc(&d, (size_t)nblk * n * n * sizeof(int16_t));
  if (e == cudaSuccess) e = cudaMalloc(&o, (size_t)nblk * sizeof(long long));
  if (e == cudaSuccess) e = cudaMemcpy(d, diff, (size_t)nblk * n * n * sizeof(int16_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_distortion(kind, n, nblk, d, o, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out, o, (size_t)nblk * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d); cudaFree(o);
  if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "b2me_distortion_blocks: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  return B2ME_OK;
}

extern "C" int b2me_ubench(int device, int kind, int iters, double *gops)
{
  if (!gops || iters <= 0) return B2ME_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return B2ME_ECUDA;
  cudaError_t e = ubench(kind, iters, gops);
  if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "ubench: %s", cudaGetErrorString(e)); return B2ME_ECUDA; }
  return B2ME_OK;
}
