"""Host-side mirror of the C ABI in include/b2me.h (ctypes; no torch types cross the boundary).

The product path is libb2me.so (hand-written sm_100a CUDA).  There is no CPU fallback: if the
library is missing or a call fails this module raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2ME_LIB") or os.path.join(_HERE, "libb2me.so")   # B2ME_LIB: development builds (tools/)
NPART = 41
DISTBLK_MAX = (2**31 - 1) << 5
PAD_X, PAD_Y = 32, 20
_vp = C.c_void_p


class B2Error(RuntimeError):
    pass


class SearchParams(C.Structure):
    _fields_ = [("lambda_factor", C.c_int32 * 3), ("restrict_mode", C.c_int32), ("metric_h", C.c_int32),
                ("metric_q", C.c_int32), ("do_subpel", C.c_int32), ("subpel_full", C.c_int32),
                ("min_mcost", C.c_int64)]


def make_params(lambda_factor, restrict_mode=2, metric_h=2, metric_q=2, do_subpel=True, min_mcost=DISTBLK_MAX, subpel_full=False):
    p = SearchParams()
    lam = [int(x) for x in np.broadcast_to(np.asarray(lambda_factor), (3,))]
    p.lambda_factor[0], p.lambda_factor[1], p.lambda_factor[2] = lam
    p.restrict_mode, p.metric_h, p.metric_q = restrict_mode, metric_h, metric_q
    p.do_subpel, p.subpel_full, p.min_mcost = int(bool(do_subpel)), int(bool(subpel_full)), int(min_mcost)
    return p


_lib = None


def lib():
    """Load libb2me.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2Error(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                          "(make -C h264_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.b2me_last_error.restype = C.c_char_p
        L.b2me_last_error.argtypes = [_vp]
        L.b2me_launch_count.restype = C.c_int64
        L.b2me_launch_count.argtypes = [_vp]
        L.b2me_destroy.argtypes = [_vp]
        L.b2me_destroy.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def _dp(t):
    """device pointer of a torch CUDA tensor (or None)"""
    return _vp(0) if t is None else _vp(t.data_ptr())


def ubench(kind, iters=2000, device=0):
    g = C.c_double()
    r = lib().b2me_ubench(C.c_int(device), C.c_int(kind), C.c_int(iters), C.byref(g))
    if r:
        raise B2Error(f"b2me_ubench failed: {r}")
    return g.value


def ubench_i8(iters=4000, device=0):
    """measured dense tcgen05.mma kind::i8 rate in Tops (b2fp_ubench_i8)"""
    t = C.c_double()
    r = lib().b2fp_ubench_i8(C.c_int(device), C.c_int(iters), C.byref(t))
    if r:
        raise B2Error(f"b2fp_ubench_i8 failed: {r}")
    return t.value


class Searcher:
    """One b2me context: a coded picture size, nrefs reference pictures, SearchRange R."""

    def __init__(self, W, H, nrefs, R, device=0):
        self.L = lib()
        self.W, self.H, self.nrefs, self.R, self.device = W, H, nrefs, R, device
        self.nmb = (W // 16) * (H // 16)
        h = _vp()
        r = self.L.b2me_create(C.byref(h), C.c_int(device), C.c_int(W), C.c_int(H), C.c_int(nrefs), C.c_int(R))
        self.h = h
        if r:
            msg = self.L.b2me_last_error(h if h else _vp(0))
            raise B2Error(f"b2me_create failed ({r}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.L.b2me_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, r, what):
        if r:
            msg = self.L.b2me_last_error(self.h)
            raise B2Error(f"{what} failed ({r}): {msg.decode() if msg else ''}")

    # ---- pictures -----------------------------------------------------------------------
    def set_cur(self, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        assert luma.shape == (self.H, self.W)
        self._chk(self.L.b2me_set_cur(self.h, _p(luma), C.c_int(self.W)), "b2me_set_cur")

    def set_ref(self, r, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        assert luma.shape == (self.H, self.W)
        self._chk(self.L.b2me_set_ref(self.h, C.c_int(r), _p(luma), C.c_int(self.W)), "b2me_set_ref")

    def set_ref_weights(self, r, weight, offset, log_denom, apply=True):
        """explicit weighted prediction of reference slot r; takes effect at the next set_ref of that slot"""
        self._chk(self.L.b2me_set_ref_weights(self.h, C.c_int(r), C.c_int(int(apply)), C.c_int(weight), C.c_int(offset), C.c_int(log_denom)),
                  "b2me_set_ref_weights")

    def set_cur_dev(self, t, stream=0):
        self._chk(self.L.b2me_set_cur_dev(self.h, _dp(t), C.c_int(t.stride(0)), _vp(stream)), "b2me_set_cur_dev")

    def set_ref_dev(self, r, t, stream=0):
        self._chk(self.L.b2me_set_ref_dev(self.h, C.c_int(r), _dp(t), C.c_int(t.stride(0)), _vp(stream)), "b2me_set_ref_dev")

    def check_errors(self, stream=0):
        """b2me_check_errors: synchronises `stream`; raises if a _dev search since the last check saw a quarter-pel centre"""
        self._chk(self.L.b2me_check_errors(self.h, _vp(stream)), "b2me_check_errors")

    def set_ref_rows_dev(self, r, t, row_first, row_count, stream=0):
        """planes of luma rows [row_first, row_first + row_count) only (MB-row bands), from a full-geometry CUDA picture"""
        self._chk(self.L.b2me_set_ref_rows_dev(self.h, C.c_int(r), _dp(t), C.c_int(t.stride(0)), C.c_int(row_first), C.c_int(row_count),
                                               _vp(stream)), "b2me_set_ref_rows_dev")

    def subplane(self, r, yy, xx):
        out = np.zeros((self.H + 2 * PAD_Y, self.W + 2 * PAD_X), np.uint8)
        self._chk(self.L.b2me_get_subplane(self.h, C.c_int(r), C.c_int(yy), C.c_int(xx), _p(out)), "b2me_get_subplane")
        return out

    # ---- search -------------------------------------------------------------------------
    def search_frame(self, pred, center, params, want_int=True):
        """Host numpy in/out through the C ABI (copies inside the call).  want_int=False (sub-pel searches only):
        the integer stage's results stay on the device, (None, None, mv_sub, cost_sub) comes back."""
        pred = np.ascontiguousarray(pred, np.int16)
        center = np.ascontiguousarray(center, np.int16)
        shape = (self.nmb, self.nrefs, NPART, 2)
        assert pred.shape == shape and center.shape == shape
        mv_int = np.zeros(shape, np.int16)
        mv_sub = np.zeros(shape, np.int16)
        cost_int = np.zeros(shape[:3], np.int64)
        cost_sub = np.zeros(shape[:3], np.int64)
        if not want_int:
            self._chk(self.L.b2me_search_frame(self.h, _p(pred), _p(center), C.byref(params), None, None,
                                               _p(mv_sub), _p(cost_sub)), "b2me_search_frame")
            return None, None, mv_sub, cost_sub
        self._chk(self.L.b2me_search_frame(self.h, _p(pred), _p(center), C.byref(params), _p(mv_int), _p(cost_int),
                                           _p(mv_sub), _p(cost_sub)), "b2me_search_frame")
        return mv_int, cost_int, mv_sub, cost_sub

    def search_frame_best(self, pred_mb, params, ref_lambda):
        """b2me_search_frame_best: pred_mb [nmb][nrefs][2] -> best_ref [nmb][21] i8, best_cost [nmb][21] i32, best_mv [nmb][41][2] i16"""
        pred_mb = np.ascontiguousarray(pred_mb, np.int16)
        assert pred_mb.shape == (self.nmb, self.nrefs, 2)
        br = np.zeros((self.nmb, 21), np.int8); bc = np.zeros((self.nmb, 21), np.int32); bm = np.zeros((self.nmb, NPART, 2), np.int16)
        self._chk(self.L.b2me_search_frame_best(self.h, _p(pred_mb), C.byref(params), C.c_int(int(ref_lambda)), _p(br), _p(bc), _p(bm)),
                  "b2me_search_frame_best")
        return br, bc, bm

    def search_frame_dev(self, pred, center, params, mv_int, cost_int, mv_sub, cost_sub, stream=0,
                         mb_first=0, mb_count=None):
        """torch CUDA tensors in/out (device pointers), asynchronous on `stream`."""
        if mb_count is None:
            mb_count = self.nmb - mb_first
        self._chk(self.L.b2me_search_mbs_dev(self.h, C.c_int(mb_first), C.c_int(mb_count), _dp(pred), _dp(center),
                                             C.byref(params), _dp(mv_int), _dp(cost_int), _dp(mv_sub), _dp(cost_sub),
                                             _vp(stream)), "b2me_search_mbs_dev")

    def mc_luma_dev(self, mb_mode, b8mode, ref8, mv, orig_blk, pred_blk, stream=0):
        """torch CUDA tensors: luma_prediction (list 0) of the picture into the b2tq block layout"""
        self._chk(self.L.b2me_mc_luma_dev(self.h, _dp(mb_mode), _dp(b8mode), _dp(ref8), _dp(mv), _dp(orig_blk), _dp(pred_blk), _vp(stream)),
                  "b2me_mc_luma_dev")

    def select_refs_dev(self, cost, ref_lambda, best_ref, best_cost, stream=0):
        """list_prediction_cost (list 0) on torch CUDA tensors: cost [nmb][nrefs][41] i64 -> best_ref [nmb][21] i8, best_cost [nmb][21] i64"""
        self._chk(self.L.b2me_select_refs_dev(self.h, _dp(cost), C.c_int(int(ref_lambda)), _dp(best_ref), _dp(best_cost), _vp(stream)),
                  "b2me_select_refs_dev")

    def set_cur_chroma(self, u, v):
        u = np.ascontiguousarray(u, np.uint8); v = np.ascontiguousarray(v, np.uint8)
        self._chk(self.L.b2me_set_cur_chroma(self.h, _p(u), _p(v), C.c_int(u.shape[1])), "b2me_set_cur_chroma")

    def set_ref_chroma(self, ref, u, v):
        u = np.ascontiguousarray(u, np.uint8); v = np.ascontiguousarray(v, np.uint8)
        self._chk(self.L.b2me_set_ref_chroma(self.h, C.c_int(ref), _p(u), _p(v), C.c_int(u.shape[1])), "b2me_set_ref_chroma")

    def mc_mb_dev(self, mb_mode, b8mode, pdir, ref8, mv_l0, mv_l1, orig_y, pred_y, orig_c, pred_c, stream=0):
        """luma + chroma prediction from either list or both on torch CUDA tensors (see include/b2me.h b2me_mc_mb_dev)"""
        self._chk(self.L.b2me_mc_mb_dev(self.h, _dp(mb_mode), _dp(b8mode), _dp(pdir), _dp(ref8), _dp(mv_l0), _dp(mv_l1), _dp(orig_y), _dp(pred_y),
                                        _dp(orig_c), _dp(pred_c), _vp(stream)), "b2me_mc_mb_dev")

    def select_refs_list_dev(self, cost, list_size, ref_lambda, best_ref, best_cost, stream=0):
        """list_prediction_cost for one list of a B slice: only the first list_size references of `cost` take part"""
        self._chk(self.L.b2me_select_refs_list_dev(self.h, _dp(cost), C.c_int(int(list_size)), C.c_int(int(ref_lambda)), _dp(best_ref), _dp(best_cost), _vp(stream)),
                  "b2me_select_refs_list_dev")

    def bipred_search(self, jobs, params, apply_weights=False, log_denom=0, test8x8=False):
        """full_search_bipred (+ sub_pel_bipred) for an array of synth.BIPRED_JOB records; returns BIPRED_RESULT records"""
        from . import synth
        jobs = np.ascontiguousarray(jobs, synth.BIPRED_JOB)
        out = np.zeros(len(jobs), synth.BIPRED_RESULT)
        self._chk(self.L.b2me_bipred_search(self.h, C.c_int(len(jobs)), _p(jobs), C.byref(params), C.c_int(int(apply_weights)),
                                            C.c_int(log_denom), C.c_int(int(test8x8)), _p(out)), "b2me_bipred_search")
        return out

    def distortion_candidates(self, cands, metric, test8x8=False):
        """computeSAD / SSE / SATD (<< 5) of synth.CANDIDATE records; int64 array"""
        from . import synth
        cands = np.ascontiguousarray(cands, synth.CANDIDATE)
        out = np.zeros(len(cands), np.int64)
        self._chk(self.L.b2me_distortion_candidates(self.h, C.c_int(metric), C.c_int(int(test8x8)), C.c_int(len(cands)), _p(cands), _p(out)),
                  "b2me_distortion_candidates")
        return out

    def bid_partition_cost(self, jobs, metric, transform8x8=False, apply_weights=False, log_denom=0):
        """b2me_bid_partition_cost: BIDPartitionCost of synth.BID_JOB records; int64 array"""
        from . import synth
        jobs = np.ascontiguousarray(jobs, synth.BID_JOB)
        out = np.zeros(len(jobs), np.int64)
        self._chk(self.L.b2me_bid_partition_cost(self.h, C.c_int(metric), C.c_int(int(transform8x8)), C.c_int(int(apply_weights)), C.c_int(log_denom),
                                                 C.c_int(len(jobs)), _p(jobs), _p(out)), "b2me_bid_partition_cost")
        return out

    def epzs_search(self, jobs, preds, patterns):
        """b2me_epzs_search: synth.EPZS_JOB records, preds [n][2] int16, synth.EPZS_PATTERN records -> synth.EPZS_RESULT records"""
        from . import synth
        jobs = np.ascontiguousarray(jobs, synth.EPZS_JOB); patterns = np.ascontiguousarray(patterns, synth.EPZS_PATTERN)
        preds = np.ascontiguousarray(preds, np.int16).reshape(-1, 2)
        out = np.zeros(len(jobs), synth.EPZS_RESULT)
        self._chk(self.L.b2me_epzs_search(self.h, C.c_int(len(jobs)), _p(jobs), C.c_int(len(preds)), _p(preds), C.c_int(len(patterns)), _p(patterns), _p(out)),
                  "b2me_epzs_search")
        return out

    def epzs_search_dev(self, njobs, jobs_t, preds_t, npatterns, patterns_t, out_t, stream=0):
        """device buffers (torch uint8 / int16 tensors holding the C records); asynchronous on `stream`"""
        self._chk(self.L.b2me_epzs_search_dev(self.h, C.c_int(njobs), C.c_void_p(jobs_t.data_ptr()), C.c_void_p(preds_t.data_ptr()), C.c_int(npatterns),
                                              C.c_void_p(patterns_t.data_ptr()), C.c_void_p(out_t.data_ptr()), C.c_void_p(stream)), "b2me_epzs_search_dev")

    def block_search(self, pos_x, pos_y, blocktype, ref, pred_mv, center_mv, params, search_range):
        pm = (C.c_int16 * 2)(int(pred_mv[0]), int(pred_mv[1]))
        cm = (C.c_int16 * 2)(int(center_mv[0]), int(center_mv[1]))
        mi, ms = (C.c_int16 * 2)(), (C.c_int16 * 2)()
        ci, cs = C.c_int64(), C.c_int64()
        self._chk(self.L.b2me_block_search(self.h, pos_x, pos_y, blocktype, ref, pm, cm, C.byref(params),
                                           C.c_int(search_range), mi, C.byref(ci), ms, C.byref(cs)), "b2me_block_search")
        return (mi[0], mi[1]), ci.value, (ms[0], ms[1]), cs.value

    # ---- instrumentation ------------------------------------------------------------------
    def launch_count(self):
        return self.L.b2me_launch_count(self.h)

    def search_stats(self, reset=True):
        out = (C.c_int64 * 3)()
        self._chk(self.L.b2me_search_stats(self.h, out, C.c_int(int(reset))), "b2me_search_stats")
        return {"exact_evals": out[0], "window_passes": out[1], "items": out[2]}

    def kernel_timing(self, enable):
        self._chk(self.L.b2me_kernel_timing(self.h, C.c_int(int(enable))), "b2me_kernel_timing")

    def kernel_time_ms(self, which):
        ms, n = C.c_double(), C.c_int64()
        self._chk(self.L.b2me_kernel_time_ms(self.h, C.c_int(which), C.byref(ms), C.byref(n)), "b2me_kernel_time_ms")
        return ms.value, n.value


FR_NODE = np.dtype([("block_type", np.int32), ("partition", np.int32), ("reference", np.int32), ("x", np.int32), ("y", np.int32),
                    ("reserved", np.int32), ("scale", np.float64), ("offset", np.float64)])     # b2fr_node


class FractalSearcher:
    """One b2fr context (include/b2me.h): version1's fractal range/domain block search for one
    picture size.  Mirrors the reference's call sequence: set_range (compute_range_Sum),
    set_domain (compute_domain_Sum), then full_search per block or search_plane for the grid."""

    def __init__(self, W, H, R, device=0):
        self.L = lib()
        self.L.b2fr_last_error.restype = C.c_char_p
        self.L.b2fr_last_error.argtypes = [_vp]
        self.L.b2fr_destroy.argtypes = [_vp]
        self.L.b2fr_destroy.restype = None
        self.L.b2fr_launch_count.restype = C.c_int64
        self.L.b2fr_launch_count.argtypes = [_vp]
        self.W, self.H, self.R = W, H, R
        h = _vp()
        r = self.L.b2fr_create(C.byref(h), C.c_int(device), C.c_int(W), C.c_int(H), C.c_int(R))
        self.h = h
        if r:
            msg = self.L.b2fr_last_error(h if h else _vp(0))
            raise B2Error(f"b2fr_create failed ({r}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.L.b2fr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, r, what):
        if r:
            msg = self.L.b2fr_last_error(self.h)
            raise B2Error(f"{what} failed ({r}): {msg.decode() if msg else ''}")

    def _planes(self, y, u, v):
        y = np.ascontiguousarray(y, np.uint8)
        assert y.shape == (self.H, self.W)
        u = None if u is None else np.ascontiguousarray(u, np.uint8)
        v = None if v is None else np.ascontiguousarray(v, np.uint8)
        return y, u, v

    def set_range(self, y, u=None, v=None):
        y, u, v = self._planes(y, u, v)
        self._chk(self.L.b2fr_set_range(self.h, _p(y), _p(u) if u is not None else _vp(0), _p(v) if v is not None else _vp(0)), "b2fr_set_range")

    def set_domain(self, plane_set, y, u=None, v=None, build_sums=True):
        y, u, v = self._planes(y, u, v)
        self._chk(self.L.b2fr_set_domain(self.h, C.c_int(plane_set), _p(y), _p(u) if u is not None else _vp(0),
                                         _p(v) if v is not None else _vp(0), C.c_int(int(build_sums))), "b2fr_set_domain")

    def grid(self, con):
        return (self.W // 16, self.H // 16) if con == 1 else (self.W // 16 // 2, self.H // 16 // 2)

    def search_plane(self, plane_set, con):
        mbw, mbh = self.grid(con)
        xy = np.zeros((mbw * mbh, NPART, 2), np.int32)
        so = np.zeros((mbw * mbh, NPART, 2), np.float64)
        rms = np.zeros((mbw * mbh, NPART), np.float64)
        self._chk(self.L.b2fr_search_plane(self.h, C.c_int(plane_set), C.c_int(con), _p(xy), _p(so), _p(rms)), "b2fr_search_plane")
        return xy, so, rms

    def full_search(self, plane_set, bx, by, bsx, bsy, con, xy=(0, 0)):
        v = (C.c_int32 * 2)(int(xy[0]), int(xy[1]))
        so = (C.c_double * 2)()
        rms = C.c_double()
        self._chk(self.L.b2fr_full_search(self.h, C.c_int(plane_set), bx, by, bsx, bsy, con, v, so, C.byref(rms)), "b2fr_full_search")
        return (v[0], v[1]), (so[0], so[1]), rms.value

    def encode_plane(self, con, tol):
        """F5: encode_one_macroblock's TRANS_NODE tree of every macroblock of component con; record array [nmb][21]"""
        mbw, mbh = self.grid(con)
        nodes = np.zeros((mbw * mbh, 21), FR_NODE)
        t = (C.c_double * 3)(*[float(x) for x in tol])
        self._chk(self.L.b2fr_encode_plane(self.h, C.c_int(con), t, _p(nodes)), "b2fr_encode_plane")
        return nodes

    def decode_plane(self, con, nodes=None):
        """F8: decode_one_macroblock of every macroblock (nodes=None: the trees encode_plane left on the device)"""
        w, h = (self.W, self.H) if con == 1 else (self.W // 2, self.H // 2)
        rec = np.zeros((h, w), np.uint8)
        if nodes is not None:
            nodes = np.ascontiguousarray(nodes, FR_NODE)
        self._chk(self.L.b2fr_decode_plane(self.h, C.c_int(con), None if nodes is None else _p(nodes), _p(rec)), "b2fr_decode_plane")
        return rec

    def domain_table(self, plane_set, con, bw, bh, squares):
        w, h = (self.W, self.H) if con == 1 else (self.W // 2, self.H // 2)
        out = np.zeros((h, w), np.int32)
        self._chk(self.L.b2fr_get_domain_table(self.h, C.c_int(plane_set), C.c_int(con), C.c_int(bw), C.c_int(bh), C.c_int(int(squares)), _p(out)), "b2fr_get_domain_table")
        return out

    def range_table(self, con, squares):
        w, h = (self.W, self.H) if con == 1 else (self.W // 2, self.H // 2)
        out = np.zeros((h // 4, w // 4), np.int32)
        self._chk(self.L.b2fr_get_range_table(self.h, C.c_int(con), C.c_int(int(squares)), _p(out)), "b2fr_get_range_table")
        return out

    def launch_count(self):
        return self.L.b2fr_launch_count(self.h)


class PoolSearcher:
    """One b2fp context (include/b2me.h): 8x8 range blocks x 8 isometries against a pool of 2:1-averaged
    16x16 domain blocks, cross terms on the tensor cores (tcgen05.mma kind::i8)."""

    def __init__(self, range_w, range_h, domain_w, domain_h, pool_size, device=0):
        self.L = lib()
        self.L.b2fp_last_error.restype = C.c_char_p
        self.L.b2fp_last_error.argtypes = [_vp]
        self.L.b2fp_destroy.argtypes = [_vp]
        self.L.b2fp_destroy.restype = None
        self.L.b2fp_launch_count.restype = C.c_int64
        self.L.b2fp_launch_count.argtypes = [_vp]
        self.rw, self.rh, self.dw, self.dh, self.nd = range_w, range_h, domain_w, domain_h, pool_size
        self.nr = (range_w // 8) * (range_h // 8)
        h = _vp()
        r = self.L.b2fp_create(C.byref(h), C.c_int(device), C.c_int(range_w), C.c_int(range_h), C.c_int(domain_w), C.c_int(domain_h), C.c_int(pool_size))
        self.h = h
        if r:
            msg = self.L.b2fp_last_error(h if h else _vp(0))
            raise B2Error(f"b2fp_create failed ({r}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.L.b2fp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, r, what):
        if r:
            msg = self.L.b2fp_last_error(self.h)
            raise B2Error(f"{what} failed ({r}): {msg.decode() if msg else ''}")

    def positions(self):
        xy = np.zeros((self.nd, 2), np.int32)
        self._chk(self.L.b2fp_pool_positions(self.h, _p(xy)), "b2fp_pool_positions")
        return xy

    def set_planes(self, range_plane, domain_plane):
        r = np.ascontiguousarray(range_plane, np.uint8); d = np.ascontiguousarray(domain_plane, np.uint8)
        assert r.shape == (self.rh, self.rw) and d.shape == (self.dh, self.dw)
        self._chk(self.L.b2fp_set_planes(self.h, _p(r), C.c_int(self.rw), _p(d), C.c_int(self.dw)), "b2fp_set_planes")

    def set_planes_dev(self, range_ptr, range_stride, domain_ptr, domain_stride, stream):
        self._chk(self.L.b2fp_set_planes_dev(self.h, _vp(range_ptr), C.c_int(range_stride), _vp(domain_ptr), C.c_int(domain_stride), _vp(stream)), "b2fp_set_planes_dev")

    def search(self):
        dom = np.zeros(self.nr, np.int32); iso = np.zeros(self.nr, np.uint8)
        aq = np.zeros(self.nr, np.int16); beta = np.zeros(self.nr, np.int16); err = np.zeros(self.nr, np.int64)
        self._chk(self.L.b2fp_search(self.h, _p(dom), _p(iso), _p(aq), _p(beta), _p(err)), "b2fp_search")
        return dom, iso, aq, beta, err

    def search_dev(self, dom, iso, aq, beta, err, stream):
        self._chk(self.L.b2fp_search_dev(self.h, _vp(dom), _vp(iso), _vp(aq), _vp(beta), _vp(err), _vp(stream)), "b2fp_search_dev")

    def kernel_time_ms(self, reset=True):
        ms, n = C.c_double(), C.c_int64()
        self._chk(self.L.b2fp_kernel_time_ms(self.h, C.byref(ms), C.byref(n), C.c_int(int(reset))), "b2fp_kernel_time_ms")
        return ms.value, n.value

    def probe_ms(self):
        ms = C.c_double()
        self._chk(self.L.b2fp_probe(self.h, C.byref(ms)), "b2fp_probe")
        return ms.value

    def stats(self, reset=True):
        out = (C.c_int64 * 3)()
        self._chk(self.L.b2fp_stats(self.h, out, C.c_int(int(reset))), "b2fp_stats")
        return {"exact_evals": out[0], "chunk_rescans": out[1], "chunks": out[2]}

    def launch_count(self):
        return self.L.b2fp_launch_count(self.h)


def tq_chroma(params, orig, pred, device=0):
    """b2tq_chroma (4:2:0 chroma, one plane): orig / pred [nmb][64] raster -> dc_level, dc_run, ac_level, ac_run, recon, cr_cbp"""
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nmb = orig.shape[0]
    dl = np.zeros((nmb, 4), np.int16); dr = np.zeros((nmb, 4), np.uint8)
    al = np.zeros((nmb, 4, 16), np.int16); ar = np.zeros((nmb, 4, 16), np.uint8)
    rec = np.zeros((nmb, 64), np.uint8); cbp = np.zeros(nmb, np.uint8)
    L = lib()
    L.b2tq_last_error.restype = C.c_char_p
    r = L.b2tq_chroma(C.c_int(device), C.byref(params), C.c_int(nmb), _p(orig), _p(pred), _p(dl), _p(dr), _p(al), _p(ar), _p(rec), _p(cbp))
    if r:
        raise B2Error(f"b2tq_chroma failed ({r}): {L.b2tq_last_error().decode()}")
    return dl, dr, al, ar, rec, cbp


def tq16x16(params, orig, pred, device=0):
    """b2tq_16x16 (Intra16x16 luma): orig / pred [nmb][256] raster -> dc_level, dc_run, ac_level, ac_run, recon, ac_coef"""
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nmb = orig.shape[0]
    dl = np.zeros((nmb, 16), np.int16); dr = np.zeros((nmb, 16), np.uint8)
    al = np.zeros((nmb, 16, 16), np.int16); ar = np.zeros((nmb, 16, 16), np.uint8)
    rec = np.zeros((nmb, 256), np.uint8); ac = np.zeros(nmb, np.uint8)
    L = lib()
    L.b2tq_last_error.restype = C.c_char_p
    r = L.b2tq_16x16(C.c_int(device), C.byref(params), C.c_int(nmb), _p(orig), _p(pred), _p(dl), _p(dr), _p(al), _p(ar), _p(rec), _p(ac))
    if r:
        raise B2Error(f"b2tq_16x16 failed ({r}): {L.b2tq_last_error().decode()}")
    return dl, dr, al, ar, rec, ac


def tq_dev(params, orig, pred, n, level, run, recon, cost, nonzero, stream=0):
    """b2tq_4x4_dev / b2tq_8x8_dev on torch CUDA tensors (orig, pred [nblk][n*n] u8; outputs preallocated)."""
    L = lib()
    L.b2tq_last_error.restype = C.c_char_p
    f = L.b2tq_4x4_dev if n == 4 else L.b2tq_8x8_dev
    r = f(C.byref(params), C.c_int(orig.shape[0]), _dp(orig), _dp(pred), _dp(level), _dp(run), _dp(recon), _dp(cost), _dp(nonzero), _vp(stream))
    if r:
        raise B2Error(f"b2tq_{n}x{n}_dev failed ({r}): {L.b2tq_last_error().decode()}")


def distortion_blocks(kind, n, diff, device=0):
    """b2me_distortion_blocks: diff [nblk][n*n] int16 -> int64 [nblk] (distortion << 5); kind 0 SAD, 1 SSE, 2 SATD."""
    d = np.ascontiguousarray(diff, np.int16)
    assert d.ndim == 2 and d.shape[1] == n * n
    out = np.zeros(d.shape[0], np.int64)
    r = lib().b2me_distortion_blocks(C.c_int(device), C.c_int(kind), C.c_int(n), C.c_int(d.shape[0]), _p(d), _p(out))
    if r:
        raise B2Error(f"b2me_distortion_blocks failed ({r})")
    return out


class TQParams(C.Structure):
    """b2tq_params (include/b2me.h)"""
    _fields_ = [("qp", C.c_int32), ("mode", C.c_int32), ("cavlc", C.c_int32), ("field_scan", C.c_int32),
                ("disthres", C.c_int32), ("reserved", C.c_int32 * 3), ("scale", C.c_int32 * 64),
                ("offset", C.c_int32 * 64), ("invscale", C.c_int32 * 64)]


def tq_default_params(n, qp, intra, mode=0, cavlc=1, field_scan=0, disthres=0):
    p = TQParams()
    r = lib().b2tq_default_params(C.byref(p), C.c_int(int(n == 8)), C.c_int(qp), C.c_int(intra), C.c_int(mode))
    if r:
        raise B2Error(f"b2tq_default_params failed ({r})")
    p.cavlc, p.field_scan, p.disthres = cavlc, field_scan, disthres
    return p


def tq_params_table(p, n):
    return np.array([list(p.scale)[:n * n], list(p.offset)[:n * n], list(p.invscale)[:n * n]], np.int32)


def tq(params, orig, pred, n, device=0):
    """Host arrays through b2tq_4x4 / b2tq_8x8: returns level, run, recon, coeff_cost, nonzero."""
    L = lib()
    L.b2tq_last_error.restype = C.c_char_p
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nblk, m = orig.shape[0], n * n
    assert orig.shape == (nblk, m) and pred.shape == (nblk, m)
    level = np.zeros((nblk, m), np.int16); run = np.zeros((nblk, m), np.uint8)
    recon = np.zeros((nblk, m), np.uint8); cost = np.zeros(nblk, np.int32); nz = np.zeros(nblk, np.uint8)
    f = L.b2tq_4x4 if n == 4 else L.b2tq_8x8
    r = f(C.c_int(device), C.byref(params), C.c_int(nblk), _p(orig), _p(pred), _p(level), _p(run), _p(recon), _p(cost), _p(nz))
    if r:
        raise B2Error(f"b2tq_{n}x{n} failed ({r}): {L.b2tq_last_error().decode()}")
    return level, run, recon, cost, nz


def deblock_frame(y, u, v, mbs, blks, device=0):
    """b2dbk_frame: DeblockFrame on the GPU; host planes (packed uint8), synth.DBK_MB / DBK_BLK records; returns filtered copies"""
    from . import synth
    y, u, v = (np.ascontiguousarray(a, np.uint8).copy() for a in (y, u, v))
    H, W = y.shape
    mbs = np.ascontiguousarray(mbs, synth.DBK_MB); blks = np.ascontiguousarray(blks, synth.DBK_BLK)
    if mbs.size != (W // 16) * (H // 16) or blks.size != (W // 4) * (H // 4):
        raise ValueError("deblock_frame: record arrays do not match the picture size")
    L = lib()
    r = L.b2dbk_frame(C.c_int(device), C.c_int(W), C.c_int(H), _p(y), _p(u), _p(v), _p(mbs), _p(blks))
    if r != 0:
        L.b2dbk_last_error.restype = C.c_char_p
        raise RuntimeError(f"b2dbk_frame failed ({r}): {L.b2dbk_last_error().decode()}")
    return y, u, v


def deblock_frame_dev(y, u, v, mbs, blks, progress, stream=0):
    """b2dbk_frame_dev on torch CUDA tensors (uint8 planes in place; mbs / blks uint8 views of the records; progress int32 [H / 16])"""
    L = lib()
    H, W = y.shape
    r = L.b2dbk_frame_dev(C.c_int(W), C.c_int(H), C.c_void_p(y.data_ptr()), C.c_int(y.stride(0)), C.c_void_p(u.data_ptr()), C.c_void_p(v.data_ptr()),
                          C.c_int(u.stride(0)), C.c_void_p(mbs.data_ptr()), C.c_void_p(blks.data_ptr()), C.c_void_p(progress.data_ptr()), C.c_void_p(stream))
    if r != 0:
        L.b2dbk_last_error.restype = C.c_char_p
        raise RuntimeError(f"b2dbk_frame_dev failed ({r}): {L.b2dbk_last_error().decode()}")
