"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for
  * liborc.so          : our CPU restatement of the reference hot path (oracle/b2_oracle*.c)
  * _ref/libjmref.so   : the UNMODIFIED JM 18.5 objects + harness (oracle/jm_harness.c)
  * _ref/libv1ref.so   : the UNMODIFIED version1 compute.c/block_enc.c + harness

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product (h264_b200) never does.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DISTBLK_MAX = (2**31 - 1) << 5
PAD_X, PAD_Y = 32, 20

_vp = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return C.CDLL(path)


def have_jmref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libjmref.so"))


def have_v1ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libv1ref.so"))


_orc = None


def orc_lib():
    global _orc
    if _orc is None:
        L = _load(os.path.join(_HERE, "liborc.so"))
        L.orc_frame_create.restype = _vp
        L.orc_frame_planes.restype = _vp
        L.orc_full_search.restype = C.c_int64
        L.orc_sub_pel.restype = C.c_int64
        L.orc_call_full_search.restype = C.c_int64
        L.orc_call_sub_pel.restype = C.c_int64
        _orc = L
    return _orc


def spiral(R):
    """JM spiral order, (2R+1)^2 x 2 int16 integer-pel offsets (mv_search.c:406-442)."""
    n = max(9, (2 * R + 1) ** 2)
    out = np.zeros((n, 2), np.int16)
    orc_lib().orc_spiral(C.c_int(R), _ptr(out))
    return out


def mvbits(d):
    return orc_lib().orc_mvbits(C.c_int(int(d)))


def spiral_xy(pos):
    """integer displacement of position pos of the JM spiral (mv_search.c:406-442)"""
    if pos == 0:
        return 0, 0
    l = 1
    while (2 * l + 1) ** 2 <= pos:
        l += 1
    k = pos - (2 * l - 1) ** 2
    if k < 2 * (2 * l - 1):
        return (k >> 1) - l + 1, (l if k & 1 else -l)
    k -= 2 * (2 * l - 1)
    return (l if k & 1 else -l), (k >> 1) - l


def partition_geometry():
    """[(blocktype, ox, oy, w, h)] for the 41 partitions of an MB."""
    L = orc_lib()
    out = []
    for p in range(41):
        v = [C.c_int() for _ in range(5)]
        L.orc_partition_geometry(C.c_int(p), *[C.byref(x) for x in v])
        out.append(tuple(x.value for x in v))
    return out


SELECT_ENTRIES = [(1, 0), (2, 0), (2, 1), (3, 0), (3, 1)] + [(m, b) for m in (4, 5, 6, 7) for b in range(4)]   # (mode, block) of the 21 entries


# (mode, block) entries of list_prediction_cost -> the partitions whose motion costs they sum (include/b2me.h, b2me_select_refs_dev)
ENTRY_PARTS = [(0,), (1,), (2,), (3,), (4,), (5,), (6,), (7,), (8,), (9, 11), (10, 12), (13, 15), (14, 16),
               (17, 18), (19, 20), (21, 22), (23, 24), (25, 26, 29, 30), (27, 28, 31, 32), (33, 34, 37, 38), (35, 36, 39, 40)]


def select_refs(cost, ref_lambda, list_size=None):
    """list_prediction_cost (list < BI_PRED) for every macroblock: cost [nmb][nrefs][41] -> (best_ref int8 [nmb][21], best_cost int64 [nmb][21]);
    list_size: listXsize of the list the costs belong to (default: all nrefs)"""
    cost = np.ascontiguousarray(cost, np.int64)
    nmb, nrefs = cost.shape[:2]
    br = np.zeros((nmb, 21), np.int8); bc = np.zeros((nmb, 21), np.int64)
    orc_lib().orc_select_refs_list(C.c_int(nmb), C.c_int(nrefs), C.c_int(nrefs if list_size is None else int(list_size)), _ptr(cost), C.c_int(int(ref_lambda)), _ptr(br), _ptr(bc))
    return br, bc


def deblock_frame(y, u, v, mbs, blks):
    """orc_deblock_frame (DeblockFrame of JM restated): returns filtered copies of the planes"""
    from h264_b200 import synth
    y, u, v = (np.ascontiguousarray(a, np.uint8).copy() for a in (y, u, v))
    H, W = y.shape
    mbs = np.ascontiguousarray(mbs, synth.DBK_MB); blks = np.ascontiguousarray(blks, synth.DBK_BLK)
    assert mbs.size == (W // 16) * (H // 16) and blks.size == (W // 4) * (H // 4)
    orc_lib().orc_deblock_frame(C.c_int(W), C.c_int(H), _ptr(y), _ptr(u), _ptr(v), _ptr(mbs), _ptr(blks))
    return y, u, v


def subpel_planes(luma):
    H, W = luma.shape
    out = np.zeros((4, 4, H + 2 * PAD_Y, W + 2 * PAD_X), np.uint8)
    luma = np.ascontiguousarray(luma, np.uint8)
    orc_lib().orc_subpel_planes(_ptr(luma), C.c_int(W), C.c_int(H), _ptr(out))
    return out


def hadamard4x4(diff):
    d = np.ascontiguousarray(diff, np.int16)
    return orc_lib().orc_hadamard4x4(_ptr(d))


def distortion_blocks(kind, n, diff):
    """distortion4x4/8x8{SAD,SSE,SATD} restated (me_distortion.c:38-134): diff [nblk][n*n] int16 -> int64 [nblk]."""
    L = orc_lib()
    L.orc_distortion.restype = C.c_int64
    d = np.ascontiguousarray(diff, np.int16)
    return np.array([L.orc_distortion(C.c_int(kind), C.c_int(n), _ptr(d[i])) for i in range(d.shape[0])], np.int64)


def hadamard8x8(diff):
    d = np.ascontiguousarray(diff, np.int16)
    return orc_lib().orc_hadamard8x8(_ptr(d))


# b2me_bipred_job / b2me_bipred_result (include/b2me.h)
BIPRED_JOB = np.dtype([("min_mcost", np.int64), ("pos_x", np.int16), ("pos_y", np.int16), ("blocktype", np.int16),
                       ("ref1", np.int16), ("ref2", np.int16), ("search_range", np.int16),
                       ("pred1", np.int16, 2), ("pred2", np.int16, 2), ("mv1", np.int16, 2), ("mv2", np.int16, 2),
                       ("weight1", np.int16), ("weight2", np.int16), ("offset_bi", np.int16), ("reserved", np.int16)], align=True)
BIPRED_RESULT = np.dtype([("cost_int", np.int64), ("cost_sub", np.int64), ("mv_int", np.int16, 2), ("mv_sub", np.int16, 2)], align=True)
assert BIPRED_JOB.itemsize == 48 and BIPRED_RESULT.itemsize == 24


class OrcFrame:
    """Restated oracle for one (current frame, reference frames) pair."""

    def __init__(self, cur, refs, R):
        self.L = orc_lib()
        cur = np.ascontiguousarray(cur, np.uint8)
        refs = np.ascontiguousarray(refs, np.uint8)
        self.H, self.W = cur.shape
        self.nrefs = refs.shape[0]
        self.R = R
        self.h = _vp(self.L.orc_frame_create(C.c_int(self.W), C.c_int(self.H), C.c_int(self.nrefs),
                                             C.c_int(R), _ptr(cur), _ptr(refs)))

    def __del__(self):
        try:
            self.L.orc_frame_destroy(self.h)
        except Exception:
            pass

    def set_weights(self, ref, weight, offset, log_denom):
        """explicit WP of reference `ref` (computeSADWP/SATDWP/SSEWP); call once, after construction"""
        self.L.orc_frame_set_weights(self.h, C.c_int(ref), C.c_int(weight), C.c_int(offset), C.c_int(log_denom))

    def mc_luma(self, mb_mode, b8mode, ref8, mv):
        """luma_prediction (list 0) of the whole picture from a search result array: (orig_blk, pred_blk) [nmb*16][16]"""
        nmb = (self.W // 16) * (self.H // 16)
        mb_mode = np.ascontiguousarray(mb_mode, np.uint8); b8mode = np.ascontiguousarray(b8mode, np.uint8)
        ref8 = np.ascontiguousarray(ref8, np.int8); mv = np.ascontiguousarray(mv, np.int16)
        assert mb_mode.shape == (nmb,) and b8mode.shape == (nmb, 4) and ref8.shape == (nmb, 4) and mv.shape == (nmb, self.nrefs, 41, 2)
        orig = np.zeros((nmb * 16, 16), np.uint8); pred = np.zeros((nmb * 16, 16), np.uint8)
        self.L.orc_mc_luma(self.h, _ptr(mb_mode), _ptr(b8mode), _ptr(ref8), _ptr(mv), _ptr(orig), _ptr(pred))
        return orig, pred

    def mc_mb(self, mb_mode, b8mode, pdir, ref8, mv0, mv1, curc, refc):
        """orc_mc_mb: luma + chroma prediction from either list or both; returns (orig_y, pred_y, orig_c, pred_c)"""
        nmb = (self.W // 16) * (self.H // 16)
        mb_mode = np.ascontiguousarray(mb_mode, np.uint8); b8mode = np.ascontiguousarray(b8mode, np.uint8); pdir = np.ascontiguousarray(pdir, np.uint8)
        ref8 = np.ascontiguousarray(ref8, np.int8); mv0 = np.ascontiguousarray(mv0, np.int16); mv1 = np.ascontiguousarray(mv1, np.int16)
        curc = np.ascontiguousarray(curc, np.uint8); refc = np.ascontiguousarray(refc, np.uint8)
        oy = np.zeros((nmb * 16, 16), np.uint8); py = np.zeros_like(oy); oc = np.zeros((nmb, 2, 4, 16), np.uint8); pc = np.zeros_like(oc)
        self.L.orc_mc_mb(self.h, _ptr(mb_mode), _ptr(b8mode), _ptr(pdir), _ptr(ref8), _ptr(mv0), _ptr(mv1), _ptr(curc), _ptr(refc),
                         _ptr(oy), _ptr(py), _ptr(oc), _ptr(pc))
        return oy, py, oc, pc

    def bipred_search(self, jobs, lam, metric_h=2, metric_q=2, do_subpel=True, test8x8=False, wp=False, log_denom=0):
        """full_search_bipred (+ sub_pel_bipred) for an array of BIPRED_JOB; returns BIPRED_RESULT array"""
        jobs = np.ascontiguousarray(jobs, BIPRED_JOB)
        out = np.zeros(len(jobs), BIPRED_RESULT)
        lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lam), (3,)), np.int32)
        self.L.orc_bipred_search(self.h, C.c_int(len(jobs)), _ptr(jobs), _ptr(lam), C.c_int(metric_h), C.c_int(metric_q),
                                 C.c_int(int(do_subpel)), C.c_int(int(test8x8)), C.c_int(int(wp)), C.c_int(log_denom), _ptr(out))   # do_subpel 2: the 81-position stage
        return out

    def distortion_candidates(self, cands, metric, test8x8=False):
        """computeSAD / SSE / SATD (<< 5) at explicit (block, ref, vector) records (h264_b200.synth.CANDIDATE layout)"""
        cands = np.ascontiguousarray(cands)
        assert cands.dtype.itemsize == 12
        out = np.zeros(len(cands), np.int64)
        self.L.orc_distortion_candidates(self.h, C.c_int(metric), C.c_int(int(test8x8)), C.c_int(len(cands)), _ptr(cands), _ptr(out))
        return out

    def bid_partition_cost(self, jobs, metric, transform8x8=False, apply_weights=False, log_denom=0):
        """orc_bid_partition_cost (BIDPartitionCost, mv_search.c:1159-1250) over h264_b200.synth.BID_JOB records"""
        jobs = np.ascontiguousarray(jobs)
        assert jobs.dtype.itemsize == 60
        out = np.zeros(len(jobs), np.int64)
        self.L.orc_bid_partition_cost(self.h, C.c_int(metric), C.c_int(int(transform8x8)), C.c_int(int(apply_weights)), C.c_int(log_denom),
                                      C.c_int(len(jobs)), _ptr(jobs), _ptr(out))
        return out

    def epzs_search(self, jobs, preds, patterns):
        """orc_epzs_search (EPZS_motion_estimation / EPZS_subMB_motion_estimation restated over include/b2me.h's job records)"""
        from h264_b200 import synth
        jobs = np.ascontiguousarray(jobs, synth.EPZS_JOB); patterns = np.ascontiguousarray(patterns, synth.EPZS_PATTERN)
        preds = np.ascontiguousarray(preds, np.int16).reshape(-1, 2)
        out = np.zeros(len(jobs), synth.EPZS_RESULT)
        self.L.orc_epzs_search(self.h, C.c_int(len(jobs)), _ptr(jobs), _ptr(preds), C.c_int(len(patterns)), _ptr(patterns), _ptr(out))
        return out

    def planes(self, r):
        Hp, Wp = self.H + 2 * PAD_Y, self.W + 2 * PAD_X
        p = self.L.orc_frame_planes(self.h, C.c_int(r))
        buf = (C.c_uint8 * (16 * Hp * Wp)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(4, 4, Hp, Wp)

    def call_full_search(self, ref, pos_x, pos_y, blocktype, pred_mv, center_mv, search_range, min_mcost, lam):
        pm = np.array(pred_mv, np.int16)
        mv = np.array(center_mv, np.int16)
        c = self.L.orc_call_full_search(self.h, C.c_int(ref), C.c_int(pos_x), C.c_int(pos_y), C.c_int(blocktype),
                                        _ptr(pm), _ptr(mv), C.c_int(search_range), C.c_int64(int(min_mcost)),
                                        C.c_int(int(lam)))
        return (int(mv[0]), int(mv[1])), c

    def call_sub_pel(self, ref, pos_x, pos_y, blocktype, pred_mv, mv_in, min_mcost, lam_h, lam_q,
                     metric_h=2, metric_q=2):
        pm = np.array(pred_mv, np.int16)
        mv = np.array(mv_in, np.int16)
        c = self.L.orc_call_sub_pel(self.h, C.c_int(ref), C.c_int(pos_x), C.c_int(pos_y), C.c_int(blocktype),
                                    _ptr(pm), _ptr(mv), C.c_int64(int(min_mcost)), C.c_int(int(lam_h)),
                                    C.c_int(int(lam_q)), C.c_int(metric_h), C.c_int(metric_q))
        return (int(mv[0]), int(mv[1])), c

    def search_frame(self, pred, center, lambda_factor, restrict_mode=2, metric_h=2, metric_q=2,
                     do_subpel=True, mb_first=0, mb_count=None):
        nmb = (self.W // 16) * (self.H // 16)
        if mb_count is None:
            mb_count = nmb - mb_first
        pred = np.ascontiguousarray(pred, np.int16)
        center = np.ascontiguousarray(center, np.int16)
        assert pred.shape == (nmb, self.nrefs, 41, 2) and center.shape == pred.shape
        lam = np.ascontiguousarray(lambda_factor, np.int32)
        mv_int = np.zeros_like(pred)
        mv_sub = np.zeros_like(pred)
        cost_int = np.zeros(pred.shape[:3], np.int64)
        cost_sub = np.zeros(pred.shape[:3], np.int64)
        self.L.orc_search_frame(self.h, C.c_int(mb_first), C.c_int(mb_count), _ptr(pred), _ptr(center),
                                _ptr(lam), C.c_int(restrict_mode), C.c_int(metric_h), C.c_int(metric_q),
                                C.c_int(int(do_subpel)), _ptr(mv_int), _ptr(cost_int), _ptr(mv_sub),
                                _ptr(cost_sub))
        return mv_int, cost_int, mv_sub, cost_sub


class JMRef:
    """The unmodified JM 18.5 objects behind oracle/jm_harness.c."""

    def __init__(self, W, H, R, nrefs, metric=(0, 2, 2), rdopt=1, restrict_mode=2):
        self.L = _load(os.path.join(_HERE, "_ref", "libjmref.so"))
        self.L.jmh_create.restype = _vp
        self.L.jmh_sad.restype = C.c_int64
        self.L.jmh_satd.restype = C.c_int64
        self.W, self.H, self.R, self.nrefs = W, H, R, nrefs
        self.h = _vp(self.L.jmh_create(W, H, R, nrefs, metric[0], metric[1], metric[2], rdopt, restrict_mode))

    def set_ref(self, r, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        assert luma.shape == (self.H, self.W)
        self.L.jmh_set_ref(self.h, C.c_int(r), _ptr(luma))

    def set_weights(self, log_denom, weights, offsets, apply=True):
        """UseWeightedReferenceME with explicit weights: per-reference luma weight / offset"""
        w = np.ascontiguousarray(weights, np.int16); o = np.ascontiguousarray(offsets, np.int16)
        self.L.jmh_set_weights(self.h, C.c_int(int(apply)), C.c_int(log_denom), _ptr(w), _ptr(o))

    def set_cur(self, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        assert luma.shape == (self.H, self.W)
        self.L.jmh_set_cur(self.h, _ptr(luma))

    def subplane(self, r, yy, xx):
        out = np.zeros((self.H + 2 * PAD_Y, self.W + 2 * PAD_X), np.uint8)
        self.L.jmh_get_subplane(self.h, C.c_int(r), C.c_int(yy), C.c_int(xx), _ptr(out))
        return out

    def spiral(self, n):
        out = np.zeros((n, 2), np.int16)
        self.L.jmh_spiral(self.h, C.c_int(n), _ptr(out))
        return out

    def mvbits(self, d):
        return self.L.jmh_mvbits(self.h, C.c_int(int(d)))

    def max_mvd(self):
        return self.L.jmh_max_mvd(self.h)

    def sad(self, pos_x, pos_y, blocktype, ref, cand_x, cand_y):
        return self.L.jmh_sad(self.h, pos_x, pos_y, blocktype, ref, cand_x, cand_y)

    def satd(self, pos_x, pos_y, blocktype, ref, cand_x, cand_y, test8x8=0):
        return self.L.jmh_satd(self.h, pos_x, pos_y, blocktype, ref, cand_x, cand_y, test8x8)

    def distortion(self, kind, n, diff):
        """the reference's distortion4x4/8x8{SAD,SSE,SATD} on one difference block"""
        self.L.jmh_distortion.restype = C.c_longlong
        d = np.ascontiguousarray(diff, np.int16)
        return int(self.L.jmh_distortion(C.c_int(kind), C.c_int(n), _ptr(d)))

    def hadamard4x4(self, diff):
        d = np.ascontiguousarray(diff, np.int16)
        return self.L.jmh_hadamard4x4(_ptr(d))

    def hadamard8x8(self, diff):
        d = np.ascontiguousarray(diff, np.int16)
        return self.L.jmh_hadamard8x8(_ptr(d))

    def bipred_search(self, jobs, lam, do_subpel=True, test8x8=False, wp=False, log_denom=0):
        """the reference's full_search_bipred_motion_estimation (+ sub_pel_bipred_motion_estimation) per BIPRED_JOB"""
        jobs = np.ascontiguousarray(jobs, BIPRED_JOB)
        out = np.zeros(len(jobs), BIPRED_RESULT)
        lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lam), (3,)), np.int32)
        self.L.jmh_bipred_search(self.h, C.c_int(len(jobs)), _ptr(jobs), _ptr(lam), C.c_int(int(do_subpel)), C.c_int(int(test8x8)),
                                 C.c_int(int(wp)), C.c_int(log_denom), _ptr(out))
        return out

    def chroma_pred4x4(self, plane, pix_c_x, opix_c_y, block_c_x, block_c_y, mv16):
        """one chroma 4x4 block through the unmodified OneComponentChromaPrediction4x4_regenerate; mv16 [4][4][2]"""
        plane = np.ascontiguousarray(plane, np.uint8); mv16 = np.ascontiguousarray(mv16, np.int16)
        out = np.zeros(16, np.uint8)
        self.L.jmh_chroma_pred4x4(self.h, C.c_int(plane.shape[1]), C.c_int(plane.shape[0]), _ptr(plane), C.c_int(pix_c_x), C.c_int(opix_c_y),
                                  C.c_int(block_c_x), C.c_int(block_c_y), _ptr(mv16), _ptr(out))
        return out.reshape(4, 4)

    def list_prediction_cost(self, mode, block, costs, ref_lambda, list=0):
        """the reference's list_prediction_cost (list 0 or 1) for one (mode, block): (best_ref, bmcost)"""
        c = np.ascontiguousarray(costs, np.int64)
        br = C.c_int(); bm = C.c_longlong()
        self.L.jmh_list_prediction_cost(self.h, C.c_int(list), C.c_int(mode), C.c_int(block), C.c_int(len(c)), _ptr(c), C.c_int(int(ref_lambda)), C.byref(br), C.byref(bm))
        return br.value, bm.value

    def search_frame(self, pred, center, lambda_factor, do_subpel=True, mb_first=0, mb_count=None):
        nmb = (self.W // 16) * (self.H // 16)
        if mb_count is None:
            mb_count = nmb - mb_first
        pred = np.ascontiguousarray(pred, np.int16)
        center = np.ascontiguousarray(center, np.int16)
        assert pred.shape == (nmb, self.nrefs, 41, 2) and center.shape == pred.shape
        lam = np.ascontiguousarray(lambda_factor, np.int32)
        mv_int = np.zeros_like(pred)
        mv_sub = np.zeros_like(pred)
        cost_int = np.zeros(pred.shape[:3], np.int64)
        cost_sub = np.zeros(pred.shape[:3], np.int64)
        self.L.jmh_search_frame(self.h, C.c_int(mb_first), C.c_int(mb_count), _ptr(pred), _ptr(center),
                                _ptr(lam), C.c_int(int(do_subpel)), _ptr(mv_int), _ptr(cost_int),
                                _ptr(mv_sub), _ptr(cost_sub))
        return mv_int, cost_int, mv_sub, cost_sub


# ---------------------------------------------------------------------------------------------
# version1 fractal search
# ---------------------------------------------------------------------------------------------
V1_SIZES = [(16, 16), (8, 8), (4, 4), (16, 8), (8, 16), (8, 4), (4, 8)]   # table order of compute_domain_Sum


class V1Ref:
    """The unmodified version1 compute.c/block_enc.c behind oracle/v1_harness.c (one per process:
    the reference keeps its state in globals)."""
    _made = False
    LIB = "libv1ref.so"

    def __init__(self, W, H, R, tol=(10.5, 8.0, 6.0)):
        if type(self)._made:
            raise RuntimeError("V1Ref: the reference state is global; one instance per process")
        self.L = _load(os.path.join(_HERE, "_ref", self.LIB))
        self.L.v1h_full_search.restype = C.c_double
        self.L.v1h_compute_rms.restype = C.c_double
        self.W, self.H, self.R = W, H, R
        r = self.L.v1h_init(C.c_int(W), C.c_int(H), C.c_int(R), C.c_double(tol[0]), C.c_double(tol[1]), C.c_double(tol[2]))
        assert r == 0
        type(self)._made = True

    def set_cur(self, y, u, v):
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        self.L.v1h_set_cur(_ptr(y), _ptr(u), _ptr(v))

    def set_ref(self, which, y, u, v, build_sums=True):
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        self.L.v1h_set_ref(C.c_int(which), _ptr(y), _ptr(u), _ptr(v), C.c_int(int(build_sums)))

    def full_search(self, which, bx, by, bsx, bsy, con):
        xy = (C.c_int * 2)(0, 0)
        so = (C.c_double * 2)()
        rms = self.L.v1h_full_search(C.c_int(which), bx, by, bsx, bsy, con, xy, so)
        return (xy[0], xy[1]), (so[0], so[1]), rms

    def compute_rms(self, which, bx, by, m, n, bsx, bsy, con):
        ab = (C.c_double * 2)()
        r = self.L.v1h_compute_rms(C.c_int(which), bx, by, m, n, bsx, bsy, con, ab)
        return r, ab[0], ab[1]

    def domain_table(self, sz, con, squares):
        """table sz (index into V1_SIZES) of plane set C as the reference holds it (double, cropped)."""
        w, h = (self.W, self.H) if con == 1 else (self.W // 2, self.H // 2)
        bw, bh = V1_SIZES[sz]
        out = np.zeros((h - bh + 1, w - bw + 1), np.float64)
        self.L.v1h_domain_table(C.c_int(sz), C.c_int(con), C.c_int(int(squares)), _ptr(out), C.c_int(out.shape[0]), C.c_int(out.shape[1]))
        return out

    def search_plane(self, which, con):
        """full_search of every range block, in the product's [mb][41] layout."""
        w, h = (self.W, self.H) if con == 1 else (self.W // 2, self.H // 2)
        mbw, mbh = (self.W // 16, self.H // 16) if con == 1 else (self.W // 16 // 2, self.H // 16 // 2)
        geo = partition_geometry()
        xy = np.zeros((mbw * mbh, 41, 2), np.int32)
        so = np.zeros((mbw * mbh, 41, 2), np.float64)
        rms = np.zeros((mbw * mbh, 41), np.float64)
        for mb in range(mbw * mbh):
            for p, (_, ox, oy, bw, bh) in enumerate(geo):
                a, b, r = self.full_search(which, (mb % mbw) * 16 + ox, (mb // mbw) * 16 + oy, bw, bh, con)
                xy[mb, p], so[mb, p], rms[mb, p] = a, b, r
        return xy, so, rms

    def decode_plane(self, con):
        """decode_one_macroblock of every macroblock of component con from the trees the last encode_mb calls left"""
        w, h = (self.W, self.H) if con == 1 else (self.W // 2, self.H // 2)
        out = np.zeros((h, w), np.uint8)
        self.L.v1h_decode_plane(C.c_int(con), _ptr(out))
        return out

    def reset_trans(self):
        """fresh TRANS_NODE trees (call between components: the harness indexes one tree array by macroblock number)"""
        self.L.v1h_reset_trans()

    def encode_mb(self, mb, con):
        out = np.zeros((21, 5), np.int32)
        outd = np.zeros((21, 2), np.float64)
        n = self.L.v1h_encode_mb(C.c_int(mb), C.c_int(con), _ptr(out), _ptr(outd), C.c_int(21))
        return out[:n], outd[:n]


V1_NODE = np.dtype([("block_type", np.int32), ("partition", np.int32), ("reference", np.int32), ("x", np.int32), ("y", np.int32),
                    ("pad", np.int32), ("scale", np.float64), ("offset", np.float64)])     # = b2fr_node (include/b2me.h)
assert V1_NODE.itemsize == 40


def v1_encode_plane(org, refC, xy, so, rms, tol):
    """F5: the partition cascade of every macroblock of one component plane from the search results of the four
    plane sets (xy/so [4][nmb][41][2], rms [4][nmb][41]); returns V1_NODE [nmb][21] (pre-order)."""
    org = np.ascontiguousarray(org, np.uint8); refC = np.ascontiguousarray(refC, np.uint8)
    h, w = org.shape
    mbw, mbh = w // 16, h // 16
    xy = np.ascontiguousarray(xy, np.int32); so = np.ascontiguousarray(so, np.float64); rms = np.ascontiguousarray(rms, np.float64)
    assert xy.shape == (4, mbw * mbh, 41, 2) and so.shape == xy.shape and rms.shape == xy.shape[:3]
    tol = np.ascontiguousarray(tol, np.float64)
    nodes = np.zeros((mbw * mbh, 21), V1_NODE)
    orc_lib().orc_v1_encode_plane(_ptr(org), _ptr(refC), C.c_int(w), C.c_int(mbw), C.c_int(mbh), _ptr(xy), _ptr(so), _ptr(rms),
                                  _ptr(tol), _ptr(nodes))
    return nodes


def v1_decode_plane(sets, nodes):
    """F8: decode_one_macroblock of every macroblock; sets = the four domain planes (C, H, M, N) of the component"""
    sets = [np.ascontiguousarray(p, np.uint8) for p in sets]
    h, w = sets[0].shape
    mbw, mbh = w // 16, h // 16
    nodes = np.ascontiguousarray(nodes, V1_NODE)
    assert nodes.shape == (mbw * mbh, 21)
    out = np.zeros((h, w), np.uint8)
    orc_lib().orc_v1_decode_plane(_ptr(sets[0]), _ptr(sets[1]), _ptr(sets[2]), _ptr(sets[3]), C.c_int(w), C.c_int(mbw), C.c_int(mbh),
                                  _ptr(nodes), _ptr(out))
    return out


def v1_box_table(img, bw, bh, squares):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h, w), np.float64)
    orc_lib().orc_v1_box_table(_ptr(img), C.c_int(w), C.c_int(h), C.c_int(bw), C.c_int(bh), C.c_int(int(squares)), _ptr(out))
    return out


def v1_range_table(img, squares):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h // 4, w // 4), np.float64)
    orc_lib().orc_v1_range_table(_ptr(img), C.c_int(w), C.c_int(h), C.c_int(int(squares)), _ptr(out))
    return out


def v1_search_plane(org, ref, R, have_sums=True, chroma=False, full_wh=None):
    """Restated full_search of every range block of one plane: (xy, scale_offset, rms) [mb][41]."""
    org = np.ascontiguousarray(org, np.uint8)
    ref = np.ascontiguousarray(ref, np.uint8)
    h, w = org.shape
    if chroma:
        W, H = full_wh
        mbw, mbh = W // 16 // 2, H // 16 // 2
    else:
        mbw, mbh = w // 16, h // 16
    xy = np.zeros((mbw * mbh, 41, 2), np.int32)
    so = np.zeros((mbw * mbh, 41, 2), np.float64)
    rms = np.zeros((mbw * mbh, 41), np.float64)
    orc_lib().orc_v1_search_plane(_ptr(org), _ptr(ref), C.c_int(w), C.c_int(h), C.c_int(mbw), C.c_int(mbh), C.c_int(R),
                                  C.c_int(int(have_sums)), _ptr(xy), _ptr(so), _ptr(rms))
    return xy, so, rms


class V1DropIn(V1Ref):
    """The same unmodified version1 objects with full_search served by libb2me.so through integration/v1/b2fr_v1_shim.c
    (oracle/_ref/libv1b2.so): call new_frame() after the pictures are set, then drive the reference's own cascade."""
    _made = False
    LIB = "libv1b2.so"

    def new_frame(self, have=(1, 0, 0, 0)):
        h = (C.c_int * 4)(*[int(x) for x in have])
        self.L.v1h_b2_new_frame(h)

    def calls(self):
        self.L.v1h_b2_calls.restype = C.c_long
        return self.L.v1h_b2_calls()


# ---------------------------------------------------------------------------------------------
# residual transform + quantisation
# ---------------------------------------------------------------------------------------------
class TQParams(C.Structure):
    """same layout as b2tq_params (include/b2me.h)"""
    _fields_ = [("qp", C.c_int32), ("mode", C.c_int32), ("cavlc", C.c_int32), ("field_scan", C.c_int32),
                ("disthres", C.c_int32), ("reserved", C.c_int32 * 3), ("scale", C.c_int32 * 64),
                ("offset", C.c_int32 * 64), ("invscale", C.c_int32 * 64)]


def tq_params(table, qp, mode=0, cavlc=1, field_scan=0, disthres=0):
    """table: int array [3][n*n] = ScaleComp, OffsetComp, InvScaleComp."""
    p = TQParams()
    p.qp, p.mode, p.cavlc, p.field_scan, p.disthres = qp, mode, cavlc, field_scan, disthres
    t = np.asarray(table, np.int64)
    for k, name in enumerate(("scale", "offset", "invscale")):
        arr = getattr(p, name)
        for i, v in enumerate(t[k].ravel()):
            arr[i] = int(v)
    return p


def tq(params, orig, pred, n):
    """restated oracle: n = 4 or 8.  Returns level, run, recon, cost, nonzero."""
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nblk = orig.shape[0]
    level = np.zeros((nblk, n * n), np.int16); run = np.zeros((nblk, n * n), np.uint8)
    recon = np.zeros((nblk, n * n), np.uint8); cost = np.zeros(nblk, np.int32); nz = np.zeros(nblk, np.uint8)
    f = orc_lib().orc_tq4x4 if n == 4 else orc_lib().orc_tq8x8
    f(C.byref(params), C.c_int(nblk), _ptr(orig), _ptr(pred), _ptr(level), _ptr(run), _ptr(recon), _ptr(cost), _ptr(nz))
    return level, run, recon, cost, nz


def tq16x16(params, orig, pred):
    """restated residual_transform_quant_luma_16x16: orig / pred [nmb][256] raster -> dc_level [nmb][16] i16, dc_run [nmb][16] u8,
    ac_level [nmb][16][16] i16, ac_run [nmb][16][16] u8, recon [nmb][256] u8, ac_coef [nmb] u8"""
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nmb = orig.shape[0]
    dl = np.zeros((nmb, 16), np.int16); dr = np.zeros((nmb, 16), np.uint8)
    al = np.zeros((nmb, 16, 16), np.int16); ar = np.zeros((nmb, 16, 16), np.uint8)
    rec = np.zeros((nmb, 256), np.uint8); ac = np.zeros(nmb, np.uint8)
    orc_lib().orc_tq16x16(C.byref(params), C.c_int(nmb), _ptr(orig), _ptr(pred), _ptr(dl), _ptr(dr), _ptr(al), _ptr(ar), _ptr(rec), _ptr(ac))
    return dl, dr, al, ar, rec, ac


def tq_chroma(params, orig, pred):
    """restated residual_transform_quant_chroma_4x4 (4:2:0, one plane): orig / pred [nmb][64] raster 8x8 -> dc_level [nmb][4] i16,
    dc_run [nmb][4] u8, ac_level [nmb][4][16] i16, ac_run [nmb][4][16] u8, recon [nmb][64] u8, cr_cbp [nmb] u8"""
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nmb = orig.shape[0]
    dl = np.zeros((nmb, 4), np.int16); dr = np.zeros((nmb, 4), np.uint8)
    al = np.zeros((nmb, 4, 16), np.int16); ar = np.zeros((nmb, 4, 16), np.uint8)
    rec = np.zeros((nmb, 64), np.uint8); cbp = np.zeros(nmb, np.uint8)
    orc_lib().orc_tq_chroma(C.byref(params), C.c_int(nmb), _ptr(orig), _ptr(pred), _ptr(dl), _ptr(dr), _ptr(al), _ptr(ar), _ptr(rec), _ptr(cbp))
    return dl, dr, al, ar, rec, cbp


class JMQuantRef:
    """The unmodified JM residual_transform_quant_luma_4x4/_8x8 behind oracle/jm_harness_tq.c."""

    def __init__(self, slice_type=0, symbol_mode=0):
        self.L = _load(os.path.join(_HERE, "_ref", "libjmref.so"))
        self.L.jmq_create.restype = _vp
        self.h = _vp(self.L.jmq_create(C.c_int(slice_type), C.c_int(symbol_mode)))

    def params(self, n, qp, intra):
        out = np.zeros((3, n * n), np.int32)
        self.L.jmq_params(self.h, C.c_int(int(n == 8)), C.c_int(qp), C.c_int(intra), _ptr(out))
        return out

    def tq(self, n, qp, intra, orig, pred):
        orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
        nblk, m = orig.shape[0], n * n
        level = np.zeros((nblk, m + 1), np.int32); run = np.zeros((nblk, m + 1), np.int32)
        recon = np.zeros((nblk, m), np.uint8); cost = np.zeros(nblk, np.int32); nz = np.zeros(nblk, np.int32)
        f = self.L.jmq_tq4x4 if n == 4 else self.L.jmq_tq8x8
        f(self.h, C.c_int(qp), C.c_int(intra), C.c_int(nblk), _ptr(orig), _ptr(pred), _ptr(level), _ptr(run), _ptr(recon), _ptr(cost), _ptr(nz))
        return level[:, :m].astype(np.int16), run[:, :m].astype(np.uint8), recon, cost, nz.astype(np.uint8)

    def tq_chroma(self, qpc, intra, uv, orig, pred):
        """the unmodified residual_transform_quant_chroma_4x4 on [nmb][64] raster 8x8 chroma blocks of plane uv"""
        orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
        nmb = orig.shape[0]
        dl = np.zeros((nmb, 5), np.int32); dr = np.zeros((nmb, 5), np.int32)
        al = np.zeros((nmb, 4, 16), np.int32); ar = np.zeros((nmb, 4, 16), np.int32)
        rec = np.zeros((nmb, 64), np.uint8); cbp = np.zeros(nmb, np.int32)
        self.L.jmq_tq_chroma(self.h, C.c_int(qpc), C.c_int(intra), C.c_int(uv), C.c_int(nmb), _ptr(orig), _ptr(pred), _ptr(dl), _ptr(dr), _ptr(al), _ptr(ar),
                             _ptr(rec), _ptr(cbp))
        return dl[:, :4].astype(np.int16), dr[:, :4].astype(np.uint8), al.astype(np.int16), ar.astype(np.uint8), rec, cbp.astype(np.uint8)

    def params_chroma(self, plane, qp, intra):
        """[3][16] table (ScaleComp, OffsetComp, InvScaleComp) of a chroma plane's 4x4 quantiser, for tq_params()"""
        out = np.zeros((3, 16), np.int32)
        self.L.jmq_params_chroma(self.h, C.c_int(plane), C.c_int(qp), C.c_int(intra), _ptr(out))
        return out

    def tq16x16(self, qp, orig, pred):
        """the unmodified residual_transform_quant_luma_16x16 (Intra16x16 luma) on [nmb][256] raster macroblocks"""
        orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
        nmb = orig.shape[0]
        dl = np.zeros((nmb, 17), np.int32); dr = np.zeros((nmb, 17), np.int32)
        al = np.zeros((nmb, 16, 16), np.int32); ar = np.zeros((nmb, 16, 16), np.int32)
        rec = np.zeros((nmb, 256), np.uint8); ac = np.zeros(nmb, np.int32)
        self.L.jmq_tq16x16(self.h, C.c_int(qp), C.c_int(nmb), _ptr(orig), _ptr(pred), _ptr(dl), _ptr(dr), _ptr(al), _ptr(ar), _ptr(rec), _ptr(ac))
        return dl[:, :16].astype(np.int16), dr[:, :16].astype(np.uint8), al.astype(np.int16), ar.astype(np.uint8), rec, ac.astype(np.uint8)


def have_v1tq():
    return os.path.exists(os.path.join(_HERE, "_ref", "libv1tq.so"))


def v1_dct_luma(qp, slice_type, orig, pred):
    """The UNMODIFIED version1 dct_luma (V1/src/block.c:836-1045) behind oracle/v1_harness_tq.c on [nblk][16] raster blocks:
    level, run, recon, cost, nonzero in the layout of tq()."""
    L = _load(os.path.join(_HERE, "_ref", "libv1tq.so"))
    orig = np.ascontiguousarray(orig, np.uint8); pred = np.ascontiguousarray(pred, np.uint8)
    nblk = orig.shape[0]
    level = np.zeros((nblk, 16), np.int32); run = np.zeros((nblk, 16), np.int32)
    recon = np.zeros((nblk, 16), np.uint8); cost = np.zeros(nblk, np.int32); nz = np.zeros(nblk, np.int32)
    L.v1tq_dct_luma(C.c_int(qp), C.c_int(slice_type), C.c_int(nblk), _ptr(orig), _ptr(pred), _ptr(level), _ptr(run), _ptr(recon), _ptr(cost), _ptr(nz))
    return level.astype(np.int16), run.astype(np.uint8), recon, cost, nz.astype(np.uint8)


# ---- fractal pool matching (oracle/b2_oracle_pool.c: DEFINES the pool-mode semantics, see its header) ----
def pool_positions(dw, dh, nd):
    xy = np.zeros((nd, 2), np.int32)
    orc_lib().orc_pool_positions(C.c_int(dw), C.c_int(dh), C.c_int(nd), _ptr(xy))
    return xy


def pool_domain_block(plane, x, y):
    plane = np.ascontiguousarray(plane, np.uint8)
    blk = np.zeros(64, np.uint8)
    orc_lib().orc_pool_domain_block(_ptr(plane), C.c_int(plane.shape[1]), C.c_int(int(x)), C.c_int(int(y)), _ptr(blk))
    return blk


def pool_iso(block64, iso):
    src = np.ascontiguousarray(block64, np.uint8).reshape(64)
    dst = np.zeros(64, np.uint8)
    orc_lib().orc_pool_iso(_ptr(src), C.c_int(iso), _ptr(dst))
    return dst


def pool_rms_double(r64, d64):
    """compute_rms's floating-point expression (V1/src/compute.c:156-182) for one pair: (rms, alpha, beta)."""
    L = orc_lib()
    L.orc_pool_rms_double.restype = C.c_double
    al, be = C.c_double(), C.c_double()
    r = np.ascontiguousarray(r64, np.uint8); d = np.ascontiguousarray(d64, np.uint8)
    rms = L.orc_pool_rms_double(_ptr(r), _ptr(d), C.byref(al), C.byref(be))
    return rms, al.value, be.value


def pool_pair(r64, d64):
    """orc_pool_pair: the exact integer fit of one (range, domain) pair -> (accepted, aq, G, beta, err_num)"""
    L = orc_lib()
    r = np.ascontiguousarray(r64, np.uint8).reshape(64); d = np.ascontiguousarray(d64, np.uint8).reshape(64)
    aq, G = C.c_int(), C.c_int64()
    ok = L.orc_pool_pair(_ptr(r), _ptr(d), C.byref(aq), C.byref(G))
    sr = int(r.astype(np.int64).sum()); sr2 = int((r.astype(np.int64) ** 2).sum())
    beta = int(L.orc_pool_quan_a(C.c_int(sr // 64)))
    err = 640000 * (sr2 - 2 * beta * sr + 64 * beta * beta) - G.value if ok else -1
    return bool(ok), aq.value, G.value if ok else None, beta, err


def pool_search(range_plane, domain_plane, nd):
    r = np.ascontiguousarray(range_plane, np.uint8); d = np.ascontiguousarray(domain_plane, np.uint8)
    rh, rw = r.shape; dh, dw = d.shape
    nr = (rw // 8) * (rh // 8)
    dom = np.zeros(nr, np.int32); iso = np.zeros(nr, np.uint8)
    aq = np.zeros(nr, np.int16); beta = np.zeros(nr, np.int16); err = np.zeros(nr, np.int64)
    orc_lib().orc_pool_search(_ptr(r), C.c_int(rw), C.c_int(rh), C.c_int(rw), _ptr(d), C.c_int(dw), C.c_int(dh), C.c_int(dw),
                              C.c_int(nd), _ptr(dom), _ptr(iso), _ptr(aq), _ptr(beta), _ptr(err))
    return dom, iso, aq, beta, err
