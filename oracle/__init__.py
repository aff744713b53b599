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


def partition_geometry():
    """[(blocktype, ox, oy, w, h)] for the 41 partitions of an MB."""
    L = orc_lib()
    out = []
    for p in range(41):
        v = [C.c_int() for _ in range(5)]
        L.orc_partition_geometry(C.c_int(p), *[C.byref(x) for x in v])
        out.append(tuple(x.value for x in v))
    return out


def subpel_planes(luma):
    H, W = luma.shape
    out = np.zeros((4, 4, H + 2 * PAD_Y, W + 2 * PAD_X), np.uint8)
    luma = np.ascontiguousarray(luma, np.uint8)
    orc_lib().orc_subpel_planes(_ptr(luma), C.c_int(W), C.c_int(H), _ptr(out))
    return out


def hadamard4x4(diff):
    d = np.ascontiguousarray(diff, np.int16)
    return orc_lib().orc_hadamard4x4(_ptr(d))


def hadamard8x8(diff):
    d = np.ascontiguousarray(diff, np.int16)
    return orc_lib().orc_hadamard8x8(_ptr(d))


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

    def hadamard4x4(self, diff):
        d = np.ascontiguousarray(diff, np.int16)
        return self.L.jmh_hadamard4x4(_ptr(d))

    def hadamard8x8(self, diff):
        d = np.ascontiguousarray(diff, np.int16)
        return self.L.jmh_hadamard8x8(_ptr(d))

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
