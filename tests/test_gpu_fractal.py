"""GPU parity tests (through the C ABI): CUDA fractal range/domain search vs golden vectors of the
unmodified version1 sources and vs the restated oracle.  Integers (x, y, visit order) and the
quantised scale/offset are compared exactly; rms is compared bit-for-bit too (the kernel evaluates
compute_rms with round-to-nearest intrinsics in the reference's operand order) -- the stated
tolerance for the floating-point quantities is therefore 0 ulp."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth
from oracle.gen_golden_v1 import CASES

pytestmark = pytest.mark.gpu


def _ctx(name):
    W, H, R, seed, shift, gain, offset, planes = CASES[name]
    ref, cur = synth.yuv_pair(W, H, seed=seed, shift=shift, gain=gain, offset=offset)
    f = api.FractalSearcher(W, H, R)
    f.set_domain(0, *ref, build_sums=True)
    f.set_range(*cur)
    return f, W, H, R, ref, cur, planes


@pytest.mark.parametrize("name", ["small", "qcif", "cif"])
def test_search_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"v1_harness_{name}.npz"))
    f, W, H, R, ref, cur, planes = _ctx(name)
    for which, con in planes:
        xy, so, rms = f.search_plane(which, con)
        assert (xy == g[f"xy_{which}_{con}"]).all(), (which, con)
        assert (so == g[f"so_{which}_{con}"]).all(), (which, con)
        assert (rms == g[f"rms_{which}_{con}"]).all(), (which, con)
    assert f.launch_count() > 0


def test_sum_tables_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "v1_harness_small.npz"))
    f, W, H, R, ref, cur, _ = _ctx("small")
    for sz, (bw, bh) in enumerate(oracle.V1_SIZES):
        for con in (1, 2):
            for sq in (0, 1):
                gt = g[f"tab_{sz}_{con}_{sq}"]
                t = f.domain_table(0, con, bw, bh, sq)
                assert (t[:gt.shape[0], :gt.shape[1]] == gt).all(), (sz, con, sq)
                assert (t[gt.shape[0]:] == 0).all() and (t[:, gt.shape[1]:] == 0).all()
    for con in (1, 2, 3):
        for sq in (0, 1):
            assert (f.range_table(con, sq) == oracle.v1_range_table(cur[con - 1], sq)).all()


def test_full_search_dropin_call(golden_dir):
    """b2fr_full_search mirrors one call of full_search(): x,y untouched when (0,0) wins (Q-F11)."""
    g = np.load(os.path.join(golden_dir, "v1_harness_qcif.npz"))
    f, W, H, R, ref, cur, _ = _ctx("qcif")
    geo = oracle.partition_geometry()
    rng = np.random.default_rng(0)
    for _ in range(200):
        mb, p = int(rng.integers(0, 99)), int(rng.integers(0, 41))
        _, ox, oy, bw, bh = geo[p]
        xy, so, rms = f.full_search(0, (mb % 11) * 16 + ox, (mb // 11) * 16 + oy, bw, bh, 1, xy=(99, -99))
        gx = g["xy_0_1"][mb, p]
        exp = (99, -99) if (gx == 0).all() else tuple(gx)
        assert xy == exp and so == tuple(g["so_0_1"][mb, p]) and rms == g["rms_0_1"][mb, p]
    with pytest.raises(api.B2Error):
        f.full_search(0, 3, 0, 8, 8, 1)               # not on the range grid


@pytest.mark.parametrize("W,H,R,seed,gain,offset,noise", [
    (352, 288, 7, 5, 0.9, 10.0, 1.0),          # BASELINE config 2, fresh seed
    (64, 64, 12, 6, 1.6, -40.0, 6.0),          # larger window than the picture margin, strong gain
    (48, 32, 3, 8, -0.5, 200.0, 0.0),          # negative correlation -> negative alpha, QUAN_A on negatives (Q-F4)
    (1920, 1088, 7, 9, 1.0, 0.0, 2.0),         # 1080p: full size through size-independent properties
])
def test_search_matches_oracle(W, H, R, seed, gain, offset, noise):
    ref, cur = synth.yuv_pair(W, H, seed=seed, shift=(2, -2), gain=gain, offset=offset, noise=noise)
    f = api.FractalSearcher(W, H, R)
    f.set_domain(0, *ref, build_sums=True)
    f.set_range(*cur)
    big = W * H > 352 * 288
    for which, con in ((0, 1), (1, 1), (0, 2), (0, 3)):
        xy, so, rms = f.search_plane(which, con)
        if big:
            # full size: the oracle is too slow; check properties that hold for every input
            assert (np.abs(xy) <= R).all()
            assert (np.round(so[..., 0] * 100) % 5 == 0).all() and (so[..., 1] % 5 == 0).all()
            ok = rms < 1e30
            assert (so[..., 0][ok] >= -2.35).all() and (so[..., 0][ok] <= 4.0).all()
            # crop invariance: an interior macroblock searched by the oracle on the 48x48 crop around
            # it (3x3 macroblocks, the block in the centre one) gives the identical result (R <= 16)
            if con == 1 and which == 0:
                rng = np.random.default_rng(1)
                mbw = W // 16
                for _ in range(40):
                    mb = int(rng.integers(0, xy.shape[0]))
                    x0, y0 = (mb % mbw) * 16, (mb // mbw) * 16
                    if x0 < 16 or y0 < 16 or x0 + 32 > W or y0 + 32 > H:
                        continue
                    sub = (slice(y0 - 16, y0 + 32), slice(x0 - 16, x0 + 32))
                    cxy, cso, crms = oracle.v1_search_plane(cur[0][sub], ref[0][sub], R, True)
                    assert (cxy[4] == xy[mb]).all() and (cso[4] == so[mb]).all() and (crms[4] == rms[mb]).all()
            continue
        org = cur[con - 1]
        dom = ref[con - 1] if which == 0 else np.zeros_like(ref[con - 1])
        exy, eso, erms = oracle.v1_search_plane(org, dom, R, have_sums=(which == 0), chroma=(con > 1), full_wh=(W, H))
        assert (xy == exy).all() and (so == eso).all() and (rms == erms).all(), (which, con)


@pytest.mark.parametrize("name", ["loaded", "zero"])
def test_cascade_matches_reference_golden(golden_dir, name):
    """F5 on the device (b2fr_encode_plane) against the trees of the unmodified encode_one_macroblock"""
    from oracle import gen_golden_cascade as gc
    g = np.load(os.path.join(golden_dir, f"v1_cascade_{name}.npz"))
    cur, sets, W, H, R, tol, loaded = gc.inputs(name)
    f = api.FractalSearcher(W, H, R)
    for s in range(4):
        if s == 0 or loaded:
            f.set_domain(s, *sets[s], build_sums=True)
    f.set_range(*cur)
    for con in (1, 2, 3):
        nodes, exp = f.encode_plane(con, tol), g[f"nodes_{con}"]
        for k in ("block_type", "partition", "reference", "x", "y", "scale", "offset"):
            assert (nodes[k] == exp[k]).all(), (con, k, int((nodes[k] != exp[k]).sum()))
        # F8: the prediction of decode_one_macroblock, from the trees still on the device and from trees handed in
        assert (f.decode_plane(con) == g[f"rec_{con}"]).all(), con
        assert (f.decode_plane(con, exp) == g[f"rec_{con}"]).all(), con


def test_cascade_matches_oracle_cif():
    """BASELINE config 2 geometry (CIF, +-7): static noisy content so that the cascade splits; oracle cascade on the oracle's searches"""
    W, H, R, tol = 352, 288, 7, (3.5, 4.5, 2.0)
    ref, cur = synth.yuv_pair(W, H, seed=11, shift=(0, 0), gain=1.0, offset=0.0, noise=4.0)
    hset = [np.roll(p, (1, -1), (0, 1)) for p in ref]
    f = api.FractalSearcher(W, H, R)
    f.set_domain(0, *ref, build_sums=True)
    f.set_domain(1, *hset, build_sums=True)
    f.set_range(*cur)
    for con in (1, 2):
        nodes = f.encode_plane(con, tol)
        org = cur[con - 1]
        doms = [ref[con - 1], hset[con - 1], np.zeros_like(org), np.zeros_like(org)]
        res = [oracle.v1_search_plane(org, doms[s], R, have_sums=(s < 2), chroma=(con > 1), full_wh=(W, H)) for s in range(4)]
        exp = oracle.v1_encode_plane(org, ref[con - 1], np.stack([r[0] for r in res]), np.stack([r[1] for r in res]),
                                     np.stack([r[2] for r in res]), tol)
        for k in ("block_type", "partition", "reference", "x", "y", "scale", "offset"):
            assert (nodes[k] == exp[k]).all(), (con, k)
        rec = f.decode_plane(con)
        assert (rec == oracle.v1_decode_plane(doms, exp)).all(), con
        # hand-made trees reach the macroblock-level rectangles the encoder never leaves (partition 1 / 2)
        alt = exp.copy(); alt[:, 0]["partition"] = np.arange(len(alt)) % 3
        assert (f.decode_plane(con, alt) == oracle.v1_decode_plane(doms, alt)).all(), con
        if con == 1:
            assert len(np.unique(nodes[:, 0]["partition"])) == 2 and len(np.unique(nodes[:, [1, 6, 11, 16]]["partition"])) >= 3
