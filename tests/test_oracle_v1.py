"""CPU tests: the restated fractal oracle (oracle/b2_oracle_v1.c) against golden vectors generated
from the UNMODIFIED version1 sources (oracle/gen_golden_v1.py) and, when oracle/_ref/libv1ref.so
is present, against the reference objects live on a fresh seed."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import synth
from oracle.gen_golden_v1 import CASES


def _planes(name):
    W, H, R, seed, shift, gain, offset, planes = CASES[name]
    ref, cur = synth.yuv_pair(W, H, seed=seed, shift=shift, gain=gain, offset=offset)
    return W, H, R, ref, cur, planes


def _restated(W, H, R, ref, cur, which, con):
    org = cur[con - 1]
    dom = ref[con - 1] if which == 0 else np.zeros_like(ref[con - 1])     # H/M/N: zero planes (Q-F3)
    return oracle.v1_search_plane(org, dom, R, have_sums=(which == 0), chroma=(con > 1), full_wh=(W, H))


@pytest.mark.parametrize("name", ["small", "qcif", "cif"])
def test_restated_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"v1_harness_{name}.npz"))
    W, H, R, ref, cur, planes = _planes(name)
    for which, con in planes:
        xy, so, rms = _restated(W, H, R, ref, cur, which, con)
        assert (xy == g[f"xy_{which}_{con}"]).all(), (which, con)
        assert (so == g[f"so_{which}_{con}"]).all(), (which, con)           # alpha, beta after QUAN_A: exact
        assert (rms == g[f"rms_{which}_{con}"]).all(), (which, con)         # same IEEE-754 doubles, bit for bit


def test_golden_exercises_the_interesting_paths(golden_dir):
    g = np.load(os.path.join(golden_dir, "v1_harness_qcif.npz"))
    xy, so, rms = g["xy_0_1"], g["so_0_1"], g["rms_0_1"]
    assert (xy != 0).any() and (np.abs(xy) <= 7).all()
    assert len(np.unique(so[..., 0])) > 5            # several quantised scales
    assert (rms >= 1e30).any() or (so[..., 0] == 0).any()
    h = g["rms_1_1"]                                  # zero H plane: alpha = 0, rms = sum (r - beta)^2 - like
    assert (g["so_1_1"][..., 0] == 0).all() and (g["xy_1_1"] == 0).all() and np.isfinite(h).all()


def test_sum_tables_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "v1_harness_small.npz"))
    W, H, R, ref, cur, _ = _planes("small")
    for sz, (bw, bh) in enumerate(oracle.V1_SIZES):
        for con in (1, 2):
            for sq in (0, 1):
                t = oracle.v1_box_table(ref[con - 1], bw, bh, sq)
                gt = g[f"tab_{sz}_{con}_{sq}"]
                assert (t[:gt.shape[0], :gt.shape[1]] == gt).all(), (sz, con, sq)


@pytest.mark.skipif(not oracle.have_v1ref(), reason="oracle/_ref/libv1ref.so not built (needs /root/reference)")
def test_restated_matches_reference_live():
    W, H, R = 64, 48, 5                                # geometry is global in the reference: one per process
    v = oracle.V1Ref(W, H, R)
    for seed in (11, 12):
        ref, cur = synth.yuv_pair(W, H, seed=seed, shift=(-2, 3), gain=1.3, offset=-20.0, noise=4.0)
        v.set_ref(0, *ref, build_sums=True)
        v.set_cur(*cur)
        for con in (1, 2, 3):
            a = v.search_plane(0, con)
            b = oracle.v1_search_plane(cur[con - 1], ref[con - 1], R, True, chroma=(con > 1), full_wh=(W, H))
            for x, y in zip(a, b):
                assert (x == y).all()


def _cascade_results(cur, sets, W, H, R, loaded, con):
    org = cur[con - 1]
    res = [oracle.v1_search_plane(org, sets[s][con - 1], R, have_sums=(s == 0 or loaded), chroma=(con > 1), full_wh=(W, H)) for s in range(4)]
    return org, np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[2] for r in res])


@pytest.mark.parametrize("name", ["loaded", "zero"])
def test_cascade_matches_reference_golden(golden_dir, name):
    """F5: the TRANS_NODE trees the unmodified encode_one_macroblock leaves (block_enc.c:508), every field of every node"""
    from oracle import gen_golden_cascade as gc
    g = np.load(os.path.join(golden_dir, f"v1_cascade_{name}.npz"))
    cur, sets, W, H, R, tol, loaded = gc.inputs(name)
    for con in (1, 2, 3):
        org, xy, so, rms = _cascade_results(cur, sets, W, H, R, loaded, con)
        nodes = oracle.v1_encode_plane(org, sets[0][con - 1], xy, so, rms, tol)
        exp = g[f"nodes_{con}"]
        for f in ("block_type", "partition", "reference", "x", "y", "scale", "offset"):
            assert (nodes[f] == exp[f]).all(), (con, f)
        # F8: decode_one_macroblock (block_dec.c:20) on those trees
        rec = oracle.v1_decode_plane([sets[s][con - 1] for s in range(4)], exp)
        assert (rec == g[f"rec_{con}"]).all(), con
    n1 = g["nodes_1"]
    if name == "loaded":      # the fixture reaches every outcome of the cascade
        assert set(np.unique(n1[:, 0]["partition"])) == {0, 3}
        assert set(np.unique(n1[:, [1, 6, 11, 16]]["partition"])) == {0, 1, 2, 3}
        assert set(np.unique(n1["reference"])) == {0, 1, 2, 3}
