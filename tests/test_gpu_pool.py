"""GPU parity tests (through the C ABI): tensor-core pool matching (k_frac_pool) vs oracle/b2_oracle_pool.c."""
import numpy as np
import pytest

import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu


def _check(rp, dp, nd):
    s = api.PoolSearcher(rp.shape[1], rp.shape[0], dp.shape[1], dp.shape[0], nd)
    assert (s.positions() == oracle.pool_positions(dp.shape[1], dp.shape[0], nd)).all()
    s.set_planes(rp, dp)
    got = s.search()
    exp = oracle.pool_search(rp, dp, nd)
    for g, e, n in zip(got, exp, ("best_dom", "best_iso", "aq", "beta", "err_num")):
        assert (g == e).all(), (n, int((g != e).sum()), np.flatnonzero(g != e)[:5], g[g != e][:5], e[g != e][:5])
    return s


@pytest.mark.parametrize("W,H,nd,seed", [(64, 48, 200, 3), (96, 64, 256, 4), (176, 144, 1000, 5), (128, 128, 513, 6)])
def test_pool_matches_oracle(W, H, nd, seed):
    (yr, _, _), (yc, _, _) = synth.yuv_pair(W, H, seed=seed, shift=(-3, 2), gain=0.8, offset=12.0)
    s = _check(yc, yr, nd)
    st = s.stats()
    assert st["chunks"] > 0 and st["exact_evals"] > 0


def test_pool_noise_flat_and_saturated():
    """Pure noise (no good match: many near-ties), flat domains (det == 0 -> alpha 0), saturated ranges."""
    rng = np.random.default_rng(1)
    rp = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    dp = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    dp[:24, :40] = 77
    rp[:8, :8] = 255; rp[8:16, :8] = 0
    _check(rp, dp, 300)
    _check(np.full((16, 16), 9, np.uint8), np.full((32, 32), 200, np.uint8), 5)     # everything flat: G == 0 ties -> index 0


def test_pool_planted_and_different_plane_sizes():
    rng = np.random.default_rng(5)
    dp = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    nd = 60
    xy = oracle.pool_positions(96, 64, nd)
    rp = np.zeros((16, 32), np.uint8)
    want = [7, 23, 41, 58, 3, 11, 30, 52]
    for k, p in enumerate(want):
        d = oracle.pool_domain_block(dp, *xy[p]).astype(np.float64)
        blk = np.clip(np.rint(0.5 * (d - d.mean()) + 120), 0, 255).astype(np.uint8).reshape(8, 8)
        rp[(k // 4) * 8:(k // 4) * 8 + 8, (k % 4) * 8:(k % 4) * 8 + 8] = blk.T if k & 1 else blk
    s = _check(rp, dp, nd)
    dom, iso, aq, beta, err = s.search()
    assert list(dom) == want and (aq == 50).all()


def test_pool_isometry_and_index_ties():
    """Range blocks that are symmetric under flips / transposition make several isometry rows identical (equal G for
    every domain: the lowest isometry must win although the rows share one threshold), and a domain plane made of
    repeated tiles makes distinct pool entries identical (equal G: the lowest pool index must win, across MMA tiles
    and epilogue groups)."""
    rng = np.random.default_rng(9)
    tile = rng.integers(0, 256, (16, 16), dtype=np.uint8)
    dp = np.tile(tile, (8, 10))                                   # 128 x 160: many identical domain blocks
    rp = np.zeros((32, 64), np.uint8)
    for k in range(32):
        b = rng.integers(0, 256, (8, 8)).astype(np.int64)
        kind = k % 4
        if kind == 0: b = b + b[:, ::-1]                          # left-right symmetric
        elif kind == 1: b = b + b.T                               # symmetric under transposition
        elif kind == 2: b = b + b[::-1, :] + b[:, ::-1] + b[::-1, ::-1]   # both flips (and the half turn)
        else: b = b + b.T + b[::-1, ::-1] + b[::-1, ::-1].T       # both diagonals
        b = (b * 255 // max(1, b.max())).astype(np.uint8)
        rp[(k // 8) * 8:(k // 8) * 8 + 8, (k % 8) * 8:(k % 8) * 8 + 8] = b
    _check(rp, dp, 1200)
    _check(rp, dp, 300)


def test_pool_error_codes():
    with pytest.raises(api.B2Error):
        api.PoolSearcher(60, 48, 64, 64, 10)          # range plane not a multiple of 8
    with pytest.raises(api.B2Error):
        api.PoolSearcher(64, 48, 8, 8, 10)            # domain plane smaller than a domain block
