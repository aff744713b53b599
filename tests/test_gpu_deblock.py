"""In-loop deblocking filter: b2dbk_frame (k_deblock, 2:1 macroblock wavefront) against the oracle and against what
the unmodified DeblockFrame of JM 18.5 left on real encoder pictures (tests/golden/jm_deblock.npz, oracle/gen_golden_dbk.py)."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "jm_deblock.npz")


def golden_pictures():
    g = np.load(GOLD)
    for tag in "pbt":
        for i in range(int(g[f"{tag}_n"])):
            yield f"{tag}{i}", [g[f"{tag}{i}_{n}0"] for n in "yuv"], g[f"{tag}{i}_mbs"], g[f"{tag}{i}_blks"], [g[f"{tag}{i}_{n}1"] for n in "yuv"]


def test_deblock_matches_reference_golden():
    """I, P and B pictures of stock lencod runs (intra / inter / skipped macroblocks, two lists, 8x8 transform, filter offsets)"""
    n = 0
    for name, before, mbs, blks, after in golden_pictures():
        got = api.deblock_frame(*before, mbs, blks)
        for g, a, pl in zip(got, after, "yuv"):
            assert (g == a).all(), (name, pl, int((g != a).sum()))
        n += 1
    assert n == 12


def random_records(W, H, rng, p_intra=0.15, p_t8=0.2, p_off=0.05):
    nmb, nb = (W // 16) * (H // 16), (W // 4) * (H // 4)
    mbs = np.zeros(nmb, synth.DBK_MB)
    mbs["intra"] = rng.random(nmb) < p_intra
    mbs["qp"] = rng.integers(24, 52, nmb); mbs["qpc_u"] = np.minimum(mbs["qp"], 39); mbs["qpc_v"] = np.maximum(mbs["qpc_u"].astype(int) - 2, 0)
    mbs["transform8x8"] = rng.random(nmb) < p_t8
    mbs["disable"] = rng.random(nmb) < p_off
    mbs["alpha_off"] = rng.integers(-3, 4) * 2; mbs["beta_off"] = rng.integers(-3, 4) * 2
    mbs["cbp_blk"] = np.where(rng.random(nmb) < 0.5, rng.integers(0, 65536, nmb), 0)
    blks = np.zeros(nb, synth.DBK_BLK)
    base = rng.integers(-6, 7, (H // 16, W // 16, 2, 2))                           # per macroblock, then small per-block deviations
    mv = np.repeat(np.repeat(base, 4, axis=0), 4, axis=1) + (rng.random((H // 4, W // 4, 2, 2)) < 0.3) * rng.integers(-5, 6, (H // 4, W // 4, 2, 2))
    blks["mv"] = mv.reshape(nb, 2, 2)
    blks["ref"] = rng.integers(-1, 3, (nb, 2))
    return mbs, blks


@pytest.mark.parametrize("W,H,seed", [(64, 48, 1), (176, 144, 2), (1920, 1088, 3)])
def test_deblock_random_pictures_match_oracle(W, H, seed):
    """Random records at sizes up to 1080p: the wavefront over 68 concurrently running macroblock rows against the serial oracle"""
    rng = np.random.default_rng(seed)
    fr = synth.yuv420_sequence(W, H, 1, seed=seed)
    y = np.frombuffer(fr, np.uint8, W * H).reshape(H, W)
    u = np.frombuffer(fr, np.uint8, W * H // 4, W * H).reshape(H // 2, W // 2)
    v = np.frombuffer(fr, np.uint8, W * H // 4, W * H * 5 // 4).reshape(H // 2, W // 2)
    y = (y.astype(int) + rng.integers(-4, 5, y.shape)).clip(0, 255).astype(np.uint8)      # blocking-like noise so that the filters act
    mbs, blks = random_records(W, H, rng)
    exp = oracle.deblock_frame(y, u, v, mbs, blks)
    for rep in range(2 if W < 1000 else 3):                                            # repeated: a race would not repeat
        got = api.deblock_frame(y, u, v, mbs, blks)
        for g, e, pl in zip(got, exp, "yuv"):
            assert (g == e).all(), (pl, rep, int((g != e).sum()))
    assert sum(int((e != o).sum()) for e, o in zip(exp, (y, u, v))) > W * H // 100


def test_deblock_errors():
    y = np.zeros((40, 64), np.uint8); u = np.zeros((20, 32), np.uint8)
    with pytest.raises(RuntimeError):
        api.deblock_frame(y, u, u, np.zeros(8, synth.DBK_MB), np.zeros(160, synth.DBK_BLK))     # 40 rows: not a multiple of 16
