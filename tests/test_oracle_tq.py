"""CPU tests: restated transform/quant oracle (oracle/b2_oracle_tq.c) and the product's default
parameter tables against golden vectors of the UNMODIFIED JM objects (oracle/gen_golden_tq.py)."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "jm_tq.npz"))
    for ci, (n, qp, intra, st, sm, seed, nblk) in enumerate(g["cases"]):
        yield g, ci, int(n), int(qp), int(intra), int(st), int(sm), int(seed), int(nblk)


def test_restated_matches_reference_golden(golden_dir):
    for g, ci, n, qp, intra, st, sm, seed, nblk in _cases(golden_dir):
        orig, pred = synth.residual_blocks(nblk, n, seed)
        p = oracle.tq_params(g[f"c{ci}_params"], qp, mode=0, cavlc=int(sm == 0))
        lv, rn, rec, cost, nz = oracle.tq(p, orig, pred, n)
        for a, name in ((lv, "level"), (rn, "run"), (rec, "recon"), (cost, "cost"), (nz, "nz")):
            assert (a == g[f"c{ci}_{name}"]).all(), (ci, name)


def test_default_params_match_reference_tables(golden_dir):
    """b2tq_default_params (pure host code of the product library) reproduces the LevelQuantParams the
    reference derives from q_matrix.c / q_offsets.c for the flat-matrix, default-offset configuration."""
    for g, ci, n, qp, intra, st, sm, seed, nblk in _cases(golden_dir):
        if st == 2 and intra == 0:
            continue                                    # inter tables are not refreshed in an I slice
        p = api.tq_default_params(n, qp, 2 if (intra and st == 2) else intra)
        assert (api.tq_params_table(p, n) == g[f"c{ci}_params"]).all(), (ci, n, qp, intra)


def test_transform_roundtrip_properties():
    """forward4x4 is linear and inverse4x4(forward4x4(x) scaled) reproduces x: the H.264 core
    transform pair satisfies  inverse(forward(x) * [16,20,25 pattern] ... ) -- checked via the
    quantiser at qp where scale*invscale is exact: at qp 0..5 reconstruction error is <= 1."""
    rng = np.random.default_rng(3)
    orig = rng.integers(0, 256, (2000, 16), dtype=np.uint8)
    pred = rng.integers(0, 256, (2000, 16), dtype=np.uint8)
    p = api.tq_default_params(4, 0, 2)
    _, _, rec, _, _ = oracle.tq(oracle.tq_params(api.tq_params_table(p, 4), 0), orig, pred, 4)
    assert np.abs(rec.astype(int) - orig.astype(int)).max() <= 1


@pytest.mark.skipif(not oracle.have_jmref(), reason="oracle/_ref/libjmref.so not built (needs /root/reference)")
def test_restated_matches_reference_live():
    r = oracle.JMQuantRef(0, 0)
    for n, qp, intra, seed in ((4, 22, 0, 101), (4, 37, 1, 102), (8, 30, 0, 103), (8, 12, 1, 104)):
        orig, pred = synth.residual_blocks(300, n, seed)
        a = r.tq(n, qp, intra, orig, pred)
        b = oracle.tq(oracle.tq_params(r.params(n, qp, intra), qp), orig, pred, n)
        for x, y in zip(a, b):
            assert (x == y).all(), (n, qp, intra)


def test_restated_intra16x16_matches_reference_golden(golden_dir):
    """residual_transform_quant_luma_16x16 (DC Hadamard, quant_dc4x4 / quant_ac4x4) restated vs the unmodified JM objects"""
    from oracle.gen_golden_tq16 import CASES, NMB, macroblocks
    g = np.load(os.path.join(golden_dir, "jm_tq16.npz"))
    for ci, (qp, sm, seed) in enumerate(CASES):
        orig, pred = macroblocks(NMB, seed)
        p = oracle.tq_params(g[f"c{ci}_params"], qp, mode=0, cavlc=int(sm == 0))
        got = oracle.tq16x16(p, orig, pred)
        for a, name in zip(got, ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "ac_coef")):
            assert (a == g[f"c{ci}_{name}"]).all(), (qp, sm, name)
        assert (api.tq_params_table(api.tq_default_params(4, qp, 2), 4) == g[f"c{ci}_params"]).all()   # the product's default intra table


def _chroma_cmp(got, exp, tag):
    dl, dr, al, ar, rec, cbp = got
    for a, b, name in zip(got, exp, ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "cr_cbp")):
        if name == "ac_run":            # the reference leaves the runs of dropped AC levels behind: compare where a level stands
            assert (a[al != 0] == b[al != 0]).all(), (tag, name)
        elif name == "dc_run":
            assert (a[dl != 0] == b[dl != 0]).all(), (tag, name)
        else:
            assert (a == b).all(), (tag, name, int((a != b).sum()))


def test_restated_chroma_matches_reference_golden(golden_dir):
    """residual_transform_quant_chroma_4x4 (hadamard2x2, quant_dc2x2_normal, quant_ac4x4_normal, the chroma coefficient-cost rule)
    restated vs the unmodified JM objects (tests/golden/jm_tqc.npz, oracle/gen_golden_tqc.py)"""
    from oracle.gen_golden_tqc import CASES, NMB, chroma_blocks
    g = np.load(os.path.join(golden_dir, "jm_tqc.npz"))
    seen = set()
    for ci, (qp, sm, intra, uv, seed) in enumerate(CASES):
        orig, pred = chroma_blocks(NMB, seed)
        p = oracle.tq_params(g[f"c{ci}_params"], qp, mode=0, cavlc=int(sm == 0))
        got = oracle.tq_chroma(p, orig, pred)
        _chroma_cmp(got, [g[f"c{ci}_{n}"] for n in ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "cr_cbp")], (qp, sm, intra, uv))
        seen |= set(got[5].tolist())
    assert seen == {0, 1, 2}


@pytest.mark.skipif(not oracle.have_jmref(), reason="oracle/_ref/libjmref.so not built")
def test_restated_chroma_matches_reference_live():
    from oracle.gen_golden_tqc import chroma_blocks
    for qp, sm, intra, uv, seed in ((5, 0, 0, 0, 71), (20, 1, 1, 1, 72), (30, 0, 0, 1, 73), (38, 0, 1, 0, 74)):
        r = oracle.JMQuantRef(2 if intra else 0, sm)
        orig, pred = chroma_blocks(250, seed)
        p = oracle.tq_params(r.params_chroma(uv + 1, qp, intra), qp, cavlc=int(sm == 0))
        _chroma_cmp(oracle.tq_chroma(p, orig, pred), r.tq_chroma(qp, intra, uv, orig, pred), (qp, sm, intra, uv))
