"""GPU parity tests (through the C ABI): CUDA transform/quant/reconstruction vs golden vectors of the
unmodified JM objects and vs the restated oracle (bit-exact: integer path)."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu
NAMES = ("level", "run", "recon", "cost", "nz")


def test_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "jm_tq.npz"))
    for ci, (n, qp, intra, st, sm, seed, nblk) in enumerate(g["cases"]):
        n, qp = int(n), int(qp)
        orig, pred = synth.residual_blocks(int(nblk), n, int(seed))
        p = api.TQParams()
        q = oracle.tq_params(g[f"c{ci}_params"], qp, cavlc=int(sm == 0))
        for f, _ in api.TQParams._fields_:
            setattr(p, f, getattr(q, f))
        out = api.tq(p, orig, pred, n)
        for a, name in zip(out, NAMES):
            assert (a == g[f"c{ci}_{name}"]).all(), (ci, name)


@pytest.mark.parametrize("n,qp,intra,mode,field,dis,nblk", [
    (4, 28, 0, 0, 0, 0, 1), (4, 28, 0, 0, 0, 0, 130),            # ragged sizes (not a multiple of the CTA)
    (4, 33, 1, 0, 1, 1, 4097), (4, 28, 0, 1, 0, 0, 2000),        # field scan + disthres; version1 dct_luma mode
    (4, 10, 1, 1, 0, 0, 2000), (8, 28, 0, 0, 0, 0, 1), (8, 26, 1, 0, 1, 1, 1025),
    (4, 28, 0, 0, 0, 0, 130560), (8, 28, 0, 0, 0, 0, 32640),     # every 4x4 / 8x8 block of a 1080p luma frame
])
def test_matches_oracle(n, qp, intra, mode, field, dis, nblk):
    orig, pred = synth.residual_blocks(nblk, n, seed=nblk + qp)
    p = api.tq_default_params(n, qp, intra, mode=mode, field_scan=field, disthres=dis)
    q = oracle.tq_params(api.tq_params_table(p, n), qp, mode=mode, cavlc=1, field_scan=field, disthres=dis)
    got = api.tq(p, orig, pred, n)
    exp = oracle.tq(q, orig, pred, n)
    for a, b, name in zip(got, exp, NAMES):
        assert (a == b).all(), name


def test_empty_and_bad_arguments():
    p = api.tq_default_params(4, 28, 0)
    z = np.zeros((0, 16), np.uint8)
    out = api.tq(p, z, z, 4)
    assert out[0].shape == (0, 16)
    p.qp = 99
    with pytest.raises(api.B2Error):
        api.tq(p, np.zeros((1, 16), np.uint8), np.zeros((1, 16), np.uint8), 4)
    with pytest.raises(api.B2Error):
        api.tq_default_params(8, 28, 0, mode=1)        # version1 has no 8x8 transform


@pytest.mark.parametrize("qp,cavlc,field,seed", [(0, 1, 0, 41), (28, 1, 0, 42), (28, 0, 1, 43), (45, 1, 0, 44)])
def test_intra16x16_matches_oracle(qp, cavlc, field, seed):
    """b2tq_16x16 (k_tq16x16: DC Hadamard across the threads, DC / AC quantisers) vs the restated oracle, itself pinned to the
    unmodified residual_transform_quant_luma_16x16 (tests/test_oracle_tq.py)"""
    from oracle.gen_golden_tq16 import macroblocks
    orig, pred = macroblocks(1001, seed)
    p = api.tq_default_params(4, qp, 2, cavlc=cavlc, field_scan=field)
    q = oracle.tq_params(api.tq_params_table(p, 4), qp, mode=0, cavlc=cavlc, field_scan=field)
    got, exp = api.tq16x16(p, orig, pred), oracle.tq16x16(q, orig, pred)
    for a, b, name in zip(got, exp, ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "ac_coef")):
        assert (a == b).all(), (name, int((a != b).sum()))


@pytest.mark.parametrize("qp,cavlc,field,intra,seed", [(0, 1, 0, 1, 1), (14, 1, 1, 0, 2), (26, 0, 0, 1, 3), (33, 1, 0, 0, 4), (39, 1, 1, 1, 5)])
def test_chroma_matches_oracle(qp, cavlc, field, intra, seed):
    """b2tq_chroma (k_tq_chroma: hadamard2x2 through shuffles, DC / AC quantisers, the chroma coefficient-cost rule) vs the restated
    oracle, itself pinned to the unmodified residual_transform_quant_chroma_4x4 (tests/test_oracle_tq.py)"""
    from oracle.gen_golden_tqc import chroma_blocks
    orig, pred = chroma_blocks(2003, seed)
    p = api.tq_default_params(4, qp, intra, cavlc=cavlc, field_scan=field)
    q = oracle.tq_params(api.tq_params_table(p, 4), qp, mode=0, cavlc=cavlc, field_scan=field)
    got, exp = api.tq_chroma(p, orig, pred), oracle.tq_chroma(q, orig, pred)
    for a, b, name in zip(got, exp, ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "cr_cbp")):
        assert (a == b).all(), (name, int((a != b).sum()))
    assert set(got[5].tolist()) == {0, 1, 2} or qp == 0
