"""F7 pinned by execution ("V1 dct_luma parity unpinned"): the restated oracle in its version1 mode
(oracle/b2_oracle_tq.c, mode 1) and the product's default version1 parameter table against the UNMODIFIED dct_luma
(V1/src/block.c:836-1045) -- golden vectors (oracle/gen_golden_v1_tq.py) and, where oracle/_ref/libv1tq.so exists, live."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth
from oracle.gen_golden_v1_tq import CASES, NBLK


def _restated(qp, st, orig, pred):
    p = api.tq_default_params(4, qp, 2 if st == 2 else 0, mode=1)       # pure host code of the product: quant_coef, (1 << q_bits) / 3, dequant_coef
    return oracle.tq(oracle.tq_params(api.tq_params_table(p, 4), qp, mode=1, cavlc=0), orig, pred, 4)


def test_restated_dct_luma_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "v1_dct_luma.npz"))
    assert [tuple(c) for c in g["cases"]] == CASES
    for ci, (qp, st, seed) in enumerate(CASES):
        orig, pred = synth.residual_blocks(NBLK, 4, seed)
        got = _restated(qp, st, orig, pred)
        for a, name in zip(got, ("level", "run", "recon", "cost", "nz")):
            assert (np.asarray(a) == g[f"c{ci}_{name}"]).all(), (qp, st, name)


@pytest.mark.skipif(not oracle.have_v1tq(), reason="oracle/_ref/libv1tq.so (the unmodified version1 block.c) is not built here")
def test_restated_dct_luma_matches_reference_live():
    for qp, st, seed in ((5, 0, 201), (19, 2, 202), (33, 0, 203), (48, 0, 204)):
        orig, pred = synth.residual_blocks(400, 4, seed)
        for a, b in zip(_restated(qp, st, orig, pred), oracle.v1_dct_luma(qp, st, orig, pred)):
            assert (np.asarray(a) == np.asarray(b)).all(), (qp, st)
