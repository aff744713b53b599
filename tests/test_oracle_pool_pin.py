"""The per-pair fit of the pool-mode fractal matching pinned to the REFERENCE: version1 has no pool search, but
the fit of every (range block under an isometry, decimated domain block) pair is its compute_rms (V1/src/compute.c:6-189).
tests/golden/v1_pool_fit.npz holds 4096 such pairs laid into the reference's own planes and what the UNMODIFIED
compute_range_Sum / compute_domain_Sum / compute_rms returned for them (oracle/gen_golden_pool_fit.py).  The oracle's exact
integer fit (oracle/b2_oracle_pool.c orc_pool_pair -- what k_frac_pool's epilogue computes, tests/test_gpu_pool.py) must
agree: alpha * 100 and beta EXACTLY, accept / reject exactly, err_num / 640000 == rms to double rounding."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle


def _check(g):
    r, d, a100, beta, rms = g["r"], g["d"], g["a100"], g["beta"], g["rms"]
    n_rej = 0
    for k in range(len(r)):
        ok, aq, G, bq, err = oracle.pool_pair(r[k], d[k])
        assert aq == a100[k], (k, aq, a100[k])                       # (int)(alpha * 100) then QUAN_A: exact
        assert bq == beta[k], (k, bq, beta[k])                       # QUAN_A(rsum1 / no): exact
        assert ok == (rms[k] < 1e29), (k, ok, rms[k])                # MIN_ALPHA / MAX_ALPHA reject
        if ok:
            assert abs(err / 640000.0 - rms[k]) <= 1e-9 * max(1.0, abs(rms[k])) + 1e-7, (k, err / 640000.0, rms[k])
        else:
            n_rej += 1
    return n_rej


def test_exact_fit_equals_unmodified_compute_rms_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "v1_pool_fit.npz"))
    assert len(g["r"]) == 4096
    n_rej = _check(g)
    assert n_rej > 100 and (g["a100"] < 0).sum() > 100 and (g["a100"] == 0).sum() > 100     # rejects, negatives, det == 0 are all there


@pytest.mark.skipif(not oracle.have_v1ref(), reason="oracle/_ref/libv1ref.so (the unmodified version1 objects) is not built here")
def test_exact_fit_equals_unmodified_compute_rms_live(tmp_path):
    """fresh pairs through the reference objects in a subprocess (the reference keeps its state in globals)"""
    out = tmp_path / "fit.npz"
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import gen_golden_pool_fit as g; "
            "r, d = g.make_pairs(seed=77, nframes=4); a, b, e = g.reference_fit(r, d); np.savez(%r, r=r, d=d, a100=a, beta=b, rms=e)"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(out)))
    subprocess.check_call([sys.executable, "-c", code])
    _check(np.load(out))
