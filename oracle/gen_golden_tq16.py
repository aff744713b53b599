"""TEST INFRASTRUCTURE ONLY.  tests/golden/jm_tq16.npz: the UNMODIFIED JM residual_transform_quant_luma_16x16
(JM/lencod/src/block.c:207-345; oracle/_ref/libjmref.so through oracle/jm_harness_tq.c) on seeded Intra16x16 macroblocks:
DC / AC level and run lists, reconstruction, ac_coef, at several QPs, CAVLC and CABAC.
Run in the build container (needs /root/reference):  python oracle/gen_golden_tq16.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
CASES = [(0, 0, 31), (12, 0, 32), (28, 0, 33), (28, 1, 34), (40, 0, 35), (51, 1, 36)]      # (qp, symbol_mode, seed)
NMB = 300


def macroblocks(nmb, seed):
    """(orig, pred) [nmb][256]: DC prediction errors, smooth ramps, noise, flat (all-zero residual), saturated macroblocks"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:16, 0:16]
    base = rng.integers(0, 256, (nmb, 1, 1))
    ramp = yy[None] * rng.integers(-6, 7, (nmb, 1, 1)) + xx[None] * rng.integers(-6, 7, (nmb, 1, 1))
    amp = rng.choice([0, 1, 3, 8, 30, 255], (nmb, 1, 1))
    orig = np.clip(base + ramp + rng.integers(-1, 2, (nmb, 16, 16)) * amp, 0, 255)
    pred = np.clip(base + rng.integers(-40, 41, (nmb, 1, 1)) * rng.choice([0, 1], (nmb, 1, 1)), 0, 255) * np.ones((1, 16, 16), int)
    pred[::11] = orig[::11]
    orig[::13] = 255; pred[::13] = 0
    return orig.reshape(nmb, 256).astype(np.uint8), pred.reshape(nmb, 256).astype(np.uint8)


if __name__ == "__main__":
    import oracle
    out = {"cases": np.array(CASES, np.int32)}
    for ci, (qp, sm, seed) in enumerate(CASES):
        r = oracle.JMQuantRef(2, sm)
        orig, pred = macroblocks(NMB, seed)
        res = r.tq16x16(qp, orig, pred)
        for a, name in zip(res, ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "ac_coef")):
            out[f"c{ci}_{name}"] = a
        out[f"c{ci}_params"] = r.params(4, qp, 1)
        print(f"qp {qp} symbol_mode {sm}: {int((res[5] != 0).sum())} of {NMB} with AC, {int((res[0][:, 0] != 0).sum())} with DC")
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "jm_tq16.npz"), **out)
