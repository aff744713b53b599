"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/jm_tq.npz from the UNMODIFIED JM objects
(oracle/_ref/libjmref.so, oracle/jm_harness_tq.c): outputs of residual_transform_quant_luma_4x4 /
_8x8 (ACLevel/ACRun lists, reconstruction, coeff_cost increment, return value) on seeded residual
blocks (h264_b200.synth.residual_blocks), and the LevelQuantParams tables the reference derives.
Run in the build container (needs /root/reference):  python oracle/gen_golden_tq.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from h264_b200 import synth  # noqa: E402

CASES = [  # (n, qp, intra, slice_type, symbol_mode, seed, nblk)
    (4, 28, 0, 0, 0, 1, 512), (4, 28, 1, 0, 0, 2, 512), (4, 0, 0, 0, 0, 3, 512), (4, 51, 1, 2, 1, 4, 512), (4, 17, 0, 0, 1, 5, 512),
    (8, 28, 0, 0, 1, 6, 256), (8, 28, 1, 2, 1, 7, 256), (8, 3, 0, 0, 1, 8, 256), (8, 44, 0, 0, 1, 9, 256),
]

if __name__ == "__main__":
    out = {"cases": np.array(CASES, np.int64)}
    refs = {}
    for ci, (n, qp, intra, st, sm, seed, nblk) in enumerate(CASES):
        r = refs.setdefault((st, sm), oracle.JMQuantRef(st, sm))
        orig, pred = synth.residual_blocks(nblk, n, seed)
        lv, rn, rec, cost, nz = r.tq(n, qp, intra, orig, pred)
        out[f"c{ci}_params"] = r.params(n, qp, intra)
        out[f"c{ci}_level"], out[f"c{ci}_run"], out[f"c{ci}_recon"], out[f"c{ci}_cost"], out[f"c{ci}_nz"] = lv, rn, rec, cost, nz
        print(ci, n, qp, "nonzero blocks", int(nz.sum()), "max |level|", int(np.abs(lv).max()))
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "jm_tq.npz"), **out)
