"""TEST INFRASTRUCTURE ONLY.  tests/golden/v1_dct_luma.npz: what the UNMODIFIED version1 dct_luma (V1/src/block.c:836-1045,
oracle/_ref/libv1tq.so, recipe in oracle/Makefile.v1) returns for seeded residual blocks (h264_b200.synth.residual_blocks)
at several QPs and both slice types: levels, runs, reconstruction, coefficient cost, nonzero flag.
Run in the build container (needs /root/reference):  python oracle/gen_golden_v1_tq.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
CASES = [(0, 2, 11), (10, 0, 12), (22, 0, 13), (28, 2, 14), (28, 0, 15), (37, 0, 16), (45, 2, 17), (51, 0, 18)]   # (qp, slice type (0 P, 2 I), seed)
NBLK = 600

if __name__ == "__main__":
    import oracle
    from h264_b200 import synth
    out = {"cases": np.array(CASES, np.int32)}
    for ci, (qp, st, seed) in enumerate(CASES):
        orig, pred = synth.residual_blocks(NBLK, 4, seed)
        lv, rn, rec, cost, nz = oracle.v1_dct_luma(qp, st, orig, pred)
        out[f"c{ci}_level"], out[f"c{ci}_run"], out[f"c{ci}_recon"], out[f"c{ci}_cost"], out[f"c{ci}_nz"] = lv, rn, rec, cost, nz
        print(f"qp {qp} type {st}: {int(nz.sum())} of {NBLK} blocks nonzero, max |level| {int(np.abs(lv).max())}")
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "v1_dct_luma.npz"), **out)
