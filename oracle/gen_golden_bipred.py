"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/jm_bipred.npz from the UNMODIFIED reference: outputs of
full_search_bipred_motion_estimation + sub_pel_bipred_motion_estimation (JM/lencod/src/me_fullsearch.c:112, :300, with
computeBiPredSAD/SSE/SATD 1 and 2 of me_distortion.c behind them) driven through oracle/jm_harness.c on seeded
synthetic frames: every (half, quarter)-pel metric pair, unweighted and weighted, 4x4 and 8x8 Hadamard.

Run in the build container (needs /root/reference):  python oracle/gen_golden_bipred.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from h264_b200 import synth  # noqa: E402

W, H, R, NR, LAM, DENOM = 96, 64, 8, 3, (187, 150, 120), 5
CASES = [(mh, mq, wp, t8) for (mh, mq) in ((2, 2), (0, 0), (1, 1), (0, 2)) for wp in (0, 1) for t8 in (0, 1)
         if not (wp and t8 and 2 in (mh, mq))]      # weighted 8x8 Hadamard = reference bug Q-J5, not restated


def frames():
    fr = synth.luma_sequence(W, H, NR + 1, seed=3)
    return fr[NR], fr[[2, 1, 0]]


def jobs_of(case_index, wp, t8):
    return synth.bipred_jobs(W, H, NR, R, 48, seed=100 + case_index, weighted=bool(wp), blocktypes=(1, 2, 3, 4) if t8 else (1, 2, 3, 4, 5, 6, 7))


if __name__ == "__main__":
    cur, refs = frames()
    out = {}
    for k, (mh, mq, wp, t8) in enumerate(CASES):
        jm = oracle.JMRef(W, H, R, NR, metric=(0, mh, mq))
        jm.set_cur(cur)
        for r in range(NR):
            jm.set_ref(r, refs[r])
        res = jm.bipred_search(jobs_of(k, wp, t8), LAM, test8x8=bool(t8), wp=bool(wp), log_denom=DENOM)
        out[f"c{k}"] = np.concatenate([res["cost_int"][:, None], res["cost_sub"][:, None], res["mv_int"], res["mv_sub"]], axis=1)
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "jm_bipred.npz"), cases=np.array(CASES), **out)
    print("wrote jm_bipred.npz:", len(CASES), "cases")
