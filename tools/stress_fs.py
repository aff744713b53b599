"""Determinism stress of the integer search: the same 1080p frame searched repeatedly, every result compared
with the first one (a timing-dependent difference means a race in k_sad_fs's pipeline)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from h264_b200 import api, synth

W, H, R, NR = bench.W, bench.H, bench.R, bench.NREFS
fr, pred, cen = bench.workload(seed=1)
rng = np.random.default_rng(11)
pred = pred.copy(); cen = cen.copy()
pj, cj = synth.predictors(W, H, NR, seed=5, spread=9, rmax=10)
sel = rng.random(pred.shape[0]) < 0.33
pred[sel] = pj[sel]; cen[sel] = cj[sel]
s = api.Searcher(W, H, NR, R)
s.set_cur(fr[NR])
for r in range(NR):
    s.set_ref(r, fr[NR - 1 - r])
p = api.make_params(bench.LAMBDA, do_subpel=int(os.environ.get("SUBPEL", "0")))
first = s.search_frame(pred, cen, p)
bad = 0
for it in range(int(os.environ.get("N", 40))):
    got = s.search_frame(pred, cen, p)
    d = (got[1] != first[1]) | (got[0] != first[0]).any(-1)
    if p.do_subpel:
        d |= (got[3] != first[3]) | (got[2] != first[2]).any(-1)
    if d.any():
        bad += 1
        idx = np.argwhere(d)
        print(f"iter {it}: {len(idx)} differences; first (mb, ref, p) = {idx[:6].tolist()}; jittered MB: {[bool(sel[i[0]]) for i in idx[:6]]}; "
              f"cost {got[1][tuple(idx[0])]} vs {first[1][tuple(idx[0])]}", flush=True)
print("iterations with differences:", bad)
