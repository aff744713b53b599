"""Profiling target: one 1080p search step (resident inputs), small enough for ncu replays."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from h264_b200 import api
NR = int(os.environ.get("NR", bench.NREFS))
fr, pred, cen = bench.workload()
s = api.Searcher(bench.W, bench.H, bench.NREFS, bench.R)
s.set_cur(fr[bench.NREFS])
for r in range(bench.NREFS):
    s.set_ref(r, fr[bench.NREFS - 1 - r])
p = api.make_params(bench.LAMBDA)
for it in range(int(os.environ.get("ITERS", 2))):
    res = s.search_frame(pred, cen, p)
print("ok", int(res[2].astype(np.int64).sum()))
