import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from h264_b200 import api
fr, pred, cen = bench.workload(seed=1)
s = api.Searcher(bench.W, bench.H, bench.NREFS, bench.R)
s.set_cur(fr[bench.NREFS])
for r in range(bench.NREFS):
    s.set_ref(r, fr[bench.NREFS - 1 - r])
p = api.make_params(bench.LAMBDA, do_subpel=1)
mvi, ci, mvs, cs = s.search_frame(pred, cen, p)
def uni(m):
    m = m.reshape(-1, 41, 2)
    return float(((m == m[:, :1]).all(-1).all(-1)).mean())
print("uniform int", uni(mvi), "uniform sub", uni(mvs))
m = mvi.reshape(-1, 41, 2)
nd = np.array([len({(int(a), int(b)) for a, b in x}) for x in m[:4000]])
print("distinct int mv per item: mean", nd.mean(), "hist", np.bincount(nd)[:12])
