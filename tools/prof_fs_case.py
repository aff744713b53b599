"""Profiling target: one integer-search launch on a bench.robustness_cases() case (CASE=pan|perpart|off8|noise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from h264_b200 import api
case = os.environ.get("CASE", "off8")
name, fr, pred, cen = [c for c in bench.robustness_cases() if c[0] == case][0]
dev = torch.device("cuda", 0)
s = api.Searcher(bench.W, bench.H, bench.NREFS, bench.R)
d = torch.from_numpy(fr).to(dev)
s.set_cur_dev(d[bench.NREFS])
for r in range(bench.NREFS):
    s.set_ref_dev(r, d[bench.NREFS - 1 - r])
nmb = s.nmb
mvi = torch.zeros((nmb, bench.NREFS, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
ci = torch.zeros((nmb, bench.NREFS, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
p = api.make_params(bench.LAMBDA, do_subpel=False)
for it in range(int(os.environ.get("ITERS", 1))):
    s.search_frame_dev(torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev), p, mvi, ci, mvs, cs)
torch.cuda.synchronize()
print("ok", int(mvi.to(torch.int64).sum().item()), s.search_stats())
