"""k_sad_fs probes on the bench workload (resident, one launch per frame): kernel time, optional task repetition
(B2ME_FS_REP) and the warp-cycle split of a -DFS_PROFILE build (B2ME_FS_PROFILE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from h264_b200 import api
fr, pred, cen = bench.workload(seed=1)
dev = torch.device("cuda", 0)
s = api.Searcher(bench.W, bench.H, bench.NREFS, bench.R)
t = lambda a: torch.from_numpy(a).to(dev)
d_fr = t(fr)
s.set_cur_dev(d_fr[bench.NREFS])
for r in range(bench.NREFS):
    s.set_ref_dev(r, d_fr[bench.NREFS - 1 - r])
nmb = s.nmb
mvi = torch.zeros((nmb, bench.NREFS, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
ci = torch.zeros((nmb, bench.NREFS, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
p = api.make_params(bench.LAMBDA, do_subpel=False)
dp, dc = t(pred), t(cen)
for _ in range(3):
    s.search_frame_dev(dp, dc, p, mvi, ci, mvs, cs)
torch.cuda.synchronize()
s.search_stats()
s.kernel_timing(True)
for _ in range(5):
    s.search_frame_dev(dp, dc, p, mvi, ci, mvs, cs)
torch.cuda.synchronize()
ms, n = s.kernel_time_ms(0)
print(f"k_sad_fs {ms / n:.4f} ms/launch  REP={os.environ.get('B2ME_FS_REP', '0')}", flush=True)
print(s.search_stats())
print("checksum", int(mvi.to(torch.int64).sum().item()), int(ci.sum().item() & 0xffffffff), flush=True)
