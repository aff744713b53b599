"""Profiling target for every kernel that bench.py's step does not dominate: one launch each at a realistic size
(1080p pictures / CIF fractal search) so that one `ncu --set full` pass can capture them (tools/gpu_round.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from h264_b200 import api, synth

dev = torch.device("cuda", 0)
W, H, NR, R = bench.W, bench.H, 2, 32
fr = synth.luma_sequence(W, H, NR + 1, seed=1)
s = api.Searcher(W, H, NR, R)
s.set_cur(fr[NR])
s.set_ref_weights(1, 40, -3, 5)                       # k_apply_wp on slot 1
for r in range(NR):
    s.set_ref(r, fr[NR - 1 - r])                      # k_half_planes, k_quarter_planes, k_search_plane
nmb = s.nmb
# k_mc_luma -> k_tq4x4 / k_tq8x8 (search -> prediction -> transform chain layout)
rng = np.random.default_rng(0)
mv = torch.from_numpy(rng.integers(-40, 41, (nmb, NR, 41, 2)).astype(np.int16)).to(dev)
mb_mode = torch.from_numpy(rng.choice([1, 2, 3, 8], nmb).astype(np.uint8)).to(dev)
b8mode = torch.from_numpy(rng.integers(4, 8, (nmb, 4)).astype(np.uint8)).to(dev)
ref8 = torch.from_numpy(rng.integers(0, NR, (nmb, 4)).astype(np.int8)).to(dev)
orig = torch.zeros((nmb * 16, 16), dtype=torch.uint8, device=dev); pred = torch.zeros_like(orig)
s.mc_luma_dev(mb_mode, b8mode, ref8, mv, orig, pred)
for n in (4, 8):
    nb = nmb * 16 if n == 4 else nmb * 4
    o, p = orig.view(nb, n * n), pred.view(nb, n * n)
    level = torch.zeros((nb, n * n), dtype=torch.int16, device=dev); run = torch.zeros((nb, n * n), dtype=torch.uint8, device=dev)
    recon = torch.zeros_like(run); cost = torch.zeros(nb, dtype=torch.int32, device=dev); nz = torch.zeros(nb, dtype=torch.uint8, device=dev)
    api.tq_dev(api.tq_default_params(n, 28, 0), o, p, n, level, run, recon, cost, nz)
torch.cuda.synchronize()
# k_distortion (mode-decision distortions on difference blocks)
diff = rng.integers(-255, 256, (nmb * 16, 16)).astype(np.int16)
api.distortion_blocks(2, 4, diff)
api.distortion_blocks(2, 8, diff.reshape(-1, 64))
# k_bipred, k_cand_dist
s.bipred_search(synth.bipred_jobs(W, H, NR, 16, 2048, seed=2), api.make_params(bench.LAMBDA))
s.distortion_candidates(synth.candidates(W, H, NR, 200000, seed=3), 2)
# fractal window search (BASELINE config 2: CIF, +-7, all levels) : k_frac_domain_sums, k_frac_range_sums, k_frac_window
(ry, ru, rv), (cy, cu, cv) = synth.yuv_pair(352, 288, seed=3, shift=(-3, 2), gain=0.8, offset=12.0)
f = api.FractalSearcher(352, 288, 7)
f.set_domain(0, ry, ru, rv)
f.set_range(cy, cu, cv)
out = f.search_plane(0, 1)
nodes = f.encode_plane(1, (3.5, 4.5, 2.0))            # k_frac_decide
rec = f.decode_plane(1)                               # k_frac_predict
print("ok", float(np.asarray(out[2]).min()), int(rec.sum()))
