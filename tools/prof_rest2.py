"""Profiling target for the kernels added in round 2 (tools/prof_rest.py covers round 1's): one launch each at 1080p so that one
`ncu --set full` pass captures k_mc_mb, k_tq16x16, k_tq_chroma, k_dbk_prep, k_deblock, k_epzs, k_sad_table, k_bicand, k_bid_cost,
k_expand_pred, k_select_refs, k_gather_best."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from h264_b200 import api, synth

dev = torch.device("cuda", 0)
# k_mc_mb, k_tq4x4, k_tq_chroma, k_deblock: the reconstruction leg of the bench (3 warm-up passes + 1)
print(bench.recon_leg(0, steps=1))

W, H, NR, R = bench.W, bench.H, 2, 32
rng = np.random.default_rng(0)
fr = synth.luma_sequence(W, H, NR + 1, seed=1)
s = api.Searcher(W, H, NR, R)
s.set_cur(fr[NR])
for r in range(NR):
    s.set_ref(r, fr[NR - 1 - r])
nmb = s.nmb
# k_tq16x16: Intra16x16 luma of every macroblock
o16 = rng.integers(0, 256, (nmb, 256), dtype=np.uint8)
p16 = np.clip(o16.astype(int) + rng.integers(-12, 13, o16.shape), 0, 255).astype(np.uint8)
api.tq16x16(api.tq_default_params(4, 28, 1), o16, p16)
# k_expand_pred, k_sad_fs, k_subpel_refine, k_select_refs, k_gather_best: the compact end-to-end entry
pred_mb = rng.integers(-8, 9, (nmb, NR, 2)).astype(np.int16)
s.search_frame_best(pred_mb, api.make_params(bench.LAMBDA), bench.LAMBDA[0])
# k_epzs: every partition of every macroblock, five predictors each
geom = np.array(synth.PART_GEOM, np.int64)
mbw = W // 16
mbs = np.arange(nmb)
nj = nmb * 41
jobs = np.zeros(nj, synth.EPZS_JOB)
mbx, mby = np.repeat(mbs % mbw, 41), np.repeat(mbs // mbw, 41)
pp = np.tile(np.arange(41), nmb)
jobs["pos_x"] = 16 * mbx + geom[pp, 1]; jobs["pos_y"] = 16 * mby + geom[pp, 2]; jobs["blocktype"] = geom[pp, 0]
jobs["mv"] = (8, 4); jobs["pred"] = (8, 4); jobs["range"] = (4 * R, 4 * R)
jobs["mv_range"] = np.where(geom[pp, 0] < 5, 10, 12)
jobs["flags"] = 4 | 8 | np.where(geom[pp, 0] > 4, 2, 0)
jobs["lambda_factor"] = bench.LAMBDA[0]
jobs["medthres"] = (geom[pp, 3] * geom[pp, 4]) << 5
jobs["stop0"] = jobs["medthres"] + 2 * bench.LAMBDA[0]; jobs["stop"] = 2 * jobs["medthres"] + 2 * bench.LAMBDA[0]
jobs["pred_first"] = 5 * np.arange(nj); jobs["npred"][:, 0] = 5; jobs["cond_host"][:, 0] = 1
jobs["pat_init"] = 2; jobs["pat_sd"] = 0; jobs["pat_sq"] = 1; jobs["pat_else"] = 2; jobs["pat_dual"] = 2
preds = rng.integers(-16, 17, (nj * 5, 2)).astype(np.int16)
res = s.epzs_search(jobs, preds, synth.epzs_patterns())
# k_sad_table: setup_fast_full_search's SAD tables of a few macroblocks; k_bicand: bi-predictive distortions at explicit candidates
import ctypes as C
tab = np.zeros((41, (2 * R + 1) ** 2), np.uint16)
for mbx, mby in ((3, 2), (60, 30), (119, 67)):
    assert s.L.b2me_sad_table(s.h, mbx, mby, 0, (C.c_int16 * 2)(8, 4), R, tab.ctypes.data_as(C.c_void_p)) == 0
bj = synth.bipred_jobs(W, H, NR, 16, 4096, seed=2)
outb = np.zeros(len(bj), np.int64)
assert s.L.b2me_bipred_distortion_candidates(s.h, 2, 0, 0, 0, len(bj), bj.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p)) == 0
# k_bid_cost: BIDPartitionCost of 16 K partitions
bid = synth.bid_jobs(W, H, NR, 16384, seed=4)
cb = s.bid_partition_cost(bid, 2)
print("ok", int(cb.sum() & 0xffff), int(res["npoints"].sum()), int(tab.sum()), int(outb.sum() & 0xffff))
