"""BASELINE config 4: ONE 4K picture, full search +-64 (1 reference), MB-row bands across the ranks with the
reconstructed-reference halo exchange over NCCL before every picture.
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/band_bench.py [--steps K]
Prints one JSON line on rank 0: whole-picture Mpel-search-points/s (device time, max over ranks), the share of
the halo exchange, and the bytes exchanged."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from h264_b200 import api, bands, synth

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--range", type=int, default=64)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29531",
                        rank=rank, world_size=world)
W, H, R = args.width, args.height // 16 * 16, args.range
fr = synth.luma_sequence(W, H, 2, seed=9)
nmb = (W // 16) * (H // 16)
pred, cen = synth.predictors(W, H, 1, seed=2, spread=0, base=np.tile(np.array([[[[8, 4]]]], np.int64), (nmb, 1, 1, 1)))
b = bands.BandSearcher(W, H, 1, R, rank, world, device=local, max_center_pel=16)
d_cur = torch.from_numpy(fr[1]).to(dev)
own = torch.from_numpy(fr[0][16 * b.first_row:16 * b.last_row]).to(dev).contiguous()
d_pred, d_cen = torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev)
mvi = torch.zeros((nmb, 1, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
ci = torch.zeros((nmb, 1, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
params = api.make_params((187, 187, 187))
stream = torch.cuda.current_stream().cuda_stream
b.set_cur_dev(d_cur, stream)
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)


def step():
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    b.set_ref_from_band(0, own, stream)          # halo exchange (NCCL p2p) + planes of the reference
    e1.record()
    b.search(d_pred, d_cen, params, mvi, ci, mvs, cs, stream)
    e2.record()
    return e0, e1, e2


for _ in range(max(3, args.warmup)):
    step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
evs = []
for i in range(args.steps):
    flush.fill_(i)
    evs.append(step())
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ms = sum(e[0].elapsed_time(e[2]) for e in evs) / args.steps
ms_x = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
t = torch.tensor([ms, ms_x], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
halo = sum((hi - lo) * W for s_, d_, lo, hi in bands.exchange_plan(world, H // 16, R) if d_ == rank)
hb = torch.tensor([halo], dtype=torch.int64, device=dev); dist.all_reduce(hb)
if rank == 0:
    pel = nmb * (2 * R + 1) ** 2 * 256
    print(json.dumps({"metric": "Mpel-search-points/s", "value": pel / (float(t[0]) * 1e-3) / 1e6, "n_gpus": world, "ms_per_step": float(t[0]),
                      "ms_halo_exchange_and_planes": float(t[1]), "halo_bytes_all_ranks": int(hb.item()),
                      "config": {"workload": f"{W}x{H} single picture, full search +-{R}, 1 ref, {world} MB-row bands, NCCL halo exchange per picture",
                                 "l2": "192 MB L2 flush between timed iterations"},
                      "mb_per_s": nmb / (float(t[0]) * 1e-3), "scaling": "strong", "checksum": int(mvs[b.mb_first:b.mb_first + b.mb_count].to(torch.int64).sum().item())}))
dist.destroy_process_group()
