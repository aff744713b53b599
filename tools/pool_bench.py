"""Pool-matching sweep (BASELINE config 5): 1080p range plane (32 400 8x8 ranges x 8 isometries) against
domain pools of 1K..64K blocks; kernel time (CUDA events), tensor-pipe utilisation, filter statistics."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from h264_b200 import api, synth

W, H = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080) // 8 * 8)
fr = synth.luma_sequence(W, H, 2, seed=3)
rp, dp = fr[1], fr[0]
out = []
for nd in [int(x) for x in os.environ.get("ND", "1024,4096,16384,65536").split(",")]:
    s = api.PoolSearcher(W, H, W, H, nd)
    s.set_planes(rp, dp)
    s.search(); s.kernel_time_ms(); s.stats()
    for _ in range(3):
        res = s.search()
    ms, n = s.kernel_time_ms()
    st = s.stats()
    pms = min(s.probe_ms() for _ in range(3))
    nr = s.nr
    ops = 2.0 * 8 * nr * nd * 64
    rec = {"pool": nd, "ranges": nr, "kernel_ms": ms / n, "tensor_only_ms": pms, "Tops": ops / (ms / n * 1e-3) / 1e12,
           "tensor_only_Tops": ops / (pms * 1e-3) / 1e12, "pairs_per_s": 8.0 * nr * nd / (ms / n * 1e-3),
           "exact_per_row": st["exact_evals"] / (3 * 8 * nr), "rescan_frac": st["chunk_rescans"] / max(1, st["chunks"]),
           "rejected": int((res[0] < 0).sum())}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    s.close()
json.dump(out, open(os.path.join("gpurun_out", "pool_bench.json"), "w"))
