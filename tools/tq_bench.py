"""k_tq4x4 / k_tq8x8 device time on one 1080p picture's blocks and on 1 M blocks (HBM roofline of the transform path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from h264_b200 import api
dev = torch.device("cuda", 0)
for nmb in (8160, 65536):
    for n in (4, 8):
        nb = nmb * (16 if n == 4 else 4)
        g = torch.Generator(device=dev); g.manual_seed(1)
        o = torch.randint(0, 256, (nb, n * n), dtype=torch.uint8, device=dev, generator=g)
        p = (o.to(torch.int16) + torch.randint(-12, 13, (nb, n * n), dtype=torch.int16, device=dev, generator=g)).clamp(0, 255).to(torch.uint8)
        level = torch.zeros((nb, n * n), dtype=torch.int16, device=dev); run = torch.zeros((nb, n * n), dtype=torch.uint8, device=dev)
        recon = torch.zeros_like(run); cost = torch.zeros(nb, dtype=torch.int32, device=dev); nz = torch.zeros(nb, dtype=torch.uint8, device=dev)
        prm = api.tq_default_params(n, 28, 0)
        for _ in range(3):
            api.tq_dev(prm, o, p, n, level, run, recon, cost, nz)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            api.tq_dev(prm, o, p, n, level, run, recon, cost, nz)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        byt = nb * (2 * n * n + 2 * n * n + n * n + n * n + 5)       # orig + pred in; level, run, recon, cost, nonzero out
        print(f"tq{n}x{n} {nb} blocks: {us:.1f} us  {byt / us / 1e3:.0f} GB/s  nonzero {int(nz.sum())} level-checksum {int(level.to(torch.int64).abs().sum())}", flush=True)
