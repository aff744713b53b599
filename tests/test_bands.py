"""Band partition + halo exchange (h264_b200/bands.py): plan logic and a world_size-2/3 gloo run on CPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from h264_b200 import bands


def test_band_rows_cover_the_picture():
    for mbh in (9, 17, 68, 135):
        for world in (1, 2, 3, 4, 8):
            rows = [bands.band_mb_rows(r, world, mbh) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == mbh
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1


def test_exchange_plan_is_exactly_the_needed_rows():
    for mbh, world, R in ((135, 8, 64), (68, 4, 32), (9, 3, 16), (9, 8, 64)):      # last: halo spans several bands
        plan = bands.exchange_plan(world, mbh, R)
        for dst in range(world):
            lo, hi = bands.needed_rows(dst, world, mbh, R)
            f, l = bands.band_mb_rows(dst, world, mbh)
            got = np.zeros(16 * mbh, int)
            got[16 * f:16 * l] += 1
            for s, d, a, b in plan:
                if d == dst:
                    sf, sl = bands.band_mb_rows(s, world, mbh)
                    assert 16 * sf <= a < b <= 16 * sl          # the source owns what it sends
                    got[a:b] += 1
            assert (got[lo:hi] == 1).all() and got[:lo].sum() == 0 and got[hi:].sum() == 0


def _worker(rank, world, port, H, W, R, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        frame = torch.from_numpy(rng.integers(0, 256, (H, W), dtype=np.uint8))
        f, l = bands.band_mb_rows(rank, world, H // 16)
        full = bands.exchange_halos(frame[16 * f:16 * l].clone(), H, W, R)
        lo, hi = bands.needed_rows(rank, world, H // 16, R)
        ok = bool((full[lo:hi] == frame[lo:hi]).all()) and int(full[:lo].sum()) == 0 and int(full[hi:].sum()) == 0
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,R", [(2, 144, 16), (3, 272, 64)])
def test_halo_exchange_gloo(world, H, R):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, 64, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


# ---- fractal pool matching across ranks (h264_b200/pool_bands.py): range-row bands, replicated domain plane ----
def test_range_bands_cover_the_picture():
    from h264_b200 import pool_bands
    for rows8 in (6, 36, 135, 270):
        for world in (1, 2, 3, 4, 8):
            rows = [pool_bands.range_band_rows(r, world, rows8) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == rows8
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rows) - min(b - a for a, b in rows) <= 1


def _pool_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from h264_b200 import pool_bands
        rng = np.random.default_rng(11)
        plane = torch.from_numpy(rng.integers(0, 256, (48, 64), dtype=np.uint8))
        mine = plane.clone() if rank == 1 else torch.zeros_like(plane)       # rank 1 owns the reconstructed picture
        pool_bands.replicate_domain(mine, src=1)
        q.put((rank, bool((mine == plane).all())))
    finally:
        dist.destroy_process_group()


def test_domain_plane_is_replicated_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_pool_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
