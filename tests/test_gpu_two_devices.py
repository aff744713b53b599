"""One process driving two GPUs (VERDICT r1 weak-7 / ADVICE r1): the >48 KB dynamic-shared-memory opt-in of k_sad_fs and
k_frac_pool is a per-device function attribute; contexts on two devices in one process must both launch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_contexts_on_two_devices_in_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    import oracle
    from h264_b200 import api, synth
    W, H, R, NR = 64, 48, 16, 2
    fr = synth.luma_sequence(W, H, NR + 1, seed=1)
    cur, refs = fr[NR], fr[[1, 0]]
    pred, cen = synth.predictors(W, H, NR, seed=2, spread=2, rmax=6)
    lam = (187, 187, 187)
    exp = oracle.OrcFrame(cur, refs, R).search_frame(pred, cen, lam)
    (yr, _, _), (yc, _, _) = synth.yuv_pair(64, 48, seed=3, shift=(-3, 2), gain=0.8, offset=12.0)
    ep = oracle.pool_search(yc, yr, 200)
    for dev in (0, 1):                       # device 1 second: its launch needs its own opt-in
        s = api.Searcher(W, H, NR, R, device=dev)
        s.set_cur(cur)
        for r in range(NR):
            s.set_ref(r, refs[r])
        got = s.search_frame(pred, cen, api.make_params(lam))
        for a, b in zip(got, exp):
            assert (a == b).all()
        ps = api.PoolSearcher(64, 48, 64, 48, 200, device=dev)
        ps.set_planes(yc, yr)
        for a, b in zip(ps.search(), ep):
            assert (a == b).all()
        s.close(); ps.close()
