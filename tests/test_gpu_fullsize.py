"""GPU parity at BASELINE.json's full sizes: the whole workload runs on the GPU, a seeded sample of it is checked
bit-exactly against the oracle (which finishes the sample in seconds), and size-independent properties are checked
on everything."""
import numpy as np
import pytest

import bench
import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_1080p_4refs_sample_matches_oracle():
    """BASELINE config 3 (bench.py's workload): 1080p, +-32, 4 refs, 41 partitions, SAD + SATD sub-pel."""
    W, H, R, NR = bench.W, bench.H, bench.R, bench.NREFS
    fr, pred, cen = bench.workload(seed=1)
    rng = np.random.default_rng(11)
    pred = pred.copy(); cen = cen.copy()
    # a third of the macroblocks get per-partition predictors (several centre groups), the rest share one
    pj, cj = synth.predictors(W, H, NR, seed=5, spread=9, rmax=10)
    sel = rng.random(pred.shape[0]) < 0.33
    pred[sel] = pj[sel]; cen[sel] = cj[sel]
    cur, refs = fr[NR], fr[[NR - 1 - r for r in range(NR)]]
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    got = s.search_frame(pred, cen, api.make_params(bench.LAMBDA))
    mv_int, cost_int, mv_sub, cost_sub = got
    # ---- properties on the whole frame ----
    nmb = (W // 16) * (H // 16)
    assert mv_int.shape == (nmb, NR, 41, 2)
    assert (cost_int > 0).all() and (cost_sub > 0).all()
    d = (mv_int.astype(np.int32) - cen.astype(np.int32))
    assert (np.abs(d) <= 4 * R).all() and (d % 4 == 0).all()                      # integer vectors inside the window
    assert (np.abs(mv_sub.astype(np.int32) - mv_int.astype(np.int32)) <= 3).all()  # +-2 then +-1 quarter-pel
    again = s.search_frame(pred, cen, api.make_params(bench.LAMBDA))                # idempotent / deterministic
    for a, b in zip(got, again):
        assert (a == b).all()
    # ---- bit-exact sample: corners, borders, interior, the jittered macroblocks ----
    of = oracle.OrcFrame(cur, refs, R)
    mbw = W // 16
    sample = sorted(set([0, mbw - 1, nmb - mbw, nmb - 1, 5 * mbw, 6 * mbw - 1] + list(rng.integers(0, nmb, 30)) + list(np.flatnonzero(sel)[:8])))
    for mb in sample:
        exp = of.search_frame(pred, cen, bench.LAMBDA, mb_first=int(mb), mb_count=1)
        for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
            assert (a[mb] == b[mb]).all(), (n, int(mb))


@pytest.mark.parametrize("nd", [16384, 65536])
def test_pool_1080p_sample_matches_oracle(nd):
    """BASELINE config 5: 1080p range plane against a 16 K / 64 K pool: sampled range blocks re-derived by brute force."""
    W, H = 1920, 1080
    fr = synth.luma_sequence(W, H, 2, seed=3)
    rp, dp = fr[1], fr[0]
    s = api.PoolSearcher(W, H, W, H, nd)
    s.set_planes(rp, dp)
    dom, iso, aq, beta, err = s.search()
    assert (dom >= -1).all() and (dom < nd).all() and (iso < 8).all()
    assert ((aq >= -235) & (aq <= 400)).all() and (err[dom >= 0] >= 0).all()
    xy = oracle.pool_positions(W, H, nd)
    pool = np.stack([oracle.pool_domain_block(dp, *p) for p in xy]).astype(np.int64)
    Sd, Sd2 = pool.sum(1), (pool * pool).sum(1)
    det = 64 * Sd2 - Sd * Sd
    rng = np.random.default_rng(2)

    def quan(a):
        c = np.trunc(a / 10).astype(np.int64); b = a - c * 10
        return np.where((b > 2) & (b < 8), c * 10 + 5, np.where(b > 7, (c + 1) * 10, c * 10))

    for ri in rng.integers(0, s.nr, 24 if nd <= 16384 else 10):
        bx, by = ri % (W // 8), ri // (W // 8)
        r0 = rp[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8].reshape(64)
        best = (-1, 0, 0, 0)
        for k in range(8):
            r = oracle.pool_iso(r0, k).astype(np.int64)
            num = 64 * (pool @ r) - r.sum() * Sd
            a = np.where(det == 0, 0, np.trunc(100 * num / np.maximum(det, 1))).astype(np.int64)   # |100 num| < 2^53: exact in float64
            q = quan(a)
            G = np.where((q >= -235) & (q <= 400), 200 * q * num - q * q * det, -1)
            j = int(np.argmax(G))                                                 # first maximum = lowest pool index
            if G[j] > best[0]:
                best = (int(G[j]), j, k, int(q[j]))
        assert (best[1], best[2], best[3]) == (dom[ri], iso[ri], aq[ri]), int(ri)
        Sr, Sr2 = int(r0.astype(np.int64).sum()), int((r0.astype(np.int64) ** 2).sum())
        assert err[ri] == 640000 * (Sr2 - 2 * int(beta[ri]) * Sr + 64 * int(beta[ri]) ** 2) - best[0]


def test_4k_range64_sample_matches_oracle():
    """BASELINE config 4's shape on one GPU: 3840x2160 (coded x2160), full search +-64, 1 reference.  Sampled macroblocks --
    picture corners / borders and the first / last MB rows of the 8 bands of an 8-GPU split -- against the oracle, bit for bit;
    window properties on every macroblock."""
    from h264_b200 import bands
    W, H, R = 3840, 2160, 64
    fr = synth.luma_sequence(W, H, 2, seed=9)
    nmb = (W // 16) * (H // 16)
    base = np.tile(np.array([[[[8, 4]]]], np.int64), (nmb, 1, 1, 1))
    pred, cen = synth.predictors(W, H, 1, seed=2, spread=3, base=base)
    s = api.Searcher(W, H, 1, R)
    s.set_cur(fr[1]); s.set_ref(0, fr[0])
    lam = (187, 187, 187)
    got = s.search_frame(pred, cen, api.make_params(lam))
    d = got[0].astype(np.int32) - cen.astype(np.int32)
    assert (np.abs(d) <= 4 * R).all() and (d % 4 == 0).all() and (got[1] > 0).all()
    of = oracle.OrcFrame(fr[1], fr[[0]], R)
    mbw, mbh = W // 16, H // 16
    rows = sorted({0, mbh - 1} | {bands.band_mb_rows(r, 8, mbh)[0] for r in range(8)} | {bands.band_mb_rows(r, 8, mbh)[1] - 1 for r in range(8)})
    rng = np.random.default_rng(4)
    sample = sorted({0, mbw - 1, nmb - mbw, nmb - 1} | {int(r * mbw + rng.integers(0, mbw)) for r in rows})
    for mb in sample:
        exp = of.search_frame(pred, cen, lam, mb_first=int(mb), mb_count=1)
        for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
            assert (a[mb] == b[mb]).all(), (n, int(mb))


def test_band_searcher_refuses_centres_beyond_its_halo():
    import torch
    from h264_b200 import bands
    W, H, R = 64, 96, 16
    b = bands.BandSearcher(W, H, 1, R, 0, 2, max_center_pel=4)
    nmb = (W // 16) * (H // 16)
    cen = torch.zeros((nmb, 1, 41, 2), dtype=torch.int16, device="cuda")
    cen[b.mb_first + 1, 0, 7, 1] = 4 * 5                       # 5 pel > max_center_pel
    z = torch.zeros_like(cen); c64 = torch.zeros((nmb, 1, 41), dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        b.search(cen, cen, api.make_params((187, 187, 187)), z, c64, z.clone(), c64.clone())


def test_pool_range_bands_equal_whole_picture():
    """h264_b200/pool_bands.py's split: the range-row bands of 1, 2, 3 and 8 ranks, searched one after the other on this GPU against
    the replicated domain plane, give the whole picture's result (the argmin is per range block)."""
    import torch
    from h264_b200 import pool_bands
    W, H, nd = 176, 144, 700
    (yr, _, _), (yc, _, _) = synth.yuv_pair(W, H, seed=8, shift=(2, -1), gain=0.9, offset=5.0)
    whole = api.PoolSearcher(W, H, W, H, nd)
    whole.set_planes(yc, yr)
    exp = whole.search()
    rp, dp = torch.from_numpy(yc).cuda(), torch.from_numpy(yr).cuda()
    for world in (1, 2, 3, 8):
        parts = []
        for rank in range(world):
            b = pool_bands.PoolBandSearcher(W, H, W, H, nd, rank, world)
            out = (torch.zeros(b.nr, dtype=torch.int32, device="cuda"), torch.zeros(b.nr, dtype=torch.uint8, device="cuda"),
                   torch.zeros(b.nr, dtype=torch.int16, device="cuda"), torch.zeros(b.nr, dtype=torch.int16, device="cuda"),
                   torch.zeros(b.nr, dtype=torch.int64, device="cuda"))
            b.search_dev(rp, dp, out)
            torch.cuda.synchronize()
            parts.append([t.cpu().numpy() for t in out])
            b.close()
        for k in range(5):
            assert (np.concatenate([p[k] for p in parts]) == exp[k]).all(), (world, k)
