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


def test_pool_1080p_16k_sample_matches_oracle():
    """BASELINE config 5: 1080p range plane against a 16 K pool: sampled range blocks re-derived by brute force."""
    W, H, nd = 1920, 1080, 16384
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

    for ri in rng.integers(0, s.nr, 24):
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
