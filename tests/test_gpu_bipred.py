"""GPU parity tests (through the C ABI): bi-predictive block search (b2me_bipred_search) against the golden vectors
captured from the unmodified JM (full_search_bipred_motion_estimation + sub_pel_bipred_motion_estimation with
computeBiPred{SAD,SSE,SATD}{1,2}) and against the oracle on other seeded inputs."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu


def _flat(r):
    return np.concatenate([r["cost_int"][:, None], r["cost_sub"][:, None], r["mv_int"], r["mv_sub"]], axis=1)


def _searcher(cur, refs, R):
    H, W = cur.shape
    s = api.Searcher(W, H, len(refs), R)
    s.set_cur(cur)
    for r in range(len(refs)):
        s.set_ref(r, refs[r])
    return s


def test_bipred_matches_reference_golden(golden_dir):
    from oracle import gen_golden_bipred as gb
    g = np.load(os.path.join(golden_dir, "jm_bipred.npz"))
    cur, refs = gb.frames()
    s = _searcher(cur, refs, gb.R)
    for k, (mh, mq, wp, t8) in enumerate(g["cases"]):
        jobs = gb.jobs_of(k, wp, t8)
        got = s.bipred_search(jobs, api.make_params(gb.LAM, metric_h=int(mh), metric_q=int(mq)), apply_weights=bool(wp),
                              log_denom=gb.DENOM, test8x8=bool(t8))
        bad = np.nonzero((_flat(got) != g[f"c{k}"]).any(1))[0]
        assert len(bad) == 0, ((mh, mq, wp, t8), jobs[bad[:2]], _flat(got)[bad[:2]], g[f"c{k}"][bad[:2]])


@pytest.mark.parametrize("W,H,R,NR,lam", [(64, 48, 16, 2, (40, 30, 30)), (176, 144, 32, 3, (700, 500, 400))])
def test_bipred_matches_oracle(W, H, R, NR, lam):
    fr = synth.luma_sequence(W, H, NR + 1, seed=W + R)
    cur, refs = fr[NR], fr[list(range(NR - 1, -1, -1))]
    s = _searcher(cur, refs, R)
    of = oracle.OrcFrame(cur, refs, R)
    for wp in (False, True):
        jobs = synth.bipred_jobs(W, H, NR, R, 64, seed=3 + wp, weighted=wp, rmax=9)
        for mh, mq in ((2, 2), (0, 2)):
            got = s.bipred_search(jobs, api.make_params(lam, metric_h=mh, metric_q=mq), apply_weights=wp, log_denom=5)
            exp = of.bipred_search(jobs, lam, metric_h=mh, metric_q=mq, wp=wp, log_denom=5)
            assert (_flat(got) == _flat(exp)).all(), (wp, mh, mq)
        got = s.bipred_search(jobs, api.make_params(lam, do_subpel=False), apply_weights=wp, log_denom=5)
        exp = of.bipred_search(jobs, lam, do_subpel=False, wp=wp, log_denom=5)
        assert (_flat(got) == _flat(exp)).all()


def test_bipred_81_positions_and_subpel_only_jobs():
    """params.subpel_full: full_sub_pel_bipred_motion_estimation (me_fullsearch.c:478); search_range -1: the sub-pel call alone"""
    W, H, R, NR, lam = 96, 64, 8, 2, (187, 150, 120)
    fr = synth.luma_sequence(W, H, NR + 1, seed=41)
    cur, refs = fr[NR], fr[[1, 0]]
    s = _searcher(cur, refs, R)
    of = oracle.OrcFrame(cur, refs, R)
    rng = np.random.default_rng(2)
    for sub_only in (False, True):
        jobs = synth.bipred_jobs(W, H, NR, R, 64, seed=8 + sub_only)
        if sub_only:
            jobs["search_range"] = -1
            jobs["mv1"] = rng.integers(-24, 25, (64, 2))
        for mh, mq in ((2, 2), (0, 2), (1, 1)):
            for full in (False, True):
                got = s.bipred_search(jobs, api.make_params(lam, metric_h=mh, metric_q=mq, subpel_full=full))
                exp = of.bipred_search(jobs, lam, metric_h=mh, metric_q=mq, do_subpel=2 if full else 1)
                assert (_flat(got) == _flat(exp)).all(), (sub_only, mh, mq, full)


def test_bipred_error_codes():
    W, H, R = 64, 48, 8
    fr = synth.luma_sequence(W, H, 2, seed=1)
    s = _searcher(fr[1], fr[[0]], R)
    p = api.make_params((100, 100, 100))
    jobs = synth.bipred_jobs(W, H, 1, R, 4, seed=1)
    assert len(s.bipred_search(jobs[:0], p)) == 0                      # empty batch
    with pytest.raises(api.B2Error):                                   # weighted SATD with the 8x8 Hadamard (Q-J5)
        s.bipred_search(jobs, p, apply_weights=True, log_denom=5, test8x8=True)
    bad = jobs.copy(); bad[1]["ref2"] = 3
    with pytest.raises(api.B2Error):
        s.bipred_search(bad, p)
    bad = jobs.copy(); bad[2]["mv1"] = (5, 0)                          # sub-pel centre is not an integer-search input
    with pytest.raises(api.B2Error):
        s.bipred_search(bad, p)
    got = s.bipred_search(jobs, p)                                     # the context is still usable
    exp = oracle.OrcFrame(fr[1], fr[[0]], R).bipred_search(jobs, (100, 100, 100))
    assert (_flat(got) == _flat(exp)).all()


def test_candidate_distortions_match_oracle():
    """b2me_distortion_candidates: computeSAD / SSE / SATD (4x4 and 8x8 Hadamard) at arbitrary candidates, weighted slot too"""
    W, H, NR = 176, 144, 3
    fr = synth.luma_sequence(W, H, NR + 1, seed=31)
    cur, refs = fr[NR], fr[[2, 1, 0]]
    s = api.Searcher(W, H, NR, 16)
    s.set_cur(cur)
    s.set_ref_weights(1, 40, -3, 5)
    for r in range(NR):
        s.set_ref(r, refs[r])
    of = oracle.OrcFrame(cur, refs, 16)
    of.set_weights(1, 40, -3, 5)
    c = synth.candidates(W, H, NR, 3000, seed=4)
    for metric in (0, 1, 2):
        assert (s.distortion_candidates(c, metric) == of.distortion_candidates(c, metric)).all(), metric
    c8 = synth.candidates(W, H, NR, 1000, seed=5, blocktypes=(1, 2, 3, 4))
    assert (s.distortion_candidates(c8, 2, test8x8=True) == of.distortion_candidates(c8, 2, test8x8=True)).all()
    assert len(s.distortion_candidates(c[:0], 0)) == 0
    bad = c[:8].copy(); bad[3]["blocktype"] = 9
    with pytest.raises(api.B2Error):
        s.distortion_candidates(bad, 0)


def test_bid_partition_cost_matches_reference_golden_and_oracle():
    """b2me_bid_partition_cost (k_bid_cost, one warp per partition) against the costs the UNMODIFIED BIDPartitionCost returned in
    stock lencod runs (tests/golden/jm_bid.npz; incl. the calls of the twin BPredPartitionCost) and against the oracle on seeded records: SAD / SSE / SATD, 4x4 and 8x8 blocks,
    weighted, macroblocks at the picture borders with vectors that leave the picture."""
    from test_oracle_jm import _bid_golden
    n = 0
    for tag, cur, refs, jobs, cost, par in _bid_golden():
        s = _searcher(cur, refs, 8)
        for p in np.unique(par, axis=0):
            m = (par == p).all(axis=1)
            got = s.bid_partition_cost(jobs[m], int(p[0]), bool(p[1]), bool(p[2]), int(p[3]))
            assert (got == cost[m]).all(), (tag, p, jobs[m][got != cost[m]][:2])
            n += int(m.sum())
        s.close()
    assert n > 1600
    W, H, NR = 96, 64, 3
    fr = synth.luma_sequence(W, H, NR + 1, seed=31)
    cur, refs = fr[NR], fr[[2, 1, 0]]
    s, of = _searcher(cur, refs, 8), oracle.OrcFrame(cur, refs, 8)
    for k, (metric, t8, wp, denom) in enumerate([(0, 0, 0, 0), (1, 0, 0, 0), (2, 0, 0, 0), (2, 1, 0, 0), (0, 1, 1, 5), (1, 1, 0, 0), (2, 0, 1, 6), (2, 1, 1, 3)]):
        jobs = synth.bid_jobs(W, H, NR, 600, seed=40 + k, weighted=bool(wp), rmax=14)
        got, exp = s.bid_partition_cost(jobs, metric, bool(t8), bool(wp), denom), of.bid_partition_cost(jobs, metric, bool(t8), bool(wp), denom)
        assert (got == exp).all(), ((metric, t8, wp, denom), jobs[got != exp][:2], got[got != exp][:2], exp[got != exp][:2])
    bad = synth.bid_jobs(W, H, NR, 4, seed=1); bad[2]["ref_l1"] = NR
    with pytest.raises(api.B2Error):
        s.bid_partition_cost(bad, 2)
