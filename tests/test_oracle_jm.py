"""CPU tests: the restated oracle (oracle/b2_oracle.c) against (a) golden vectors captured from the
unmodified reference and (b) the reference objects themselves when oracle/_ref exists."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import synth


def test_spiral_and_mvbits_closed_forms():
    sp = oracle.spiral(16)
    assert sp.shape == (33 * 33, 2) and tuple(sp[0]) == (0, 0)
    assert [tuple(x) for x in sp[1:9]] == [(0, -1), (0, 1), (-1, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (1, 1)]
    assert len({tuple(x) for x in sp}) == 33 * 33
    # ring l occupies indices (2l-1)^2 .. (2l+1)^2-1
    ring = np.maximum(np.abs(sp[:, 0]), np.abs(sp[:, 1]))
    for l in range(1, 17):
        assert (ring[(2 * l - 1) ** 2:(2 * l + 1) ** 2] == l).all()
    assert [oracle.mvbits(d) for d in (0, 1, -1, 2, 3, 4, 7, 8, 255, 256, -511)] == [1, 3, 3, 5, 5, 7, 7, 9, 17, 19, 19]


def test_oracle_reproduces_boundary_logged_lencod_run(golden_dir):
    """Every full_search / sub_pel call of a stock lencod run on the bundled foreman clip."""
    g = np.load(os.path.join(golden_dir, "jm_wrap_foreman.npz"))
    pocs = list(g["pocs"])
    ints, subs = g["int_calls"], g["sub_calls"]
    assert len(ints) == 8118 and len(subs) == 8118       # 99 MB x 41 partitions x 2 P frames
    for fi, poc in enumerate(pocs):
        of = oracle.OrcFrame(g["cur"][fi], g["ref"][fi][None], 16)
        rows = ints[ints[:, 0] == poc]
        step = 1 if os.environ.get("B2_FULL") else 7
        for v in rows[::step]:
            _, px, py, bt, ref, pdx, pdy, cx, cy, sr, lam, min_in, ox, oy, cost = [int(x) for x in v]
            mv, c = of.call_full_search(ref, px, py, bt, (pdx, pdy), (cx, cy), sr, min_in, lam)
            assert (mv, c) == ((ox, oy), cost), v
        rows = subs[subs[:, 0] == poc]
        for v in rows[::step]:
            _, px, py, bt, ref, pdx, pdy, ix, iy, lh, lq, min_in, ox, oy, cost = [int(x) for x in v]
            mv, c = of.call_sub_pel(ref, px, py, bt, (pdx, pdy), (ix, iy), min_in, lh, lq)
            assert (mv, c) == ((ox, oy), cost), v


@pytest.mark.parametrize("name", ["a", "b"])
def test_oracle_reproduces_harness_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "jm_harness_qcif.npz"))
    W, H, R, NR = 176, 144, 16, 2
    fr = synth.luma_sequence(W, H, 3, seed=7)
    cur, refs = fr[2], fr[[1, 0]]
    restrict, spread, rmax, l0, l1, l2 = [int(x) for x in g[f"{name}_cfg"]]
    pred, cen = synth.predictors(W, H, NR, seed=11, spread=spread, rmax=rmax)
    of = oracle.OrcFrame(cur, refs, R)
    nmb = 24 if not os.environ.get("B2_FULL") else 99
    mi, ci, ms, cs = of.search_frame(pred, cen, (l0, l1, l2), restrict_mode=restrict, mb_count=nmb)
    assert (mi[:nmb] == g[f"{name}_mv_int"][:nmb]).all()
    assert (ci[:nmb] == g[f"{name}_cost_int"][:nmb]).all()
    assert (ms[:nmb] == g[f"{name}_mv_sub"][:nmb]).all()
    assert (cs[:nmb] == g[f"{name}_cost_sub"][:nmb]).all()


@pytest.mark.skipif(not oracle.have_jmref(), reason="oracle/_ref/libjmref.so only exists where /root/reference was built")
def test_oracle_vs_reference_objects_live():
    """Sub-pel planes, Hadamards, tables and a whole-frame search against the reference itself."""
    W, H, R, NR = 64, 48, 7, 1
    fr = synth.luma_sequence(W, H, 2, seed=3)
    jm = oracle.JMRef(W, H, R, NR)
    jm.set_ref(0, fr[0]); jm.set_cur(fr[1])
    of = oracle.OrcFrame(fr[1], fr[:1], R)
    assert (jm.spiral(15 * 15) // 4 == oracle.spiral(R)).all()
    md = jm.max_mvd()
    assert all(jm.mvbits(d) == oracle.mvbits(d) for d in range(-md, md + 1))
    P = of.planes(0)
    for yy in range(4):
        for xx in range(4):
            assert (jm.subplane(0, yy, xx) == P[yy, xx]).all(), (yy, xx)
    rng = np.random.default_rng(1)
    for _ in range(300):
        d = rng.integers(-255, 256, 16).astype(np.int16)
        assert jm.hadamard4x4(d) == oracle.hadamard4x4(d)
        d = rng.integers(-255, 256, 64).astype(np.int16)
        assert jm.hadamard8x8(d) == oracle.hadamard8x8(d)
    pred, cen = synth.predictors(W, H, NR, seed=2, spread=5, rmax=30)   # far predictors: clamped windows
    a = jm.search_frame(pred, cen, np.array([120, 100, 100], np.int32))
    b = of.search_frame(pred, cen, (120, 100, 100))
    for x, y in zip(a, b):
        assert (x == y).all()


@pytest.mark.skipif(not oracle.have_jmref(), reason="needs oracle/_ref/libjmref.so (built from /root/reference)")
def test_full_sub_pel_restatement_matches_reference():
    """orc_full_sub_pel == the unmodified full_sub_pel_motion_estimation (me_fullsearch.c:409-469), all 41 partitions."""
    W, H, R, NR = 64, 48, 8, 1
    fr = synth.luma_sequence(W, H, 2, seed=17)
    cur, refs = fr[1], fr[[0]]
    pred, cen = synth.predictors(W, H, NR, seed=6, spread=5, rmax=5)
    ref = oracle.JMRef(W, H, R, NR)
    ref.set_ref(0, refs[0]); ref.set_cur(cur)
    lam = np.array([120, 90, 70], np.int32)
    exp = ref.search_frame(pred, cen, lam, do_subpel=2)
    got = oracle.OrcFrame(cur, refs, R).search_frame(pred, cen, lam, do_subpel=2)
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))


@pytest.mark.skipif(not oracle.have_jmref(), reason="needs oracle/_ref/libjmref.so (built from /root/reference)")
def test_distortion_blocks_restatement_matches_reference():
    rng = np.random.default_rng(3)
    ref = oracle.JMRef(64, 48, 8, 1)
    for n in (4, 8):
        diff = rng.integers(-255, 256, (40, n * n)).astype(np.int16)
        diff[0] = 0; diff[1] = 255; diff[2] = -255
        for kind in (0, 1, 2):
            got = oracle.distortion_blocks(kind, n, diff)
            exp = np.array([ref.distortion(kind, n, diff[i]) for i in range(diff.shape[0])], np.int64)
            assert (got == exp).all(), (n, kind)


@pytest.mark.skipif(not oracle.have_jmref(), reason="needs oracle/_ref/libjmref.so (built from /root/reference)")
def test_sse_subpel_restatement_matches_reference():
    """orc_sse inside the sub-pel refinement == the unmodified JM with MEDistortionHPel = QPel = SSE."""
    W, H, R, NR = 64, 48, 8, 1
    fr = synth.luma_sequence(W, H, 2, seed=19)
    cur, refs = fr[1], fr[[0]]
    pred, cen = synth.predictors(W, H, NR, seed=7, spread=4, rmax=5)
    ref = oracle.JMRef(W, H, R, NR, metric=(0, 1, 1))
    ref.set_ref(0, refs[0]); ref.set_cur(cur)
    lam = np.array([120, 60, 60], np.int32)
    exp = ref.search_frame(pred, cen, lam)
    got = oracle.OrcFrame(cur, refs, R).search_frame(pred, cen, lam, metric_h=1, metric_q=1)
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))


@pytest.mark.skipif(not oracle.have_jmref(), reason="needs oracle/_ref/libjmref.so (built from /root/reference)")
@pytest.mark.parametrize("metric", [(0, 2, 2), (0, 0, 0), (0, 1, 1)])
def test_weighted_prediction_restatement_matches_reference(metric):
    """Weighted planes + plain distortions == the unmodified computeSADWP / SATDWP / SSEWP (UseWeightedReferenceME)."""
    W, H, R, NR = 64, 48, 8, 2
    fr = synth.luma_sequence(W, H, 3, seed=23, gain=0.93, offset=4.0)      # fading clip: weights matter
    cur, refs = fr[2], fr[[1, 0]]
    pred, cen = synth.predictors(W, H, NR, seed=7, spread=4, rmax=5)
    denom, wts, offs = 5, [30, 27], [3, -6]
    ref = oracle.JMRef(W, H, R, NR, metric=metric)
    for r in range(NR):
        ref.set_ref(r, refs[r])
    ref.set_cur(cur)
    ref.set_weights(denom, wts, offs)
    lam = np.array([120, 90, 90], np.int32)
    exp = ref.search_frame(pred, cen, lam)
    of = oracle.OrcFrame(cur, refs, R)
    for r in range(NR):
        of.set_weights(r, wts[r], offs[r], denom)
    got = of.search_frame(pred, cen, lam, metric_h=metric[1], metric_q=metric[2])
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))
    plain = oracle.OrcFrame(cur, refs, R).search_frame(pred, cen, lam, metric_h=metric[1], metric_q=metric[2])
    assert (plain[1] != got[1]).any()                                     # the weights really change the costs


def test_mc_luma_restatement_is_the_distortions_block():
    """orc_mc_luma reads through the same UMVLine4X address the pinned computeSAD does: SAD(orig, pred) of a
    partition's prediction equals the search's own SAD at that vector."""
    W, H, R, NR = 64, 48, 8, 1
    fr = synth.luma_sequence(W, H, 2, seed=29)
    cur, refs = fr[1], fr[[0]]
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, NR, seed=3, spread=0, rmax=40)            # far vectors: clamped origins
    mv_int, cost_int, _, _ = of.search_frame(pred, cen, (0, 0, 0), do_subpel=False)   # lambda 0: cost = SAD << 5
    nmb = (W // 16) * (H // 16)
    for mode, plist in ((1, [0]), (2, [1, 2]), (3, [3, 4])):
        orig, prd = of.mc_luma(np.full(nmb, mode, np.uint8), np.full((nmb, 4), 4, np.uint8), np.zeros((nmb, 4), np.int8), mv_int)
        sad4 = np.abs(orig.astype(int) - prd.astype(int)).sum(1).reshape(nmb, 16)      # per 4x4 block
        for m in range(nmb):
            for p in plist:
                bt, ox, oy, w, h = oracle.partition_geometry()[p]
                blocks = [by * 4 + bx for by in range(oy // 4, (oy + h) // 4) for bx in range(ox // 4, (ox + w) // 4)]
                assert sad4[m, blocks].sum() << 5 == cost_int[m, 0, p], (mode, m, p)


def _bipred_cases(golden_dir):
    from oracle import gen_golden_bipred as gb
    g = np.load(os.path.join(golden_dir, "jm_bipred.npz"))
    cur, refs = gb.frames()
    for k, (mh, mq, wp, t8) in enumerate(g["cases"]):
        yield gb, cur, refs, gb.jobs_of(k, wp, t8), int(mh), int(mq), bool(wp), bool(t8), g[f"c{k}"]


def test_bipred_oracle_reproduces_reference_golden(golden_dir):
    """full_search_bipred + sub_pel_bipred with computeBiPred{SAD,SSE,SATD}{1,2}: vectors captured from the unmodified JM"""
    n = 0
    for gb, cur, refs, jobs, mh, mq, wp, t8, exp in _bipred_cases(golden_dir):
        of = oracle.OrcFrame(cur, refs, gb.R)
        r = of.bipred_search(jobs, gb.LAM, metric_h=mh, metric_q=mq, test8x8=t8, wp=wp, log_denom=gb.DENOM)
        got = np.concatenate([r["cost_int"][:, None], r["cost_sub"][:, None], r["mv_int"], r["mv_sub"]], axis=1)
        assert (got == exp).all(), (mh, mq, wp, t8)
        n += len(jobs)
    assert n >= 14 * 48


@pytest.mark.skipif(not oracle.have_jmref(), reason="oracle/_ref not built (no /root/reference here)")
def test_bipred_oracle_equals_live_reference():
    W, H, R, NR = 64, 48, 6, 2
    fr = synth.luma_sequence(W, H, NR + 1, seed=17)
    cur, refs = fr[NR], fr[[1, 0]]
    of = oracle.OrcFrame(cur, refs, R)
    for metric in ((0, 2, 2), (0, 0, 0)):
        jm = oracle.JMRef(W, H, R, NR, metric=metric)
        jm.set_cur(cur)
        for r in range(NR):
            jm.set_ref(r, refs[r])
        for wp in (False, True):
            jobs = synth.bipred_jobs(W, H, NR, R, 40, seed=9 + wp, weighted=wp)
            a = jm.bipred_search(jobs, (90, 80, 70), wp=wp, log_denom=6)
            b = of.bipred_search(jobs, (90, 80, 70), metric_h=metric[1], metric_q=metric[2], wp=wp, log_denom=6)
            assert (a == b).all()


@pytest.mark.skipif(not oracle.have_jmref(), reason="oracle/_ref not built (no /root/reference here)")
def test_candidate_distortions_equal_live_reference():
    """computeSAD / computeSATD of the unmodified JM at arbitrary quarter-pel candidates (beyond the pad too)"""
    W, H, NR = 64, 48, 2
    fr = synth.luma_sequence(W, H, NR + 1, seed=23)
    cur, refs = fr[NR], fr[[1, 0]]
    of = oracle.OrcFrame(cur, refs, 4)
    jm = oracle.JMRef(W, H, 4, NR)
    jm.set_cur(cur)
    for r in range(NR):
        jm.set_ref(r, refs[r])
    c = synth.candidates(W, H, NR, 150, seed=2)
    sad, satd = of.distortion_candidates(c, 0), of.distortion_candidates(c, 2)
    c8 = synth.candidates(W, H, NR, 60, seed=3, blocktypes=(1, 2, 3, 4))
    satd8 = of.distortion_candidates(c8, 2, test8x8=True)
    for i, k in enumerate(c):
        q = (int(k["pos_x"]) * 4 + int(k["mv"][0]), int(k["pos_y"]) * 4 + int(k["mv"][1]))
        assert jm.sad(int(k["pos_x"]), int(k["pos_y"]), int(k["blocktype"]), int(k["ref"]), *q) == sad[i]
        assert jm.satd(int(k["pos_x"]), int(k["pos_y"]), int(k["blocktype"]), int(k["ref"]), *q) == satd[i]
    for i, k in enumerate(c8):
        q = (int(k["pos_x"]) * 4 + int(k["mv"][0]), int(k["pos_y"]) * 4 + int(k["mv"][1]))
        assert jm.satd(int(k["pos_x"]), int(k["pos_y"]), int(k["blocktype"]), int(k["ref"]), *q, test8x8=1) == satd8[i]


@pytest.mark.skipif(not oracle.have_jmref(), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_selection_equals_live_reference():
    """list_prediction_cost (mode_decision.c:275, list 0) of the unmodified JM for every (mode, block) entry, ties included"""
    parts = [[0], [1], [2], [3], [4], [5], [6], [7], [8], [9, 11], [10, 12], [13, 15], [14, 16], [17, 18], [19, 20], [21, 22], [23, 24],
             [25, 26, 29, 30], [27, 28, 31, 32], [33, 34, 37, 38], [35, 36, 39, 40]]
    rng = np.random.default_rng(3)
    for nrefs in (1, 3, 5):
        jm = oracle.JMRef(64, 48, 4, nrefs)
        cost = rng.integers(0, 4000, (5, nrefs, 41)).astype(np.int64)
        cost[0] = 777                                           # every reference ties: the first one stays
        cost[1] = rng.integers(995, 1005, (nrefs, 41))          # within the reference-cost differences
        br, bc = oracle.select_refs(cost, 187)
        for mb in range(5):
            for e, (mode, blk) in enumerate(oracle.SELECT_ENTRIES):
                c = [int(cost[mb, r, parts[e]].sum()) for r in range(nrefs)]
                assert jm.list_prediction_cost(mode, blk, c, 187) == (int(br[mb, e]), int(bc[mb, e])), (nrefs, mb, e)
        # list 1 of a B slice with fewer pictures than the cost array holds: the loop stops at listXsize[1]
        for ls in range(1, nrefs + 1):
            br1, bc1 = oracle.select_refs(cost, 187, list_size=ls)
            for mb in (1, 3):
                for e, (mode, blk) in enumerate(oracle.SELECT_ENTRIES):
                    c = [int(cost[mb, r, parts[e]].sum()) for r in range(ls)]
                    assert jm.list_prediction_cost(mode, blk, c, 187, list=1) == (int(br1[mb, e]), int(bc1[mb, e])), (nrefs, ls, mb, e)


def _epzs_golden():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jm_epzs.npz"))
    for tag in ("a", "b"):
        W, H, R, nrefs = int(g[f"{tag}_W"]), int(g[f"{tag}_H"]), int(g[f"{tag}_R"]), int(g[f"{tag}_nrefs"])
        jobs, preds, res, jpoc, pats = g[f"{tag}_jobs"], g[f"{tag}_preds"], g[f"{tag}_results"], g[f"{tag}_job_poc"], g[f"{tag}_patterns"]
        for poc in g[f"{tag}_pocs"]:
            refs = [g[f"{tag}_ref_{poc}_{r}"] for r in range(nrefs) if f"{tag}_ref_{poc}_{r}" in g]
            sel = np.nonzero(jpoc == poc)[0]
            yield tag, int(poc), g[f"{tag}_cur_{poc}"], refs, R, jobs[sel], preds, pats, res[sel]


def test_epzs_restatement_matches_reference_golden():
    """orc_epzs_search against what the UNMODIFIED EPZS_motion_estimation / EPZS_subMB_motion_estimation returned for the same
    jobs in stock lencod runs (tests/golden/jm_epzs.npz, oracle/gen_golden_epzs.py): cost and vector of 5 684 calls -- two
    references with the prevSad exits, every predictor group, the extended-diamond and the PMVFAST pattern chains, the dual round."""
    n = 0
    kinds = set()
    for tag, poc, cur, refs, R, jobs, preds, pats, res in _epzs_golden():
        fr = oracle.OrcFrame(cur, refs, R)
        out = fr.epzs_search(jobs, preds, pats)
        assert (out["cost"] == res[:, 0]).all(), (tag, poc)
        assert (out["mv"][:, 0] == res[:, 1]).all() and (out["mv"][:, 1] == res[:, 2]).all(), (tag, poc)
        n += len(jobs)
        kinds |= {(int(e), int(f) & 2) for e, f in zip(out["early"], jobs["flags"])}
    assert n > 5000 and len(kinds) == 4          # early and final returns of both variants occur


def test_deblock_restatement_matches_reference_golden():
    """orc_deblock_frame against the planes the UNMODIFIED DeblockFrame (JM/lencod/src/loopFilter.c:63, loop_filter_normal.c) left
    on the pictures of stock lencod runs (tests/golden/jm_deblock.npz): I / P / B pictures, two lists, 8x8 transform, filter offsets."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jm_deblock.npz"))
    n = 0
    for tag in "pbt":
        for i in range(int(g[f"{tag}_n"])):
            got = oracle.deblock_frame(*[g[f"{tag}{i}_{p}0"] for p in "yuv"], g[f"{tag}{i}_mbs"], g[f"{tag}{i}_blks"])
            for a, p in zip(got, "yuv"):
                assert (a == g[f"{tag}{i}_{p}1"]).all(), (tag, i, p)
            n += 1
    assert n == 12


def _bid_golden():
    """(tag, cur, refs, jobs, cost, par) per captured B picture of tests/golden/jm_bid.npz"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jm_bid.npz"))
    for tag in "stwp":
        for i in range(int(g[f"{tag}_n"])):
            yield tag, g[f"{tag}{i}_cur"], g[f"{tag}{i}_refs"], g[f"{tag}{i}_jobs"], g[f"{tag}{i}_cost"], g[f"{tag}{i}_par"]


def test_bid_partition_cost_restatement_matches_reference_golden():
    """orc_bid_partition_cost against what the UNMODIFIED BIDPartitionCost (JM/lencod/src/mv_search.c:1159-1250) returned for the
    calls of stock lencod runs (tests/golden/jm_bid.npz, oracle/gen_golden_bid.py): every block type, two references per list,
    8x8 transform on, implicit weighted bi-prediction on a fading clip; with bi-predictive motion estimation on, also the calls of its
    twin BPredPartitionCost (:589-700, the same cost on the bipred_mv vectors; last column of `par`)."""
    n = twin = 0
    for tag, cur, refs, jobs, cost, par in _bid_golden():
        of = oracle.OrcFrame(cur, refs, 8)
        for p in np.unique(par, axis=0):
            m = (par == p).all(axis=1)
            got = of.bid_partition_cost(jobs[m], int(p[0]), bool(p[1]), bool(p[2]), int(p[3]))
            assert (got == cost[m]).all(), (tag, p, jobs[m][got != cost[m]][:2])
            n += int(m.sum()); twin += int(m.sum()) * int(p[4])
    assert n > 1600 and twin > 50


def test_chroma_prediction_restatement_matches_reference():
    """orc_chroma_sample (the eighth-sample bilinear interpolation of chroma prediction) against the UNMODIFIED
    OneComponentChromaPrediction4x4_regenerate (JM/lencod/src/mc_prediction.c:292-353) through the harness: random planes,
    vectors that leave the picture on every side (negative coordinates: the reference divides with C truncation), different
    vectors for the four 2x2 groups of a block."""
    import ctypes as C
    rng = np.random.default_rng(8)
    jm = oracle.JMRef(64, 48, 4, 1)
    L = oracle.orc_lib()
    Wc, Hc = 32, 24
    plane = rng.integers(0, 256, (Hc, Wc)).astype(np.uint8)
    n = 0
    for trial in range(300):
        mbx, mby = int(rng.integers(0, Wc // 8)), int(rng.integers(0, Hc // 8))
        span = 40 if trial % 3 else 400
        mv16 = rng.integers(-span, span + 1, (4, 4, 2)).astype(np.int16)
        for bcx in (0, 4):
            for bcy in (0, 4):
                ref = jm.chroma_pred4x4(plane, mbx * 8, mby * 8, bcx, bcy, mv16)
                for j in range(4):
                    for i in range(4):
                        ci, cj = bcx + i, bcy + j
                        v = mv16[cj >> 1, ci >> 1]
                        got = L.orc_chroma_sample(plane.ctypes.data_as(C.c_void_p), C.c_int(Wc), C.c_int(Hc), C.c_int(8 * (mbx * 8 + ci) + int(v[0])),
                                                  C.c_int(8 * (mby * 8 + cj) + int(v[1])))
                        assert got == int(ref[j, i]), (trial, bcx, bcy, i, j)
                        n += 1
    assert n == 300 * 64
