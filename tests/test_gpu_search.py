"""GPU parity tests (through the C ABI): CUDA motion search vs the oracle / golden vectors."""
import os

import numpy as np
import pytest

import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu


def _setup(W, H, R, NR, seed):
    fr = synth.luma_sequence(W, H, NR + 1, seed=seed)
    cur, refs = fr[NR], fr[list(range(NR - 1, -1, -1))]
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    return s, cur, refs


def test_subpel_planes_bit_exact():
    W, H = 176, 144
    s, cur, refs = _setup(W, H, 16, 2, seed=5)
    for r in range(2):
        P = oracle.subpel_planes(refs[r])
        for yy in range(4):
            for xx in range(4):
                assert (s.subplane(r, yy, xx) == P[yy, xx]).all(), (r, yy, xx)


@pytest.mark.parametrize("name", ["a", "b"])
def test_search_matches_reference_golden(golden_dir, name):
    """Golden outputs of the unmodified JM objects (oracle/gen_golden_jm.py)."""
    g = np.load(os.path.join(golden_dir, "jm_harness_qcif.npz"))
    W, H, R, NR = 176, 144, 16, 2
    fr = synth.luma_sequence(W, H, 3, seed=7)
    cur, refs = fr[2], fr[[1, 0]]
    restrict, spread, rmax, l0, l1, l2 = [int(x) for x in g[f"{name}_cfg"]]
    pred, cen = synth.predictors(W, H, NR, seed=11, spread=spread, rmax=rmax)
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    mi, ci, ms, cs = s.search_frame(pred, cen, api.make_params((l0, l1, l2), restrict_mode=restrict))
    assert (mi == g[f"{name}_mv_int"]).all()
    assert (ci == g[f"{name}_cost_int"]).all()
    assert (ms == g[f"{name}_mv_sub"]).all()
    assert (cs == g[f"{name}_cost_sub"]).all()


@pytest.mark.parametrize("W,H,R,NR,spread,rmax,lam", [
    (64, 48, 7, 1, 0, 3, (187, 187, 187)),        # shared centre per MB
    (64, 48, 7, 2, 6, 40, (40, 30, 30)),          # far, per-partition predictors: clamped windows
    (96, 64, 32, 1, 2, 8, (1200, 900, 900)),      # +-32, high lambda
    (48, 48, 16, 1, 0, 0, (4, 4, 4)),             # tiny lambda: many ties / improvements
])
def test_search_matches_oracle(W, H, R, NR, spread, rmax, lam):
    s, cur, refs = _setup(W, H, R, NR, seed=W + R)
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, NR, seed=3, spread=spread, rmax=rmax)
    got = s.search_frame(pred, cen, api.make_params(lam))
    exp = of.search_frame(pred, cen, lam)
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))


@pytest.mark.parametrize("shift,rmax", [((3, -2), 2), ((-9, 6), 12), ((0, 0), 0)])
def test_pan_clip_uniform_vectors(shift, rmax):
    """A noiseless pan: every partition of a macroblock lands on the same vector, which is the case k_subpel_refine
    serves by its uniform-vector path (tile SATDs summed through the block-size tree); frame borders, vectors that
    point outside the picture (per-tile clamp) and a second reference with different content included."""
    W, H, R, NR = 80, 64, 16, 2
    rng = np.random.default_rng(17)
    big = rng.integers(0, 256, (H + 64, W + 64)).astype(np.float64)
    big = (big + np.roll(big, 1, 0) + np.roll(big, 1, 1) + np.roll(big, (1, 1), (0, 1))) / 4
    ref0 = big[32:32 + H, 32:32 + W].astype(np.uint8)
    cur = big[32 + shift[1]:32 + shift[1] + H, 32 + shift[0]:32 + shift[0] + W].astype(np.uint8)
    ref1 = rng.integers(0, 256, (H, W), dtype=np.uint8)
    refs = np.stack([ref0, ref1])
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, NR, seed=2, spread=0, rmax=rmax)
    got = s.search_frame(pred, cen, api.make_params((60, 45, 45)))
    exp = of.search_frame(pred, cen, (60, 45, 45))
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))
    mi = got[0][:, 0]
    if rmax <= 2:                                                   # (far random predictors leave the true vector outside many windows)
        assert ((mi == mi[:, :1]).all(-1).all(-1)).mean() > 0.5    # the uniform path really ran


def test_flat_and_noise_content():
    """Flat frames (all-zero SAD, ties decided by spiral order) and pure noise."""
    W, H, R = 48, 32, 7
    rng = np.random.default_rng(0)
    for cur, ref in ((np.full((H, W), 77, np.uint8), np.full((H, W), 77, np.uint8)),
                     (rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8))):
        s = api.Searcher(W, H, 1, R)
        s.set_cur(cur); s.set_ref(0, ref)
        of = oracle.OrcFrame(cur, ref[None], R)
        pred, cen = synth.predictors(W, H, 1, seed=9, spread=5, rmax=5)
        got = s.search_frame(pred, cen, api.make_params((90, 70, 70)))
        exp = of.search_frame(pred, cen, (90, 70, 70))
        for a, b in zip(got, exp):
            assert (a == b).all()


def test_sad_subpel_metric_and_no_subpel():
    W, H, R = 64, 48, 7
    s, cur, refs = _setup(W, H, R, 1, seed=21)
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, 1, seed=4, spread=3, rmax=4)
    got = s.search_frame(pred, cen, api.make_params((100, 100, 100), metric_h=0, metric_q=0))
    exp = of.search_frame(pred, cen, (100, 100, 100), metric_h=0, metric_q=0)
    for a, b in zip(got, exp):
        assert (a == b).all()
    got = s.search_frame(pred, cen, api.make_params((100, 100, 100), metric_h=0, metric_q=2))
    exp = of.search_frame(pred, cen, (100, 100, 100), metric_h=0, metric_q=2)
    for a, b in zip(got, exp):
        assert (a == b).all()
    got = s.search_frame(pred, cen, api.make_params((100, 100, 100), do_subpel=False))
    exp = of.search_frame(pred, cen, (100, 100, 100), do_subpel=False)
    assert (got[0] == exp[0]).all() and (got[1] == exp[1]).all()


def test_block_search_dropin_matches_boundary_log(golden_dir):
    """b2me_block_search == one full_search + sub_pel call of the stock lencod run."""
    g = np.load(os.path.join(golden_dir, "jm_wrap_foreman.npz"))
    ints, subs = g["int_calls"], g["sub_calls"]
    poc = int(g["pocs"][0])
    s = api.Searcher(176, 144, 1, 16)
    s.set_cur(g["cur"][0]); s.set_ref(0, g["ref"][0])
    rows_i = ints[ints[:, 0] == poc]
    rows_s = subs[subs[:, 0] == poc]
    idx = np.linspace(0, len(rows_i) - 1, 160).astype(int)
    for k in idx:
        _, px, py, bt, ref, pdx, pdy, cx, cy, sr, lam, min_in, ox, oy, cost = [int(x) for x in rows_i[k]]
        v = [int(x) for x in rows_s[k]]
        assert v[1:5] == [px, py, bt, ref]
        lh, lq, sox, soy, scost = v[9], v[10], v[12], v[13], v[14]
        p = api.make_params((lam, lh, lq), min_mcost=min_in)
        mi, ci, ms, cs = s.block_search(px, py, bt, ref, (pdx, pdy), (cx, cy), p, sr)
        assert (mi, ci) == ((ox, oy), cost), rows_i[k]
        assert (ms, cs) == ((sox, soy), scost), rows_s[k]


def test_error_codes():
    s = api.Searcher(64, 48, 1, 7)
    with pytest.raises(api.B2Error):
        api.Searcher(60, 48, 1, 7)                      # not a multiple of 16
    with pytest.raises(api.B2Error):
        s.block_search(0, 0, 9, 0, (0, 0), (0, 0), api.make_params(100), 7)   # bad blocktype
    with pytest.raises(api.B2Error):
        s.search_frame(*synth.predictors(64, 48, 1), api.make_params(100, metric_h=3))  # no such metric
    pred, cen = synth.predictors(64, 48, 1)
    cen = cen.copy(); cen[0, 0, 0, 0] = 2               # quarter-pel centre
    s.set_cur(np.zeros((48, 64), np.uint8)); s.set_ref(0, np.zeros((48, 64), np.uint8))
    with pytest.raises(api.B2Error):
        s.search_frame(pred, cen, api.make_params(100))


def test_device_entry_error_flag_and_weighted_row_ranges():
    """The _dev searches never synchronise: a quarter-pel centre is reported by b2me_check_errors (once, then cleared).  A weighted
    reference slot refuses a partial row range (its plane set is stored weighted as a whole)."""
    import torch
    W, H = 64, 48
    s = api.Searcher(W, H, 1, 7)
    dev = torch.device("cuda", 0)
    z = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    s.set_cur_dev(z); s.set_ref_dev(0, z)
    pred, cen = synth.predictors(W, H, 1)
    bad = cen.copy(); bad[0, 0, 0, 0] = 2
    nmb = s.nmb
    mvi = torch.zeros((nmb, 1, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
    ci = torch.zeros((nmb, 1, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
    p = api.make_params(100)
    s.search_frame_dev(torch.from_numpy(pred).to(dev), torch.from_numpy(bad).to(dev), p, mvi, ci, mvs, cs)
    assert s.L.b2me_check_errors(s.h, None) == -1        # B2ME_EINVAL
    assert s.L.b2me_check_errors(s.h, None) == 0         # cleared
    s.search_frame_dev(torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev), p, mvi, ci, mvs, cs)
    assert s.L.b2me_check_errors(s.h, None) == 0
    s.set_ref_weights(0, 40, -3, 5)
    with pytest.raises(api.B2Error):
        s.set_ref_rows_dev(0, z, 16, 16)
    s.set_ref_rows_dev(0, z, 0, H)                       # the whole picture is fine
    torch.cuda.synchronize()


def test_search_range_64_kernel():
    """+-64 takes the other k_sad_fs instantiation (window pitch 160, 12 worker warps)."""
    W, H, R = 64, 48, 64
    s, cur, refs = _setup(W, H, R, 1, seed=33)
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, 1, seed=8, spread=4, rmax=9)
    got = s.search_frame(pred, cen, api.make_params((120, 100, 100)))
    exp = of.search_frame(pred, cen, (120, 100, 100))
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))


def test_banded_search_equals_full_frame():
    """MB-row bands with halo rows only (h264_b200/bands.py): every band's results equal the full-frame search."""
    import torch
    from h264_b200 import bands
    W, H, R, world = 96, 144, 16, 3
    fr = synth.luma_sequence(W, H, 2, seed=12)
    cur, ref = fr[1], fr[0]
    pred, cen = synth.predictors(W, H, 1, seed=5, spread=3, rmax=8)        # |centre| <= 8 pel + rounding
    s = api.Searcher(W, H, 1, R)
    s.set_cur(cur); s.set_ref(0, ref)
    full = s.search_frame(pred, cen, api.make_params((150, 120, 120)))
    dev = torch.device("cuda", 0)
    d_pred, d_cen = torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev)
    nmb = (W // 16) * (H // 16)
    for rank in range(world):
        b = bands.BandSearcher(W, H, 1, R, rank, world, max_center_pel=12)
        lo, hi = bands.needed_rows(rank, world, H // 16, R, 12)
        part = np.zeros_like(ref); part[lo:hi] = ref[lo:hi]                # what the halo exchange delivers
        b.set_cur_dev(torch.from_numpy(cur).to(dev))
        if rank == 1:
            b.s.set_ref_dev(0, torch.from_numpy(part).to(dev))
        else:
            # planes of the needed rows only (b2me_set_ref_rows_dev) over planes of ANOTHER picture: stale rows must not be read
            b.s.set_ref_dev(0, torch.from_numpy(np.ascontiguousarray(cur[::-1])).to(dev))
            b.s.set_ref_rows_dev(0, torch.from_numpy(part).to(dev), lo, hi - lo)
        mvi = torch.zeros((nmb, 1, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
        ci = torch.zeros((nmb, 1, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
        b.search(d_pred, d_cen, api.make_params((150, 120, 120)), mvi, ci, mvs, cs)
        torch.cuda.synchronize()
        sl = slice(b.mb_first, b.mb_first + b.mb_count)
        for got, exp in zip((mvi, ci, mvs, cs), full):
            assert (got.cpu().numpy()[sl] == exp[sl]).all(), rank


def test_full_sub_pel_mode_matches_oracle():
    """b2me_search_params.subpel_full: full_sub_pel_motion_estimation (81 quarter-pel positions), SATD and SAD."""
    W, H, R = 64, 48, 8
    s, cur, refs = _setup(W, H, R, 2, seed=41)
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, 2, seed=9, spread=5, rmax=5)
    for mq in (2, 0):
        got = s.search_frame(pred, cen, api.make_params((130, 100, 80), metric_h=mq, metric_q=mq, subpel_full=True))
        exp = of.search_frame(pred, cen, (130, 100, 80), metric_h=mq, metric_q=mq, do_subpel=2)
        for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
            assert (a == b).all(), (mq, n, int((a != b).sum()))


def test_distortion_blocks_match_oracle():
    rng = np.random.default_rng(8)
    for n in (4, 8):
        diff = rng.integers(-255, 256, (1000, n * n)).astype(np.int16)
        diff[0] = 0; diff[1] = 255; diff[2] = -255
        for kind in (0, 1, 2):
            assert (api.distortion_blocks(kind, n, diff) == oracle.distortion_blocks(kind, n, diff)).all(), (n, kind)


def test_sse_subpel_metric_matches_oracle():
    """MEDistortionHPel/QPel = SSE (computeSSE, me_distortion.c:1190-1255) and the mixed SAD/SSE/SATD settings."""
    W, H, R = 64, 48, 7
    s, cur, refs = _setup(W, H, R, 1, seed=52)
    of = oracle.OrcFrame(cur, refs, R)
    pred, cen = synth.predictors(W, H, 1, seed=4, spread=3, rmax=4)
    for mh, mq in ((1, 1), (1, 2), (0, 1), (2, 1)):
        got = s.search_frame(pred, cen, api.make_params((100, 60, 60), metric_h=mh, metric_q=mq))
        exp = of.search_frame(pred, cen, (100, 60, 60), metric_h=mh, metric_q=mq)
        for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
            assert (a == b).all(), (mh, mq, n, int((a != b).sum()))


def test_weighted_prediction_matches_oracle():
    """b2me_set_ref_weights: explicit WP of the single-list search (computeSADWP / SATDWP), two references."""
    W, H, R, NR = 64, 48, 8, 2
    fr = synth.luma_sequence(W, H, 3, seed=23, gain=0.93, offset=4.0)
    cur, refs = fr[2], fr[[1, 0]]
    pred, cen = synth.predictors(W, H, NR, seed=7, spread=4, rmax=5)
    denom, wts, offs = 5, [30, 27], [3, -6]
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    of = oracle.OrcFrame(cur, refs, R)
    for r in range(NR):
        s.set_ref_weights(r, wts[r], offs[r], denom)
        s.set_ref(r, refs[r])
        of.set_weights(r, wts[r], offs[r], denom)
    got = s.search_frame(pred, cen, api.make_params((120, 90, 90)))
    exp = of.search_frame(pred, cen, (120, 90, 90))
    for a, b, n in zip(got, exp, ("mv_int", "cost_int", "mv_sub", "cost_sub")):
        assert (a == b).all(), (n, int((a != b).sum()))
    s.set_ref_weights(0, 32, 0, 5, apply=False); s.set_ref(0, refs[0])      # off again: plain planes
    assert (s.subplane(0, 0, 0) == oracle.subpel_planes(refs[0])[0, 0]).all()


def test_refined_results_only_and_upload_ordering():
    """b2me_search_frame with mv_int = cost_int = NULL (the refined results alone come down), directly behind host-pointer
    b2me_set_ref calls whose plane build is still running on the context's stream; the device-pointer call on ANOTHER
    stream must wait for that build too."""
    import torch
    W, H, R, NR = 128, 96, 16, 2
    fr = synth.luma_sequence(W, H, NR + 2, seed=5)
    pred, cen = synth.predictors(W, H, NR, seed=6, spread=2, rmax=6)
    lam = (187, 150, 150)
    p = api.make_params(lam)
    s = api.Searcher(W, H, NR, R)
    for rep in range(3):                            # new pictures every round: stale planes would show
        cur, refs = fr[NR + (rep & 1)], fr[[1 + (rep & 1), rep & 1]]
        exp = oracle.OrcFrame(cur, refs, R).search_frame(pred, cen, lam)
        s.set_cur(cur)
        for r in range(NR):
            s.set_ref(r, refs[r])
        got = s.search_frame(pred, cen, p, want_int=False)
        assert got[0] is None and got[1] is None
        assert (got[2] == exp[2]).all() and (got[3] == exp[3]).all()
        for r in range(NR):
            s.set_ref(r, refs[r])
        st = torch.cuda.Stream()
        dev = torch.device("cuda", 0)
        dp, dc = torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev)
        mvi = torch.zeros((s.nmb, NR, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
        ci = torch.zeros((s.nmb, NR, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
        torch.cuda.synchronize()
        for r in range(NR):
            s.set_ref(r, refs[r])
        s.search_frame_dev(dp, dc, p, mvi, ci, mvs, cs, stream=st.cuda_stream)
        st.synchronize()
        assert (mvs.cpu().numpy() == exp[2]).all() and (cs.cpu().numpy() == exp[3]).all()
    import ctypes as C
    assert s.L.b2me_search_frame(s.h, pred.ctypes.data_as(C.c_void_p), cen.ctypes.data_as(C.c_void_p),
                                 C.byref(api.make_params(lam, do_subpel=False)), None, None, None, None) == -1


def test_sad_table_matches_oracle_distortions():
    """b2me_sad_table (setup_fast_full_search's BlockSAD tables): sampled (partition, position) entries against the oracle's
    computeSAD at that candidate, border macroblocks and a centre that pushes the window outside the picture included."""
    import ctypes as C
    W, H, R, NR = 96, 64, 16, 2
    fr = synth.luma_sequence(W, H, NR + 1, seed=12)
    cur, refs = fr[NR], fr[[1, 0]]
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    of = oracle.OrcFrame(cur, refs, R)
    geo = oracle.partition_geometry()
    rng = np.random.default_rng(5)
    for (mbx, mby, ref, cen, sr) in ((0, 0, 0, (-40, -28), 16), (5, 3, 1, (36, 44), 16), (2, 1, 0, (0, 4), 7)):
        npos = (2 * sr + 1) ** 2
        tab = np.zeros((41, npos), np.uint16)
        cm = (C.c_int16 * 2)(*cen)
        r = s.L.b2me_sad_table(s.h, mbx, mby, ref, cm, sr, tab.ctypes.data_as(C.c_void_p))
        assert r == 0
        cands = np.zeros(0, synth.CANDIDATE)
        picks = [(int(p), int(pos)) for p, pos in zip(rng.integers(0, 41, 300), rng.integers(0, npos, 300))] + [(0, 0), (40, npos - 1)]
        cands = np.zeros(len(picks), synth.CANDIDATE)
        for k, (p, pos) in enumerate(picks):
            bt, ox, oy, bw, bh = geo[p]
            sx, sy = oracle.spiral_xy(pos)
            cands[k]["pos_x"], cands[k]["pos_y"], cands[k]["blocktype"], cands[k]["ref"] = mbx * 16 + ox, mby * 16 + oy, bt, ref
            cands[k]["mv"] = (cen[0] + 4 * sx, cen[1] + 4 * sy)
        exp = of.distortion_candidates(cands, 0) >> 5
        got = np.array([tab[p, pos] for p, pos in picks], np.int64)
        assert (got == exp).all(), np.flatnonzero(got != exp)[:5]


def test_epzs_search_matches_oracle_and_reference_golden():
    """b2me_epzs_search (k_epzs, one warp per job) against the oracle on the jobs captured from stock lencod runs, and against what
    the unmodified EPZS_motion_estimation / EPZS_subMB_motion_estimation returned for them (tests/golden/jm_epzs.npz):
    cost, vector, exit path and the number of search points, bit for bit."""
    from test_oracle_jm import _epzs_golden
    n = 0
    for tag, poc, cur, refs, R, jobs, preds, pats, res in _epzs_golden():
        H, W = cur.shape
        s = api.Searcher(W, H, len(refs), R)
        s.set_cur(cur)
        for r, rf in enumerate(refs):
            s.set_ref(r, rf)
        out = s.epzs_search(jobs, preds, pats)
        exp = oracle.OrcFrame(cur, refs, R).epzs_search(jobs, preds, pats)
        for k in ("cost", "mv", "early", "npoints"):
            assert (out[k] == exp[k]).all(), (tag, poc, k)
        assert (out["cost"] == res[:, 0]).all() and (out["mv"][:, 0] == res[:, 1]).all() and (out["mv"][:, 1] == res[:, 2]).all()
        n += len(jobs)
    assert n > 5000


def test_epzs_search_synthetic_jobs_and_errors():
    """Jobs the captures do not hold: blocks at the picture border with vectors far outside (the padded planes' clamp), every block
    type, search windows of +-64, a pattern that chains to another one; then the error codes (pattern index out of range,
    predictor range outside the array)."""
    rng = np.random.default_rng(11)
    W, H, R = 128, 96, 64
    fr = synth.luma_sequence(W, H, 3, seed=4)
    s = api.Searcher(W, H, 2, R)
    s.set_cur(fr[2]); s.set_ref(0, fr[1]); s.set_ref(1, fr[0])
    o = oracle.OrcFrame(fr[2], [fr[1], fr[0]], R)
    pats = np.zeros(3, synth.EPZS_PATTERN)
    sd = [(0, 4, 3, 3), (4, 0, 0, 3), (0, -4, 1, 3), (-4, 0, 2, 3)]
    sq = [(0, 4, 7, 3), (4, 4, 7, 5), (4, 0, 1, 3), (4, -4, 1, 5), (0, -4, 3, 3), (-4, -4, 3, 5), (-4, 0, 5, 3), (-4, 4, 5, 5)]
    ld = [(0, 8, 6, 5), (4, 4, 0, 3), (8, 0, 0, 5), (4, -4, 2, 3), (0, -8, 2, 5), (-4, -4, 4, 3), (-8, 0, 4, 5), (-4, 4, 6, 3)]
    for i, (pts, stop, nxt) in enumerate(((sd, 1, 0), (sq, 1, 1), (ld, 0, 0))):       # the third chains to the small diamond
        pats[i]["npoints"] = len(pts); pats[i]["stop_search"] = stop; pats[i]["next_last"] = 1; pats[i]["next_pattern"] = nxt
        pats[i]["pt"][:len(pts)] = pts
    sizes = {1: (16, 16), 2: (16, 8), 3: (8, 16), 4: (8, 8), 5: (8, 4), 6: (4, 8), 7: (4, 4)}
    n = 600
    jobs = np.zeros(n, synth.EPZS_JOB)
    preds = rng.integers(-4 * 70, 4 * 70, (n * 12, 2)).astype(np.int16)
    for i in range(n):
        bt = 1 + i % 7
        bw, bh = sizes[bt]
        j = jobs[i]
        j["blocktype"] = bt; j["ref"] = i % 2
        j["pos_x"] = rng.integers(0, (W - bw) // 4 + 1) * 4 if i % 5 else (0 if i % 2 else W - bw)
        j["pos_y"] = rng.integers(0, (H - bh) // 4 + 1) * 4 if i % 3 else (0 if i % 2 else H - bh)
        j["mv"] = rng.integers(-40, 41, 2) * 4
        j["pred"] = j["mv"] + rng.integers(-3, 4, 2)
        j["range"] = (4 * R, 4 * R) if i % 4 else (4 * 8, 4 * 16)
        j["mv_range"] = 10 if bt < 5 else 12
        j["flags"] = (1 if j["ref"] else 0) | (2 if bt > 4 else 0) | (4 if i % 6 else 0) | (8 if i % 3 else 0)
        j["lambda_factor"] = int(rng.integers(1, 2000))
        j["medthres"] = int(rng.integers(0, 3000)) << 5
        j["stop0"] = j["medthres"] + 2 * j["lambda_factor"]
        j["stop"] = int(rng.integers(0, 20000)) << 5 if i % 7 else 0
        j["prev_sad"] = int(rng.integers(0, 30000)) << 5
        j["pred_first"] = 12 * i
        j["npred"] = (5, 2, 3, 2); j["cond_host"] = (1, i % 2, (i >> 1) % 2, i % 4); j["fixed_edge"] = 1 if i % 11 == 0 else 0
        j["pat_init"] = 2 if i % 2 else 1; j["pat_sd"] = 0; j["pat_sq"] = 1; j["pat_else"] = 1 if i % 3 else 2; j["pat_dual"] = 2 if i % 5 else 0
    out = s.epzs_search(jobs, preds, pats)
    exp = o.epzs_search(jobs, preds, pats)
    for k in ("cost", "mv", "early", "npoints"):
        assert (out[k] == exp[k]).all(), k
    assert out["npoints"].max() > 30 and (out["early"] == 0).any() and (out["early"] == 1).any()
    bad = jobs[:4].copy(); bad["pat_dual"] = 3
    with pytest.raises(RuntimeError, match="out of range"):
        s.epzs_search(bad, preds, pats)
    bad = jobs[:4].copy(); bad["pred_first"] = len(preds) - 3
    with pytest.raises(RuntimeError):
        s.epzs_search(bad, preds, pats)
    assert (s.epzs_search(jobs[:8], preds, pats)["cost"] == exp["cost"][:8]).all()      # the context still works afterwards
