"""Search -> motion-compensated prediction -> transform/quant on the device: the vectors never leave
the GPU between the stages; every stage's output is compared with the oracle's chain."""
import numpy as np
import pytest
import torch

import oracle
from h264_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_search_mc_tq_chain_matches_oracle():
    W, H, R, NR = 96, 64, 12, 2
    fr = synth.luma_sequence(W, H, NR + 1, seed=31)
    cur, refs = fr[NR], fr[[1, 0]]
    nmb = (W // 16) * (H // 16)
    pred, cen = synth.predictors(W, H, NR, seed=3, spread=4, rmax=30)          # some vectors reach far outside the picture
    rng = np.random.default_rng(4)
    mb_mode = rng.choice([1, 2, 3, 8], nmb).astype(np.uint8)
    b8mode = rng.integers(4, 8, (nmb, 4)).astype(np.uint8)
    ref8 = rng.integers(0, NR, (nmb, 4)).astype(np.int8)
    ref8[mb_mode == 1] = ref8[mb_mode == 1][:, :1]                               # one reference per partition
    ref8[mb_mode == 2, 1] = ref8[mb_mode == 2, 0]; ref8[mb_mode == 2, 3] = ref8[mb_mode == 2, 2]
    ref8[mb_mode == 3, 2] = ref8[mb_mode == 3, 0]; ref8[mb_mode == 3, 3] = ref8[mb_mode == 3, 1]
    lam = (140, 100, 100)
    # ---- GPU chain, device pointers only ----
    dev = torch.device("cuda", 0)
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    t = lambda a: torch.from_numpy(a).to(dev)
    d_mvi = torch.zeros((nmb, NR, 41, 2), dtype=torch.int16, device=dev); d_mvs = torch.zeros_like(d_mvi)
    d_ci = torch.zeros((nmb, NR, 41), dtype=torch.int64, device=dev); d_cs = torch.zeros_like(d_ci)
    s.search_frame_dev(t(pred), t(cen), api.make_params(lam), d_mvi, d_ci, d_mvs, d_cs)
    d_orig = torch.zeros((nmb * 16, 16), dtype=torch.uint8, device=dev); d_pred = torch.zeros_like(d_orig)
    s.mc_luma_dev(t(mb_mode), t(b8mode), t(ref8), d_mvs, d_orig, d_pred)
    p = api.tq_default_params(4, 28, 0)
    d_level = torch.zeros((nmb * 16, 16), dtype=torch.int16, device=dev); d_run = torch.zeros((nmb * 16, 16), dtype=torch.uint8, device=dev)
    d_recon = torch.zeros_like(d_orig); d_cost = torch.zeros(nmb * 16, dtype=torch.int32, device=dev); d_nz = torch.zeros(nmb * 16, dtype=torch.uint8, device=dev)
    api.tq_dev(p, d_orig, d_pred, 4, d_level, d_run, d_recon, d_cost, d_nz)
    torch.cuda.synchronize()
    # ---- oracle chain ----
    of = oracle.OrcFrame(cur, refs, R)
    exp = of.search_frame(pred, cen, lam)
    assert (d_mvs.cpu().numpy() == exp[2]).all()
    o_orig, o_pred = of.mc_luma(mb_mode, b8mode, ref8, exp[2])
    assert (d_orig.cpu().numpy() == o_orig).all()
    assert (d_pred.cpu().numpy() == o_pred).all()
    q = oracle.tq_params(api.tq_params_table(p, 4), 28, cavlc=1)
    e_level, e_run, e_recon, e_cost, e_nz = oracle.tq(q, o_orig, o_pred, 4)
    for g, e, n in zip((d_level, d_run, d_recon, d_cost, d_nz), (e_level, e_run, e_recon, e_cost, e_nz), ("level", "run", "recon", "cost", "nz")):
        assert (g.cpu().numpy() == e).all(), n
    # the prediction is a good one: the reconstruction is close to the original
    assert np.abs(d_recon.cpu().numpy().astype(int) - o_orig.astype(int)).mean() < 8.0


def test_reference_selection_on_the_device_matches_oracle():
    """b2me_select_refs_dev (list_prediction_cost, list 0) on the cost array the search left on the device"""
    import torch
    W, H, R, NR = 96, 64, 8, 4
    fr = synth.luma_sequence(W, H, NR + 1, seed=7)
    cur, refs = fr[NR], fr[[3, 2, 1, 0]]
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    pred, cen = synth.predictors(W, H, NR, seed=3, spread=2, rmax=4)
    dev = torch.device("cuda", 0)
    dp, dc = torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev)
    mvi = torch.zeros((s.nmb, NR, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
    ci = torch.zeros((s.nmb, NR, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
    s.search_frame_dev(dp, dc, api.make_params((187, 187, 187)), mvi, ci, mvs, cs)
    br = torch.zeros((s.nmb, 21), dtype=torch.int8, device=dev); bc = torch.zeros((s.nmb, 21), dtype=torch.int64, device=dev)
    for lam in (187, 4000):
        for ls in range(1, cs.shape[1] + 1):            # list 1 of a B slice: the loop stops at listXsize
            s.select_refs_list_dev(cs, ls, lam, br, bc)
            torch.cuda.synchronize()
            ebr, ebc = oracle.select_refs(cs.cpu().numpy(), lam, list_size=ls)
            assert (br.cpu().numpy() == ebr).all() and (bc.cpu().numpy() == ebc).all(), ls
        s.select_refs_dev(cs, lam, br, bc)
        torch.cuda.synchronize()
        ebr, ebc = oracle.select_refs(cs.cpu().numpy(), lam)
        assert (br.cpu().numpy() == ebr).all() and (bc.cpu().numpy() == ebc).all(), lam
    assert len(np.unique(br.cpu().numpy())) > 1


def test_compact_frame_search_equals_full_arrays_plus_selection():
    """b2me_search_frame_best (one predictor per (MB, ref) up; best reference, cost and vector per entry / partition down) ==
    b2me_search_frame on the expanded predictors followed by the oracle's list_prediction_cost and the gather of the chosen vectors."""
    W, H, R, NR = 96, 64, 16, 3
    fr = synth.luma_sequence(W, H, NR + 1, seed=5)
    cur, refs = fr[NR], fr[[NR - 1 - r for r in range(NR)]]
    s = api.Searcher(W, H, NR, R)
    s.set_cur(cur)
    for r in range(NR):
        s.set_ref(r, refs[r])
    rng = np.random.default_rng(3)
    nmb = (W // 16) * (H // 16)
    pred_mb = rng.integers(-30, 31, (nmb, NR, 2)).astype(np.int16)
    lam = 187
    p = api.make_params((lam, lam, lam))
    br, bc, bm = s.search_frame_best(pred_mb, p, lam)
    pred = np.broadcast_to(pred_mb[:, :, None, :], (nmb, NR, 41, 2)).copy()
    cen = (((pred.astype(np.int32) + 2) >> 2) * 4).astype(np.int16)
    _, _, mv_sub, cost_sub = s.search_frame(pred, cen, p)
    ebr, ebc = oracle.select_refs(cost_sub, lam)
    assert (br == ebr).all() and (bc == np.minimum(ebc, 0x7fffffff)).all()
    for e, parts in enumerate(oracle.ENTRY_PARTS):
        for q in parts:
            assert (bm[:, q] == mv_sub[np.arange(nmb), ebr[:, e], q]).all(), (e, q)


def test_luma_and_chroma_prediction_of_both_lists_matches_oracle():
    """b2me_mc_mb_dev: luma + chroma prediction of 4:2:0 macroblocks from list 0, list 1 or both (bi_prediction), every macroblock
    mode and 8x8 sub-mode, vectors that leave the picture; the luma of list-0 macroblocks equals b2me_mc_luma_dev's."""
    W, H, R, NR = 96, 64, 12, 3
    rng = np.random.default_rng(14)
    yuv = [np.frombuffer(synth.yuv420_sequence(W, H, 1, seed=60 + i), np.uint8) for i in range(NR + 1)]
    Y = [f[:W * H].reshape(H, W) for f in yuv]
    U = [f[W * H:W * H * 5 // 4].reshape(H // 2, W // 2) for f in yuv]
    V = [f[W * H * 5 // 4:].reshape(H // 2, W // 2) for f in yuv]
    nmb = (W // 16) * (H // 16)
    mb_mode = rng.choice([1, 2, 3, 8], nmb).astype(np.uint8)
    b8mode = rng.integers(4, 8, (nmb, 4)).astype(np.uint8)
    pdir = rng.integers(0, 3, (nmb, 4)).astype(np.uint8)
    ref8 = rng.integers(0, NR, (nmb, 2, 4)).astype(np.int8)
    mv0 = rng.integers(-60, 61, (nmb, NR, 41, 2)).astype(np.int16); mv1 = rng.integers(-60, 61, (nmb, NR, 41, 2)).astype(np.int16)
    mv0[:3] = rng.integers(-500, 501, (3, NR, 41, 2)); mv1[-3:] = rng.integers(-500, 501, (3, NR, 41, 2))      # far outside the picture
    dev = torch.device("cuda", 0)
    s = api.Searcher(W, H, NR, R)
    s.set_cur(Y[NR]); s.set_cur_chroma(U[NR], V[NR])
    for r in range(NR):
        s.set_ref(r, Y[r]); s.set_ref_chroma(r, U[r], V[r])
    t = lambda a: torch.from_numpy(a).to(dev)
    oy = torch.zeros((nmb * 16, 16), dtype=torch.uint8, device=dev); py = torch.zeros_like(oy)
    oc = torch.zeros((nmb, 2, 4, 16), dtype=torch.uint8, device=dev); pc = torch.zeros_like(oc)
    s.mc_mb_dev(t(mb_mode), t(b8mode), t(pdir), t(ref8), t(mv0), t(mv1), oy, py, oc, pc)
    torch.cuda.synchronize()
    of = oracle.OrcFrame(Y[NR], [Y[r] for r in range(NR)], R)
    curc = np.stack([U[NR], V[NR]]); refc = np.stack([np.stack([U[r], V[r]]) for r in range(NR)])
    e = of.mc_mb(mb_mode, b8mode, pdir, ref8, mv0, mv1, curc, refc)
    for g, x, n in zip((oy, py, oc, pc), e, ("orig_y", "pred_y", "orig_c", "pred_c")):
        assert (g.cpu().numpy() == x).all(), n
    # list-0 only: the luma equals the single-list entry's
    pd0 = np.zeros_like(pdir)
    s.mc_mb_dev(t(mb_mode), t(b8mode), t(pd0), t(ref8), t(mv0), t(mv1), oy, py, oc, pc)
    o2 = torch.zeros_like(oy); p2 = torch.zeros_like(py)
    s.mc_luma_dev(t(mb_mode), t(b8mode), t(np.ascontiguousarray(ref8[:, 0])), t(mv0), o2, p2)
    torch.cuda.synchronize()
    assert (py == p2).all() and (oy == o2).all()
