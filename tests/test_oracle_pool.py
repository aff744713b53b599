"""CPU tests of the pool-mode oracle (oracle/b2_oracle_pool.c).  version1 has no pool search, so the oracle DEFINES
the semantics ("parity unpinned" by the reference, see its header); what can be pinned is that its exact-integer
collage error equals the reference's floating-point compute_rms expression for the same pair."""
import numpy as np

import oracle
from h264_b200 import synth


def _planes(W=64, H=48, seed=3):
    (yr, _, _), (yc, _, _) = synth.yuv_pair(W, H, seed=seed, shift=(-3, 2), gain=0.8, offset=12.0)
    return yc, yr          # range plane = current, domain plane = reference


def test_pool_positions_and_blocks():
    xy = oracle.pool_positions(64, 48, 37)
    assert xy.shape == (37, 2) and (xy >= 0).all() and (xy[:, 0] <= 64 - 16).all() and (xy[:, 1] <= 48 - 16).all()
    assert len({tuple(p) for p in xy}) == 37
    rng = np.random.default_rng(0)
    plane = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    b = oracle.pool_domain_block(plane, 5, 7).reshape(8, 8)
    ref = (plane[7:23:2, 5:21:2].astype(int) + plane[7:23:2, 6:22:2] + plane[8:24:2, 5:21:2] + plane[8:24:2, 6:22:2] + 2) >> 2
    assert (b == ref).all()


def test_isometries_are_the_dihedral_group():
    blk = np.arange(64, dtype=np.uint8)
    seen = {oracle.pool_iso(blk, i).tobytes() for i in range(8)}
    assert len(seen) == 8
    m = blk.reshape(8, 8)
    assert (oracle.pool_iso(blk, 1).reshape(8, 8) == m[:, ::-1]).all()
    assert (oracle.pool_iso(blk, 2).reshape(8, 8) == m[::-1, :]).all()
    assert (oracle.pool_iso(blk, 4).reshape(8, 8) == m.T).all()
    assert (oracle.pool_iso(blk, 5).reshape(8, 8) == np.rot90(m, -1)).all()
    assert (oracle.pool_iso(blk, 6).reshape(8, 8) == np.rot90(m, 1)).all()


def test_exact_error_matches_compute_rms_expression():
    """err_num / 640000 == the reference's double expression (V1/src/compute.c:181-182) within 1e-6 relative,
    and (aq, beta) == 100*alpha, beta of that expression, for the winning pair of every range block."""
    rp, dp = _planes()
    nd = 200
    dom, iso, aq, beta, err = oracle.pool_search(rp, dp, nd)
    xy = oracle.pool_positions(dp.shape[1], dp.shape[0], nd)
    assert (dom >= 0).any()
    for ri in range(len(dom)):
        if dom[ri] < 0:
            continue
        bx, by = ri % (rp.shape[1] // 8), ri // (rp.shape[1] // 8)
        r = oracle.pool_iso(rp[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8].reshape(64), int(iso[ri]))
        d = oracle.pool_domain_block(dp, *xy[dom[ri]])
        rms, al, be = oracle.pool_rms_double(r, d)
        assert abs(rms - err[ri] / 640000.0) <= 1e-6 * max(1.0, rms), (ri, rms, err[ri] / 640000.0)
        # alpha: exact rational truncation vs the double cast can differ only at exact multiples of 0.01
        assert abs(round(al * 100) - aq[ri]) <= 5 and be == beta[ri]


def test_planted_match_is_found():
    """A range plane made of scaled/offset copies of pool blocks: the search returns those blocks."""
    rng = np.random.default_rng(5)
    dp = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    nd = 60
    xy = oracle.pool_positions(96, 64, nd)
    rp = np.zeros((16, 32), np.uint8)
    want = [7, 23, 41, 58, 3, 11, 30, 52]
    for k, p in enumerate(want):
        d = oracle.pool_domain_block(dp, *xy[p]).astype(np.float64)
        blk = np.clip(np.rint(0.5 * (d - d.mean()) + 120), 0, 255).astype(np.uint8).reshape(8, 8)
        if k & 1:
            blk = blk.T                          # isometry 4
        rp[(k // 4) * 8:(k // 4) * 8 + 8, (k % 4) * 8:(k % 4) * 8 + 8] = blk
    dom, iso, aq, beta, err = oracle.pool_search(rp, dp, nd)
    assert list(dom) == want
    assert all(i == (4 if k & 1 else 0) for k, i in enumerate(iso))
    assert (aq == 50).all() and (beta == 120).all()


def _trunc_div(n, d):
    q = abs(n) // d
    return q if n >= 0 else -q


def _quan_a(x):
    c = _trunc_div(x, 10)
    b = x - 10 * c
    if 2 < b < 8:
        b = 5
    elif b > 7:
        b, c = 0, c + 1
    else:
        b = 0
    return c * 10 + b


def test_tie_break_is_first_maximum_in_isometry_then_pool_order():
    """Independent restatement (python integers) of the search's selection rule on inputs full of exact ties: range
    blocks symmetric under flips / transposition (several isometries give the same G for every domain) against a domain
    plane of repeated tiles (several pool entries give the same G): the oracle returns the first maximum in
    (isometry, pool index) order.  The GPU path shares one threshold among a range's 8 isometry rows and fetches the
    original pool index only on ties, so this rule is what tests/test_gpu_pool.py::test_pool_isometry_and_index_ties
    holds it to."""
    rng = np.random.default_rng(9)
    tile = rng.integers(0, 256, (16, 16), dtype=np.uint8)
    dp = np.tile(tile, (4, 5))                                      # 64 x 80
    rp = np.zeros((16, 32), np.uint8)
    for k in range(8):
        b = rng.integers(0, 256, (8, 8)).astype(np.int64)
        b = (b + b[:, ::-1], b + b.T, b + b[::-1, :] + b[:, ::-1] + b[::-1, ::-1], b + b.T + b[::-1, ::-1] + b[::-1, ::-1].T)[k % 4]
        rp[(k // 4) * 8:(k // 4) * 8 + 8, (k % 4) * 8:(k % 4) * 8 + 8] = (b * 255 // max(1, b.max())).astype(np.uint8)
    nd = 150
    dom, iso, aq, beta, err = oracle.pool_search(rp, dp, nd)
    xy = oracle.pool_positions(dp.shape[1], dp.shape[0], nd)
    doms = [oracle.pool_domain_block(dp, *p).astype(np.int64) for p in xy]
    assert len({d.tobytes() for d in doms}) < nd                    # the pool really holds duplicates
    for ri in range(8):
        blk = rp[(ri // 4) * 8:(ri // 4) * 8 + 8, (ri % 4) * 8:(ri % 4) * 8 + 8].reshape(64)
        best = (-1, 0, -1, 0)                                       # (G, iso, dom, aq)
        for i in range(8):
            r = oracle.pool_iso(blk, i).astype(np.int64)
            sr = int(r.sum())
            for j, d in enumerate(doms):
                sd, sd2 = int(d.sum()), int((d * d).sum())
                num, det = 64 * int((r * d).sum()) - sr * sd, 64 * sd2 - sd * sd
                q = _quan_a(0 if det == 0 else _trunc_div(100 * num, det))
                if q < -235 or q > 400:
                    continue
                G = 200 * q * num - q * q * det
                if G > best[0]:                                     # strict: the first maximum in (iso, index) order stays
                    best = (G, i, j, q)
        assert (dom[ri], iso[ri], aq[ri]) == (best[2], best[1], best[3]), (ri, dom[ri], iso[ri], aq[ri], best)
