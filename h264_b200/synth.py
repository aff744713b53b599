"""Deterministic synthetic YUV 4:2:0 luma sequences: smoothed-noise texture
panning a few pixels per frame, a moving object, gain/offset drift and sensor noise.  Used by
tests, bench.py and the golden-vector generators so that every side sees the same input."""
import numpy as np


def _smooth(a, k):
    for _ in range(k):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5.0
    return a


def luma_sequence(W, H, nframes, seed=20261018, pan=(2, 1), obj=(3, 1), noise=1.5, gain=1.0, offset=0.0):
    """Returns uint8 [nframes, H, W]."""
    rng = np.random.default_rng(seed)
    big = rng.normal(0, 1, (H + 64, W + 64))
    tex = _smooth(big, 3)
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    fine = rng.normal(0, 1, (H + 64, W + 64))
    tex = 40 + 160 * tex + 14 * _smooth(fine, 1)
    ob = 128 + 90 * _smooth(rng.normal(0, 1, (32, 32)), 1)
    frames = np.zeros((nframes, H, W), np.uint8)
    for t in range(nframes):
        sx, sy = (pan[0] * t) % 32, (pan[1] * t) % 32
        f = tex[16 + sy:16 + sy + H, 16 + sx:16 + sx + W].copy()
        ox, oy = (W // 3 + obj[0] * t) % (W - 32), (H // 3 + obj[1] * t) % (H - 32)
        f[oy:oy + 32, ox:ox + 32] = ob
        f = (gain ** t) * f + offset * t + rng.normal(0, noise, f.shape)
        frames[t] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return frames


def predictors(W, H, nrefs, seed=1, spread=0, base=None, rmax=6):
    """Synthetic quarter-pel MV predictors [nmb, nrefs, 41, 2] int16 and the integer search centres
    JM derives from them (mv_search.c:931-932: ((p+2)>>2)*4).  spread=0: one predictor per
    (MB, ref) shared by its 41 partitions; spread>0: per-partition jitter of +-spread qpel."""
    rng = np.random.default_rng(seed)
    nmb = (W // 16) * (H // 16)
    p = rng.integers(-4 * rmax, 4 * rmax + 1, (nmb, nrefs, 1, 2)) if base is None else base
    p = np.broadcast_to(p, (nmb, nrefs, 41, 2)).copy()
    if spread:
        p += rng.integers(-spread, spread + 1, p.shape)
    p = p.astype(np.int16)
    c = (((p.astype(np.int32) + 2) >> 2) * 4).astype(np.int16)
    return p, c


def yuv_pair(W, H, seed=20261018, shift=(-3, 2), gain=0.9, offset=10.0, noise=1.0):
    """A (reference, current) 4:2:0 frame pair for the fractal search:
    current = gain * reference displaced by `shift` + offset + noise, with a flat patch (variance 0
    -> alpha = 0 path) and a saturated patch.  Returns ((Yr,Ur,Vr), (Yc,Uc,Vc)) uint8."""
    rng = np.random.default_rng(seed)

    def tex(h, w, k):
        a = _smooth(rng.normal(0, 1, (h + 32, w + 32)), k)
        a = (a - a.min()) / (a.max() - a.min())
        return 30 + 190 * a + 10 * _smooth(rng.normal(0, 1, (h + 32, w + 32)), 1)

    out = []
    planes = [tex(H, W, 3), tex(H // 2, W // 2, 2), tex(H // 2, W // 2, 2)]
    ref, cur = [], []
    for i, t in enumerate(planes):
        h, w = (H, W) if i == 0 else (H // 2, W // 2)
        sx, sy = shift if i == 0 else (shift[0] // 2, shift[1] // 2)
        r = t[16:16 + h, 16:16 + w]
        c = gain * t[16 + sy:16 + sy + h, 16 + sx:16 + sx + w] + offset + rng.normal(0, noise, (h, w))
        r = r.copy(); c = c.copy()
        r[h // 2:h // 2 + 20, w // 4:w // 4 + 24] = 77          # flat domain area: det == 0
        c[h // 4:h // 4 + 18, w // 2:w // 2 + 20] = 200          # flat range area
        c[:12, :12] = 255
        ref.append(np.clip(np.rint(r), 0, 255).astype(np.uint8))
        cur.append(np.clip(np.rint(c), 0, 255).astype(np.uint8))
    return tuple(ref), tuple(cur)


def residual_blocks(nblk, n, seed=1):
    """(orig, pred) uint8 [nblk][n*n]: a mix of smooth, noisy, identical (all-zero residual),
    single-coefficient and saturated blocks for the transform/quant path."""
    rng = np.random.default_rng(seed)
    m = n * n
    base = rng.integers(0, 256, (nblk, 1)).astype(np.int64)
    ramp = (np.arange(m) % n)[None, :] * rng.integers(-6, 7, (nblk, 1)) + (np.arange(m) // n)[None, :] * rng.integers(-6, 7, (nblk, 1))
    amp = rng.choice([0, 1, 2, 4, 8, 20, 60, 255], (nblk, 1))
    orig = np.clip(base + ramp + rng.integers(-1, 2, (nblk, m)) * amp, 0, 255)
    pred = np.clip(base + rng.integers(-1, 2, (nblk, m)) * rng.choice([0, 1, 3, 10, 255], (nblk, 1)), 0, 255)
    same = rng.random(nblk) < 0.1
    pred[same] = orig[same]
    pred[::17] = 0; orig[::17] = 255          # maximal residual
    return orig.astype(np.uint8), pred.astype(np.uint8)


def yuv420_sequence(W, H, nframes, seed=20261018):
    """Planar 4:2:0 clip as bytes (Y, U, V per frame): luma_sequence for Y, U = 128, V = a smooth
    function of Y (SURVEY 8(d) config 1 input)."""
    y = luma_sequence(W, H, nframes, seed=seed)
    out = bytearray()
    for t in range(nframes):
        out += y[t].tobytes()
        out += np.full((H // 2, W // 2), 128, np.uint8).tobytes()
        v = (y[t][::2, ::2].astype(np.int32) // 4 + 96).astype(np.uint8)
        out += v.tobytes()
    return bytes(out)


# b2me_bipred_job / b2me_bipred_result of include/b2me.h as numpy record types
BIPRED_JOB = np.dtype([("min_mcost", np.int64), ("pos_x", np.int16), ("pos_y", np.int16), ("blocktype", np.int16),
                       ("ref1", np.int16), ("ref2", np.int16), ("search_range", np.int16),
                       ("pred1", np.int16, 2), ("pred2", np.int16, 2), ("mv1", np.int16, 2), ("mv2", np.int16, 2),
                       ("weight1", np.int16), ("weight2", np.int16), ("offset_bi", np.int16), ("reserved", np.int16)], align=True)
# b2me_bid_job (include/b2me.h): one BIDPartitionCost call
BID_JOB = np.dtype([("mb_x", np.int16), ("mb_y", np.int16), ("blocktype", np.int16), ("block8x8", np.int16), ("ref_l0", np.int16), ("ref_l1", np.int16),
                    ("mv_l0", np.int16, (4, 2)), ("mv_l1", np.int16, (4, 2)), ("weight_l0", np.int16), ("weight_l1", np.int16), ("offset_bi", np.int16),
                    ("reserved", np.int16), ("mvd_bits", np.int32), ("lambda_factor", np.int32)], align=True)
assert BID_JOB.itemsize == 60
BIPRED_RESULT = np.dtype([("cost_int", np.int64), ("cost_sub", np.int64), ("mv_int", np.int16, 2), ("mv_sub", np.int16, 2)], align=True)
# (blocktype, ox, oy, w, h) of the 41 partitions in include/b2me.h's order
PART_GEOM = ([(1, 0, 0, 16, 16), (2, 0, 0, 16, 8), (2, 0, 8, 16, 8), (3, 0, 0, 8, 16), (3, 8, 0, 8, 16)] +
             [(4, x, y, 8, 8) for y in (0, 8) for x in (0, 8)] + [(5, x, y, 8, 4) for y in (0, 4, 8, 12) for x in (0, 8)] +
             [(6, x, y, 4, 8) for y in (0, 8) for x in (0, 4, 8, 12)] + [(7, x, y, 4, 4) for y in (0, 4, 8, 12) for x in (0, 4, 8, 12)])
# the refinement patterns of JM's EPZS as b2me_epzs_pattern records: (points (dx, dy, start_nmbr, next_points), stop_search, next_last, next_pattern)
EPZS_SDIAMOND = [(0, 4, 3, 3), (4, 0, 0, 3), (0, -4, 1, 3), (-4, 0, 2, 3)]
EPZS_SQUARE = [(0, 4, 7, 3), (4, 4, 7, 5), (4, 0, 1, 3), (4, -4, 1, 5), (0, -4, 3, 3), (-4, -4, 3, 5), (-4, 0, 5, 3), (-4, 4, 5, 5)]
EPZS_EDIAMOND = [(-4, 4, 10, 5), (0, 8, 10, 8), (0, 4, 10, 7), (4, 4, 1, 5), (8, 0, 1, 8), (4, 0, 1, 7), (4, -4, 4, 5), (0, -8, 4, 8),
                 (0, -4, 4, 7), (-4, -4, 7, 5), (-8, 0, 7, 8), (-4, 0, 7, 7)]


# include/b2me.h b2me_epzs_job / b2me_epzs_result / b2me_epzs_pattern (C layout, 104 / 16 / 112 bytes)
EPZS_JOB = np.dtype([("pos_x", np.int16), ("pos_y", np.int16), ("blocktype", np.int16), ("ref", np.int16), ("mv", np.int16, 2),
                     ("pred", np.int16, 2), ("range", np.int16, 2), ("mv_range", np.int16), ("flags", np.int16), ("lambda_factor", np.int32),
                     ("stop0", np.int64), ("stop", np.int64), ("medthres", np.int64), ("prev_sad", np.int64), ("pred_first", np.int32),
                     ("npred", np.int16, 4), ("cond_host", np.int16, 4), ("fixed_edge", np.int16), ("pat_init", np.int16), ("pat_sd", np.int16),
                     ("pat_sq", np.int16), ("pat_else", np.int16), ("pat_dual", np.int16), ("pad_", np.int16)], align=True)
EPZS_RESULT = np.dtype([("cost", np.int64), ("mv", np.int16, 2), ("early", np.int16), ("npoints", np.int16)], align=True)
EPZS_PATTERN = np.dtype([("npoints", np.int32), ("stop_search", np.int32), ("next_last", np.int32), ("next_pattern", np.int32),
                         ("pt", np.int16, (12, 4))], align=True)
def epzs_patterns():
    """[small diamond, square, extended diamond], each ending in itself (stop_search = next_last = 1)"""
    pats = np.zeros(3, EPZS_PATTERN)
    for i, pts in enumerate((EPZS_SDIAMOND, EPZS_SQUARE, EPZS_EDIAMOND)):
        pats[i]["npoints"] = len(pts); pats[i]["stop_search"] = 1; pats[i]["next_last"] = 1; pats[i]["next_pattern"] = i
        pats[i]["pt"][:len(pts)] = pts
    return pats


# include/b2me.h b2dbk_mb / b2dbk_blk (12 bytes each)
DBK_MB = np.dtype([("intra", np.uint8), ("qp", np.uint8), ("qpc_u", np.uint8), ("qpc_v", np.uint8), ("transform8x8", np.uint8), ("disable", np.uint8),
                   ("alpha_off", np.int8), ("beta_off", np.int8), ("cbp_blk", np.uint16), ("pad_", np.uint16)], align=True)
DBK_BLK = np.dtype([("mv", np.int16, (2, 2)), ("ref", np.int16, 2)], align=True)
CANDIDATE = np.dtype([("pos_x", np.int16), ("pos_y", np.int16), ("blocktype", np.int16), ("ref", np.int16), ("mv", np.int16, 2)])
_BS = {1: (16, 16), 2: (16, 8), 3: (8, 16), 4: (8, 8), 5: (8, 4), 6: (4, 8), 7: (4, 4)}


def bid_jobs(W, H, nrefs, n, seed=1, weighted=False, rmax=9):
    """n seeded BIDPartitionCost calls: any macroblock (picture borders included), any blocktype and partition index, vectors at any
    quarter-pel (far enough to leave the picture at the borders), plausible mvd bit counts"""
    rng = np.random.default_rng(seed)
    j = np.zeros(n, BID_JOB)
    j["mb_x"] = 16 * rng.integers(0, W // 16, n); j["mb_y"] = 16 * rng.integers(0, H // 16, n)
    j["blocktype"] = rng.integers(1, 8, n)
    nblk = np.where(j["blocktype"] == 1, 1, np.where(j["blocktype"] < 4, 2, 4))
    j["block8x8"] = rng.integers(0, 4, n) % nblk
    j["ref_l0"] = rng.integers(0, nrefs, n); j["ref_l1"] = rng.integers(0, nrefs, n)
    j["mv_l0"] = rng.integers(-4 * rmax, 4 * rmax + 1, (n, 4, 2)); j["mv_l1"] = rng.integers(-4 * rmax, 4 * rmax + 1, (n, 4, 2))
    j["mvd_bits"] = rng.integers(2, 60, n); j["lambda_factor"] = rng.integers(20, 600, n)
    if weighted:
        j["weight_l0"] = rng.integers(10, 60, n); j["weight_l1"] = rng.integers(10, 60, n); j["offset_bi"] = rng.integers(-8, 9, n)
    return j


def bipred_jobs(W, H, nrefs, R, n, seed=1, weighted=False, blocktypes=(1, 2, 3, 4, 5, 6, 7), rmax=5):
    """n seeded calls of the bi-predictive block search: blocks on their own grid anywhere in the picture (borders
    included), the searched vector's centre integer-pel (BiPredBlockMotionSearch rounds it, mv_search.c:1079-1085),
    the static vector of the other list at any quarter-pel, ranges 0..R, some calls with a finite incoming bound."""
    rng = np.random.default_rng(seed)
    j = np.zeros(n, BIPRED_JOB)
    for i in range(n):
        bt = int(rng.choice(blocktypes)); w, h = _BS[bt]
        j[i]["blocktype"] = bt
        j[i]["pos_x"] = w * rng.integers(0, W // w); j[i]["pos_y"] = h * rng.integers(0, H // h)
        j[i]["ref1"] = rng.integers(0, nrefs); j[i]["ref2"] = rng.integers(0, nrefs)
        j[i]["search_range"] = rng.integers(0, R + 1) if i % 3 else R
        j[i]["pred1"] = rng.integers(-4 * rmax, 4 * rmax + 1, 2); j[i]["pred2"] = rng.integers(-4 * rmax, 4 * rmax + 1, 2)
        j[i]["mv1"] = 4 * rng.integers(-rmax, rmax + 1, 2)
        j[i]["mv2"] = rng.integers(-4 * rmax, 4 * rmax + 1, 2)
        j[i]["min_mcost"] = (0x7fffffff << 5) if i % 4 else int(rng.integers(2000, 60000))
        if weighted:
            j[i]["weight1"] = rng.integers(10, 60); j[i]["weight2"] = rng.integers(10, 60); j[i]["offset_bi"] = rng.integers(-8, 9)
        else:
            j[i]["weight1"] = j[i]["weight2"] = 32
    return j


def candidates(W, H, nrefs, n, seed=1, blocktypes=(1, 2, 3, 4, 5, 6, 7), reach=40):
    """n seeded (block, reference, vector) triples, vectors up to +-reach pel at any quarter-pel (beyond the pad too)"""
    rng = np.random.default_rng(seed)
    c = np.zeros(n, CANDIDATE)
    for i in range(n):
        bt = int(rng.choice(blocktypes)); w, h = _BS[bt]
        c[i]["blocktype"] = bt
        c[i]["pos_x"] = w * rng.integers(0, W // w); c[i]["pos_y"] = h * rng.integers(0, H // h)
        c[i]["ref"] = rng.integers(0, nrefs)
        c[i]["mv"] = rng.integers(-4 * reach, 4 * reach + 1, 2)
    return c
