#!/usr/bin/env python
"""CPU model of k_sad_fs's candidate filters (design tool, no GPU): for a sample of (MB, ref) items of the bench
workload it computes the sixteen 4x4 SADs of every candidate with numpy, walks the kernel's task order with evolving
per-partition bounds and counts, per item,
  exact : candidates that pass the exact 41-way filter (what the round-1 kernel re-evaluates),
  g4/g8 : lane groups (4 rows x 1 column / 4 rows x 2 columns) that pass the GROUP lower bound
          LB_P = sum over the 4x4 blocks k of P of min over the group's candidates of sad_k  (<= min over the group of sad_P)
and the candidates those groups hand to the exact stage.
usage: fs_filter_model.py [scenario] [nitems]   scenario: pan | off8 | noise | jitter"""
import sys
import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from h264_b200 import synth  # noqa: E402

R, LAM = 32, 187
W, H, NREFS = 1920, 1088, 4


def part_blocks():
    P = [(0, 0, 16, 16), (0, 0, 16, 8), (0, 8, 16, 8), (0, 0, 8, 16), (8, 0, 8, 16)]
    P += [(x, y, 8, 8) for y in (0, 8) for x in (0, 8)]
    P += [(x, y, 8, 4) for y in (0, 4, 8, 12) for x in (0, 8)]
    P += [(x, y, 4, 8) for y in (0, 8) for x in (0, 4, 8, 12)]
    P += [(x, y, 4, 4) for y in (0, 4, 8, 12) for x in (0, 4, 8, 12)]
    A = np.zeros((41, 16), np.int64)
    for p, (x, y, w, h) in enumerate(P):
        for by in range(y // 4, (y + h) // 4):
            for bx in range(x // 4, (x + w) // 4):
                A[p, by * 4 + bx] = 1
    return A


A = part_blocks()


def mvbits(d):
    d = np.abs(d)
    out = np.ones_like(d)
    nz = d > 0
    out[nz] = 2 * np.floor(np.log2(d[nz])).astype(np.int64) + 3
    return out


def sad_maps(cur, ref, mbx, mby, cx, cy):
    """s[dy, dx, k]: 4x4 SADs of the MB at displacement (cx + dx - R, cy + dy - R), coordinates clamped per pixel"""
    x0, y0 = mbx * 16 + cx - R, mby * 16 + cy - R
    ys = np.clip(np.arange(y0, y0 + 2 * R + 16), 0, H - 1)
    xs = np.clip(np.arange(x0, x0 + 2 * R + 16), 0, W - 1)
    win = ref[np.ix_(ys, xs)].astype(np.int16)
    c = cur[mby * 16:mby * 16 + 16, mbx * 16:mbx * 16 + 16].astype(np.int16)
    v = np.lib.stride_tricks.sliding_window_view(win, (16, 16))          # [65, 65, 16, 16]
    d = np.abs(v - c)
    return d.reshape(2 * R + 1, 2 * R + 1, 4, 4, 4, 4).sum(axis=(3, 5)).reshape(2 * R + 1, 2 * R + 1, 16).astype(np.int64)


def task_order():
    n = 2 * R + 1
    ngy = (n + 3) // 4
    gc = min(ngy - 1, R // 4)
    order, lo, hi = [gc], gc - 1, gc + 1
    while lo >= 0 or hi < ngy:
        if hi < ngy:
            order.append(hi); hi += 1
        if lo >= 0:
            order.append(lo); lo -= 1
    return order


def run_item(cur, ref, mbx, mby, pred, cen):
    """pred, cen [41, 2] quarter-pel; single centre group assumed (centres equal)"""
    n = 2 * R + 1
    cx, cy = int(cen[0, 0]) >> 2, int(cen[0, 1]) >> 2
    s = sad_maps(cur, ref, mbx, mby, cx, cy)
    sp = s @ A.T                                                          # [n, n, 41]
    dxs = 4 * (cx + np.arange(n) - R); dys = 4 * (cy + np.arange(n) - R)
    bits = mvbits(dxs[None, :, None] - pred[None, None, :, 0]) + mvbits(dys[:, None, None] - pred[None, None, :, 1])   # [n, n, 41]
    cost = (sp << 5) + LAM * bits
    # lower bound of the mv cost per candidate over the partitions' predictors (the kernel's m)
    lox, hix, loy, hiy = pred[:, 0].min(), pred[:, 0].max(), pred[:, 1].min(), pred[:, 1].max()
    distx = np.maximum(0, np.maximum(lox - dxs, dxs - hix)); disty = np.maximum(0, np.maximum(loy - dys, dys - hiy))
    m = np.minimum((LAM * mvbits(distx)) >> 5, 2047)[None, :] + np.minimum((LAM * mvbits(disty)) >> 5, 2047)[:, None]   # [n, n]
    best = cost[R, R].copy()                                              # exact pre-pass of the centre
    res = dict(exact=0, g4=0, g4c=0, g8=0, g8c=0, g2=0, g2c=0)
    for gy in task_order():
        r0, r1 = gy * 4, min(gy * 4 + 4, n)
        B = best >> 5
        blk = slice(r0, r1)
        ex = ((sp[blk] - B - 1 + m[blk][:, :, None]) < 0).any(axis=2)     # [rows, n]
        res["exact"] += int(ex.sum())
        # groups of one column x 4 rows
        M = s[blk].min(axis=0)                                            # [n, 16]
        LB = M @ A.T
        g4 = ((LB - B - 1 + m[blk].min(axis=0)[:, None]) < 0).any(axis=1)
        res["g4"] += int(g4.sum()); res["g4c"] += int(g4.sum()) * (r1 - r0)
        # groups of one column x 2 rows
        for q in range(r0, r1, 2):
            b2 = slice(q, min(q + 2, r1))
            M2 = s[b2].min(axis=0)
            g2 = (((M2 @ A.T) - B - 1 + m[b2].min(axis=0)[:, None]) < 0).any(axis=1)
            res["g2"] += int(g2.sum()); res["g2c"] += int(g2.sum()) * (b2.stop - b2.start)
        # groups of two columns (dx, dx + 4) x 4 rows, as a lane owns them
        cols = [c for c in range(64) if (c & 4) == 0]
        for c in cols:
            M8 = np.minimum(M[c], M[c + 4]); m8 = min(m[blk, c].min(), m[blk, c + 4].min())
            if (((A @ M8) - B - 1 + m8) < 0).any():
                res["g8"] += 1; res["g8c"] += 2 * (r1 - r0)
        if g4[64]:
            res["g8"] += 1; res["g8c"] += (r1 - r0)
        best = np.minimum(best, cost[blk].reshape(-1, 41).min(axis=0))
    return res


def main():
    scen = sys.argv[1] if len(sys.argv) > 1 else "pan"
    nit = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rng = np.random.default_rng(5)
    fr = synth.luma_sequence(W, H, NREFS + 2, seed=1)
    if scen == "noise":
        fr = rng.integers(0, 256, fr.shape).astype(np.uint8)
    if scen == "flat":
        fr = (128 + rng.normal(0, 2.0, fr.shape)).clip(0, 255).astype(np.uint8)
    tot = None
    for i in range(nit):
        mbx, mby, ref = int(rng.integers(0, W // 16)), int(rng.integers(0, H // 16)), int(rng.integers(0, NREFS))
        base = np.array([4 * 2 * (ref + 1), 4 * 1 * (ref + 1)])
        pred = np.tile(base, (41, 1)).astype(np.int64)
        if scen == "off8":
            pred += 4 * rng.integers(-8, 9, 2)
        if scen == "jitter":
            pred += rng.integers(-6, 7, (41, 2))
        cen = ((pred[0:1] + 2) >> 2) * 4
        cen = np.tile(cen, (41, 1))
        r = run_item(fr[NREFS], fr[NREFS - 1 - ref], mbx, mby, pred, cen)
        tot = r if tot is None else {k: tot[k] + v for k, v in r.items()}
    print(scen, {k: round(v / nit, 2) for k, v in tot.items()})


if __name__ == "__main__":
    main()


def density(scen, nit=20):
    """per task (in walk order): candidates and (candidate, partition) pairs passing the exact filter at task-start bounds"""
    rng = np.random.default_rng(5)
    fr = synth.luma_sequence(W, H, NREFS + 2, seed=1)
    if scen == "noise":
        fr = rng.integers(0, 256, fr.shape).astype(np.uint8)
    n = 2 * R + 1
    order = task_order()
    cands = np.zeros(len(order)); pairs = np.zeros(len(order)); parts = np.zeros(len(order))
    for i in range(nit):
        mbx, mby, ref = int(rng.integers(0, W // 16)), int(rng.integers(0, H // 16)), int(rng.integers(0, NREFS))
        base = np.array([4 * 2 * (ref + 1), 4 * 1 * (ref + 1)])
        pred = np.tile(base, (41, 1)).astype(np.int64)
        if scen == "off8":
            pred += 4 * rng.integers(-8, 9, 2)
        cen = np.tile(((pred[0:1] + 2) >> 2) * 4, (41, 1))
        cx, cy = int(cen[0, 0]) >> 2, int(cen[0, 1]) >> 2
        s = sad_maps(fr[NREFS], fr[NREFS - 1 - ref], mbx, mby, cx, cy)
        sp = s @ A.T
        dxs = 4 * (cx + np.arange(n) - R); dys = 4 * (cy + np.arange(n) - R)
        bits = mvbits(dxs[None, :, None] - pred[None, None, :, 0]) + mvbits(dys[:, None, None] - pred[None, None, :, 1])
        cost = (sp << 5) + LAM * bits
        m = (LAM * bits[:, :, 0]) >> 5
        best = cost[R, R].copy()
        for t, gy in enumerate(order):
            blk = slice(gy * 4, min(gy * 4 + 4, n))
            ps = (sp[blk] - (best >> 5) - 1 + m[blk][:, :, None]) < 0
            cands[t] += ps.any(axis=2).sum(); pairs[t] += ps.sum(); parts[t] += ps.any(axis=(0, 1)).sum()
            best = np.minimum(best, cost[blk].reshape(-1, 41).min(axis=0))
    print(scen, "per task in walk order: passing candidates / pairs / partitions improved")
    for t in range(len(order)):
        print(f"  task {t:2d} rowgroup {order[t]:2d}: {cands[t] / nit:7.1f} {pairs[t] / nit:8.1f} {parts[t] / nit:6.1f}")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "density":
    density(sys.argv[1], int(sys.argv[2]))
