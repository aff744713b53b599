"""TEST INFRASTRUCTURE ONLY.  Pins the per-pair fit of the pool-mode fractal matching (oracle/b2_oracle_pool.c orc_pool_pair,
what k_frac_pool's epilogue evaluates) to the reference ITSELF: the unmodified version1 objects
(oracle/_ref/libv1ref.so = V1/src/compute.c + block_enc.c + oracle/v1_harness.c).

version1 has no pool / isometry search (SURVEY Q-F1), but the pool search is nothing else than its compute_rms
(V1/src/compute.c:6-189) applied to (range block under an isometry, 2:1-decimated domain block) pairs.  So the pairs are laid
into the reference's own planes -- range blocks on the 8x8 grid of imgY_org, domain blocks at the same grid positions of the
C reference plane --, the unmodified compute_range_Sum / compute_domain_Sum (compute.c:686, 277) build the sum tables, and the
unmodified compute_rms returns (alpha, beta, rms) per pair.  Pairs: random textures, the 8 isometries of smooth blocks
against decimated 16x16 blocks of a real-looking plane, anti-correlated pairs (negative alpha), flat domains (det == 0),
steep fits (alpha beyond MAX_ALPHA / MIN_ALPHA: rejected), saturated blocks.

Writes tests/golden/v1_pool_fit.npz: r [n][64], d [n][64] uint8 and the reference's a100 = round(100 alpha), beta, rms.
Run in the build container (needs /root/reference):  python oracle/gen_golden_pool_fit.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
W = H = 128                       # 256 pairs per frame on the 8x8 grid


def make_pairs(seed=20261019, nframes=16):
    import oracle
    from h264_b200 import synth
    rng = np.random.default_rng(seed)
    n = nframes * (W // 8) * (H // 8)
    r = np.zeros((n, 64), np.uint8); d = np.zeros((n, 64), np.uint8)
    (yr, _, _), (yc, _, _) = synth.yuv_pair(352, 288, seed=5, shift=(-3, 2), gain=0.8, offset=12.0)
    for k in range(n):
        kind = k % 8
        if kind == 0:                                   # independent random textures
            r[k] = rng.integers(0, 256, 64); d[k] = rng.integers(0, 256, 64)
        elif kind in (1, 2, 3):                         # range block of the current plane under an isometry vs a decimated domain block
            bx, by = rng.integers(0, 352 // 8), rng.integers(0, 288 // 8)
            blk = yc[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8].reshape(64)
            r[k] = oracle.pool_iso(blk, int(rng.integers(0, 8)))
            d[k] = oracle.pool_domain_block(yr, int(rng.integers(0, 352 - 16)), int(rng.integers(0, 288 - 16)))
        elif kind == 4:                                 # strongly related: r = a * d + b + noise, a in [-3, 5] (negatives, rejects)
            d[k] = rng.integers(0, 256, 64)
            a, b = rng.uniform(-3.0, 5.0), rng.uniform(-40, 200)
            r[k] = np.clip(np.rint(a * (d[k].astype(float) - d[k].mean()) * rng.uniform(0.05, 1.0) + b + rng.normal(0, 2, 64)), 0, 255)
        elif kind == 5:                                 # flat domain (det == 0) or flat range
            if rng.random() < 0.5:
                d[k] = rng.integers(0, 256); r[k] = rng.integers(0, 256, 64)
            else:
                r[k] = rng.integers(0, 256); d[k] = rng.integers(0, 256, 64)
        elif kind == 6:                                 # low-variance domain vs high-variance range: |alpha| large
            d[k] = np.clip(128 + rng.integers(-2, 3, 64), 0, 255)
            r[k] = np.clip(128 + 40 * (d[k].astype(int) - 128) * rng.choice([-1, 1]) + rng.integers(-3, 4, 64), 0, 255)
        else:                                           # saturated / two-level blocks
            d[k] = rng.choice([0, 255], 64); r[k] = rng.choice([0, 255], 64) if rng.random() < 0.5 else 255 - d[k]
    return r, d


def reference_fit(r, d):
    """the unmodified compute_rms on every pair: (a100, beta, rms)"""
    import oracle
    v = oracle.V1Ref(W, H, 7)
    n = len(r); per = (W // 8) * (H // 8)
    a100 = np.zeros(n, np.int32); beta = np.zeros(n, np.float64); rms = np.zeros(n, np.float64)
    zc = np.zeros((H // 2, W // 2), np.uint8)
    for f in range(0, n, per):
        org = np.zeros((H, W), np.uint8); ref = np.zeros((H, W), np.uint8)
        for k in range(f, min(f + per, n)):
            gx, gy = ((k - f) % (W // 8)) * 8, ((k - f) // (W // 8)) * 8
            org[gy:gy + 8, gx:gx + 8] = r[k].reshape(8, 8); ref[gy:gy + 8, gx:gx + 8] = d[k].reshape(8, 8)
        v.set_ref(0, ref, zc, zc, build_sums=True)     # compute_domain_Sum (unmodified)
        v.set_cur(org, zc, zc)                         # compute_range_Sum (unmodified)
        for k in range(f, min(f + per, n)):
            gx, gy = ((k - f) % (W // 8)) * 8, ((k - f) // (W // 8)) * 8
            e, al, be = v.compute_rms(0, gx, gy, gx, gy, 8, 8, 1)      # compute_rms (unmodified), domain at the same grid position
            a100[k] = int(round(al * 100)); beta[k] = be; rms[k] = e
    return a100, beta, rms


if __name__ == "__main__":
    r, d = make_pairs()
    a100, beta, rms = reference_fit(r, d)
    np.savez_compressed(os.path.join(GOLD, "v1_pool_fit.npz"), r=r, d=d, a100=a100, beta=beta, rms=rms)
    print(f"{len(r)} pairs: {int((rms > 1e29).sum())} rejected, {int((a100 < 0).sum())} negative alpha, "
          f"{int((a100 == 0).sum())} zero alpha, a100 range [{a100.min()}, {a100.max()}]")
