"""Host-side (no GPU) checks of the two identities k_subpel_refine's SATD path rests on, against the oracle:
 * HadamardSAD4x4 (me_distortion.c:175-258) == sum over the four horizontal +-1 patterns k of
   max(|p|, |r|) + max(|q|, |t|), with p, q = R0k +- R1k and r, t = R2k +- R3k, where Rrk is the dot product of row r
   of (cur - ref) with pattern k -- the form the kernel evaluates with dp4a (tile_satd in csrc/subpel_refine.cu);
 * the 41 partition sums of per-tile values through the 4x4 -> 8x4 / 4x8 -> 8x8 -> 16x8 / 8x16 -> 16x16 tree of xor
   shuffles land on the partition indices of oracle.partition_geometry() (the uniform-vector path)."""
import numpy as np

import oracle

PATTERNS = np.array([[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]], np.int64)


def satd_dp4a_form(cur, ref):
    """cur, ref: uint8 [4][4]; mirrors tile_satd: the reference row enters with the negated pattern."""
    R = np.zeros((4, 4), np.int64)
    for r in range(4):
        for k in range(4):
            R[r, k] = int((cur[r].astype(np.int64) * PATTERNS[k]).sum()) + int((ref[r].astype(np.int64) * -PATTERNS[k]).sum())
    s = 0
    for k in range(4):
        p, q, r_, t = R[0, k] + R[1, k], R[0, k] - R[1, k], R[2, k] + R[3, k], R[2, k] - R[3, k]
        s += max(abs(p), abs(r_)) + max(abs(q), abs(t))
    return int(s)


def test_dp4a_form_equals_hadamard_sad4x4():
    rng = np.random.default_rng(3)
    cases = [(rng.integers(0, 256, (4, 4), dtype=np.uint8), rng.integers(0, 256, (4, 4), dtype=np.uint8)) for _ in range(400)]
    cases += [(np.full((4, 4), 255, np.uint8), np.zeros((4, 4), np.uint8)), (np.zeros((4, 4), np.uint8), np.full((4, 4), 255, np.uint8)),
              (np.full((4, 4), 7, np.uint8), np.full((4, 4), 7, np.uint8))]
    chk = np.indices((4, 4)).sum(0) % 2
    cases.append(((chk * 255).astype(np.uint8), ((1 - chk) * 255).astype(np.uint8)))       # largest alternating coefficient
    for cur, ref in cases:
        diff = cur.astype(np.int16) - ref.astype(np.int16)
        assert satd_dp4a_form(cur, ref) == oracle.hadamard4x4(diff.reshape(-1))


def _shfl_xor(v, m):
    return v[np.arange(32) ^ m]


def test_shuffle_tree_lands_on_the_partition_indices():
    geo = oracle.partition_geometry()                               # (blocktype, ox, oy, w, h) per partition
    rng = np.random.default_rng(5)
    tiles = rng.integers(0, 5000, (9, 16)).astype(np.int64)         # [candidate][tile k = ty * 4 + tx]
    want = np.zeros((41, 9), np.int64)
    for p, (_, ox, oy, w, h) in enumerate(geo):
        for ty in range(oy // 4, (oy + h) // 4):
            for tx in range(ox // 4, (ox + w) // 4):
                want[p] += tiles[:, ty * 4 + tx]
    got = np.full((41, 9), -1, np.int64)
    lane = np.arange(32)
    k, half = lane & 15, lane >> 4
    tx, ty = k & 3, k >> 2
    for ps in range(5):
        c = 2 * ps + half
        v = np.where(c < 9, tiles[np.minimum(c, 8), k], 0)
        h84 = v + _shfl_xor(v, 1)
        v48 = v + _shfl_xor(v, 4)
        e88 = h84 + _shfl_xor(h84, 4)
        s168 = e88 + _shfl_xor(e88, 2)
        s816 = e88 + _shfl_xor(e88, 8)
        s1616 = s168 + _shfl_xor(s168, 8)
        for l in range(32):
            if c[l] >= 9:
                continue
            got[25 + k[l], c[l]] = v[l]
            if not tx[l] & 1: got[9 + ty[l] * 2 + (tx[l] >> 1), c[l]] = h84[l]
            if not ty[l] & 1: got[17 + (ty[l] >> 1) * 4 + tx[l], c[l]] = v48[l]
            if not tx[l] & 1 and not ty[l] & 1: got[5 + (ty[l] >> 1) * 2 + (tx[l] >> 1), c[l]] = e88[l]
            if tx[l] == 0 and not ty[l] & 1: got[1 + (ty[l] >> 1), c[l]] = s168[l]
            if ty[l] == 0 and not tx[l] & 1: got[3 + (tx[l] >> 1), c[l]] = s816[l]
            if k[l] == 0: got[0, c[l]] = s1616[l]
    assert (got == want).all()
