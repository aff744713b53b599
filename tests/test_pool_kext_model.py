"""Host-side model (numpy, no GPU) of the next step planned for k_frac_pool: moving the
-Sr*Sd/64 correction of the fractal fit into the u8 x u8 -> s32 contraction by extra K columns, so that the epilogue's
filter needs no per-output FFMA and no column constants.  The test pins the algebra the kernel will rely on:

  acc' = sum_k r_k d_k + 64 * Mr * Md' + Mr * Fd' + Fr * Md'        (all operands u8; Mr = Sr >> 6, Fr = Sr & 63,
                                                                      Md' = 255 - (Sd >> 6), Fd' = 63 - (Sd & 63))
       = num / 64 + RowConst(r) + Fr * Fd / 64,   RowConst = (64 * 255 + 63) * Mr + 255 * Fr,   0 <= Fr * Fd / 64 < 63

so |num / 64| <= |acc' - RowConst| + 63 (a conservative filter bound from raw accumulators), and the exact cross term is
recovered in integers: sum r d = acc' - (64 * Mr * Md' + Mr * Fd' + Fr * Md').  K = 64 + 64 + 2 = 130 columns (160 padded:
five kind::i8 K steps); acc' < 2^23."""
import numpy as np


def extended_operands(r, d):
    sr, sd = int(r.sum()), int(d.sum())
    Mr, Fr, Md, Fd = sr >> 6, sr & 63, sd >> 6, sd & 63
    a = np.concatenate([r.astype(np.int64), np.full(64, Mr), [Mr], [Fr]])
    b = np.concatenate([d.astype(np.int64), np.full(64, 255 - Md), [63 - Fd], [255 - Md]])
    assert a.min() >= 0 and a.max() <= 255 and b.min() >= 0 and b.max() <= 255      # u8 operands
    return a, b, (64 * 255 + 63) * Mr + 255 * Fr, 64 * Mr * (255 - Md) + Mr * (63 - Fd) + Fr * (255 - Md)


def test_extended_contraction_bounds_and_recovers_num():
    rng = np.random.default_rng(7)
    blocks = [rng.integers(0, 256, 64) for _ in range(300)]
    blocks += [np.zeros(64, np.int64), np.full(64, 255), np.arange(64) * 4, np.full(64, 63), np.full(64, 64)]
    worst = 0
    for i in range(0, len(blocks) - 1):
        r, d = blocks[i], blocks[(i * 7 + 3) % len(blocks)]
        a, b, rowc, extra = extended_operands(r, d)
        acc = int((a * b).sum())
        srd = int((r.astype(np.int64) * d).sum())
        num = 64 * srd - int(r.sum()) * int(d.sum())
        assert 0 <= acc < 2 ** 23
        assert acc - extra == srd                                   # exact recovery of the cross term
        eps = (acc - rowc) - num / 64.0
        assert 0 <= eps < 63
        assert abs(num) / 64.0 <= abs(acc - rowc) + 63              # the filter's bound is conservative
        worst = max(worst, eps)
    assert worst > 1                                                # the slack is real: the margin is needed
