"""TEST INFRASTRUCTURE ONLY.  tests/golden/jm_tqc.npz: the UNMODIFIED JM residual_transform_quant_chroma_4x4
(JM/lencod/src/block.c:953-1200; oracle/_ref/libjmref.so through oracle/jm_harness_tq.c) on seeded 8x8 chroma blocks of 4:2:0
macroblocks: DC / AC level and run lists, reconstruction, cr_cbp, at several chroma QPs, intra and inter quantisers, both planes,
CAVLC and CABAC.  Run in the build container (needs /root/reference):  python oracle/gen_golden_tqc.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
CASES = [(0, 0, 1, 0, 51), (14, 0, 0, 1, 52), (26, 0, 1, 0, 53), (26, 1, 0, 1, 54), (33, 0, 0, 0, 55), (39, 1, 1, 1, 56)]   # (qpc, symbol_mode, intra, uv, seed)
NMB = 400


def chroma_blocks(nmb, seed):
    """(orig, pred) [nmb][64]: flat offsets (DC only), faint textures (AC levels of 1 that the cost rule drops), strong textures,
    identical blocks (nothing coded), saturated blocks"""
    rng = np.random.default_rng(seed)
    pred = rng.integers(0, 256, (nmb, 1)) + rng.integers(-3, 4, (nmb, 64))
    amp = rng.choice([0, 1, 2, 4, 12, 60], (nmb, 1))
    orig = pred + rng.integers(-12, 13, (nmb, 1)) * rng.choice([0, 1], (nmb, 1)) + rng.integers(-1, 2, (nmb, 64)) * amp
    orig = orig + (rng.random((nmb, 64)) < 0.05) * rng.integers(-40, 41, (nmb, 64))
    orig[::11] = pred[::11]
    orig[::13] = 255; pred[::13] = 0
    return np.clip(orig, 0, 255).astype(np.uint8), np.clip(pred, 0, 255).astype(np.uint8)


if __name__ == "__main__":
    import oracle
    out = {"cases": np.array(CASES, np.int32)}
    for ci, (qp, sm, intra, uv, seed) in enumerate(CASES):
        r = oracle.JMQuantRef(2 if intra else 0, sm)
        orig, pred = chroma_blocks(NMB, seed)
        res = r.tq_chroma(qp, intra, uv, orig, pred)
        for a, name in zip(res, ("dc_level", "dc_run", "ac_level", "ac_run", "recon", "cr_cbp")):
            out[f"c{ci}_{name}"] = a
        out[f"c{ci}_params"] = r.params_chroma(uv + 1, qp, intra)
        print(f"qpc {qp} symbol_mode {sm} intra {intra} uv {uv}: cr_cbp histogram {np.bincount(res[5], minlength=3).tolist()}")
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "jm_tqc.npz"), **out)
