"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/v1_cascade.npz from the UNMODIFIED reference: the TRANS_NODE tree
encode_one_macroblock (V1/src/block_enc.c:508) leaves for every macroblock of seeded synthetic frames (driven through
oracle/v1_harness.c) and the planes decode_one_macroblock (V1/src/block_dec.c:20) reconstructs from them, luma and chroma, with the four plane sets C/H/M/N loaded (case "loaded") and with H/M/N left
as the shipped program leaves them, all zero (case "zero").  Tolerances are chosen so that every outcome of the
cascade occurs: 16x16 kept, 8x8 kept, 8x4 / 4x8 accepted, 4x4, winners from every plane set.

Run in the build container (needs /root/reference):  python oracle/gen_golden_cascade.py [case]
"""
import os
import subprocess
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from h264_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = {   # name: W, H, R, seed, gain, offset, noise, tol, H/M/N loaded
    "loaded": (96, 64, 4, 5, 1.0, 0.0, 5.0, (4.0, 5.0, 2.0), True),
    "zero": (96, 64, 5, 4, 0.9, 10.0, 4.0, (3.0, 2.5, 2.0), False),
}


def inputs(name):
    """(current planes, [4 plane sets of (y, u, v)], W, H, R, tol, loaded) of a case -- needs no reference"""
    W, H, R, seed, gain, offset, noise, tol, loaded = CASES[name]
    ref, cur = synth.yuv_pair(W, H, seed=seed, shift=(0, 0), gain=gain, offset=offset, noise=noise)
    rng = np.random.default_rng(5)
    sets = [ref]
    for s in range(1, 4):
        if loaded:
            sets.append([np.clip(np.roll(p.astype(np.int32), (s, -s), (0, 1)) + rng.integers(-3, 4, p.shape), 0, 255).astype(np.uint8) for p in ref])
        else:
            sets.append([np.zeros_like(p) for p in ref])
    return cur, sets, W, H, R, tol, loaded


def node_rows(a, b):
    n = np.zeros(len(a), oracle.V1_NODE)
    n["block_type"], n["partition"], n["reference"], n["x"], n["y"] = a.T
    n["scale"], n["offset"] = b.T
    return n


def run_case(name):
    cur, sets, W, H, R, tol, loaded = inputs(name)
    v = oracle.V1Ref(W, H, R, tol=tol)
    for s in range(4):
        if s == 0 or loaded:
            v.set_ref(s, *sets[s], build_sums=True)
    v.set_cur(*cur)
    out = {}
    for con in (1, 2, 3):
        v.reset_trans()
        nmb = (W // 16) * (H // 16) if con == 1 else (W // 32) * (H // 32)
        out[f"nodes_{con}"] = np.stack([node_rows(*v.encode_mb(mb, con)) for mb in range(nmb)])
        out[f"rec_{con}"] = v.decode_plane(con)        # decode_one_macroblock (V1/src/block_dec.c:20) on those trees
    np.savez_compressed(os.path.join(GOLD, f"v1_cascade_{name}.npz"), **out)
    n1 = out["nodes_1"]
    print(name, "root partitions", np.bincount(n1[:, 0]["partition"], minlength=4), "8x8 partitions",
          np.bincount(n1[:, [1, 6, 11, 16]]["partition"].ravel(), minlength=4), "references", np.bincount(n1["reference"].ravel(), minlength=4))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for n in CASES:   # one process per case: the reference keeps its state in globals
            subprocess.check_call([sys.executable, os.path.abspath(__file__), n])
