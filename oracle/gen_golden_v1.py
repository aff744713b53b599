"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/v1_harness_*.npz from the UNMODIFIED version1
sources (oracle/_ref/libv1ref.so = V1/src/compute.c + V1/src/block_enc.c + oracle/v1_harness.c):
full_search results (x, y, scale, offset, rms) of every range block of every level on seeded
synthetic 4:2:0 frame pairs (h264_b200.synth.yuv_pair), for plane set C (sum tables built) and
plane set H (zero plane, tables never built -- what the shipped program searches, SURVEY Q-F3),
plus the reference's sum tables and the partition cascade of encode_one_macroblock.

Run in the build container (needs /root/reference):  python oracle/gen_golden_v1.py
The reference keeps its state in globals, so each geometry runs in its own process.
"""
import os
import subprocess
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (W, H, R, seed, shift, gain, offset, planes to search)
    "cif": (352, 288, 7, 20261018, (-3, 2), 0.9, 10.0, ((0, 1),)),                       # BASELINE config 2 geometry, luma / C
    "qcif": (176, 144, 7, 7, (2, -1), 1.1, -6.0, ((0, 1), (1, 1), (0, 2), (0, 3), (1, 2))),  # luma+chroma, C and H sets
    "small": (64, 48, 5, 3, (1, 1), 0.7, 30.0, ((0, 1), (1, 1), (0, 2))),                 # every block touches a border
}


def run_case(name):
    import oracle
    from h264_b200 import synth
    W, H, R, seed, shift, gain, offset, planes = CASES[name]
    ref, cur = synth.yuv_pair(W, H, seed=seed, shift=shift, gain=gain, offset=offset)
    v = oracle.V1Ref(W, H, R)
    v.set_ref(0, *ref, build_sums=True)            # C: previous reconstructed frame (code.c:256-257)
    v.set_cur(*cur)                                # image.c:458
    out = {"cfg": np.array([W, H, R, seed, shift[0], shift[1]], np.int64), "gain_offset": np.array([gain, offset])}
    for which, con in planes:
        xy, so, rms = v.search_plane(which, con)
        out[f"xy_{which}_{con}"], out[f"so_{which}_{con}"], out[f"rms_{which}_{con}"] = xy, so, rms
    if name == "small":
        for sz in range(7):
            for con in (1, 2):
                for sq in (0, 1):
                    out[f"tab_{sz}_{con}_{sq}"] = v.domain_table(sz, con, sq)
        nodes, nodes_d = [], []
        for mb in range((W // 16) * (H // 16)):
            a, b = v.encode_mb(mb, 1)
            nodes.append(a); nodes_d.append(b)
        out["cascade_i"], out["cascade_d"] = np.array(nodes), np.array(nodes_d)
    np.savez_compressed(os.path.join(GOLD, f"v1_harness_{name}.npz"), **out)
    print(name, {k: getattr(x, "shape", None) for k, x in out.items() if k.startswith(("xy", "casc"))})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for n in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), n])
