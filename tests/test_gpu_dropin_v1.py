"""Drop-in test of the version1 boundary: the unmodified version1 objects (compute.c, block_enc.c, block_dec.c)
with `full_search` replaced by integration/v1/b2fr_v1_shim.c + libb2me.so (oracle/_ref/libv1b2.so, built by
oracle/Makefile.v1 in the build container).  The reference's own encode_one_macroblock then drives its cascade on GPU
searches and must leave the TRANS_NODE trees and the reconstruction of the stock code (golden: tests/golden/v1_cascade_*.npz,
captured from the unmodified program)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
have = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libv1b2.so"))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import oracle
from oracle import gen_golden_cascade as gc
name = sys.argv[1]
g = np.load(%r %% name)
cur, sets, W, H, R, tol, loaded = gc.inputs(name)
v = oracle.V1DropIn(W, H, R, tol=tol)
for s in range(4):
    if s == 0 or loaded:
        v.set_ref(s, *sets[s], build_sums=True)
v.set_cur(*cur)
v.new_frame([1, loaded, loaded, loaded])
for con in (1, 2, 3):
    v.reset_trans()
    nmb = (W // 16) * (H // 16) if con == 1 else (W // 32) * (H // 32)
    nodes = np.stack([gc.node_rows(*v.encode_mb(mb, con)) for mb in range(nmb)])
    exp = g["nodes_%%d" %% con]
    for f in ("block_type", "partition", "reference", "x", "y", "scale", "offset"):
        assert (nodes[f] == exp[f]).all(), (con, f)
    assert (v.decode_plane(con) == g["rec_%%d" %% con]).all(), con
assert v.calls() > 1000
print("ok", v.calls())
"""


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref/libv1b2.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", ["loaded", "zero"])
def test_version1_cascade_on_cuda_full_search_is_identical(name):
    # one process per case: the reference keeps its state in globals
    code = CHILD % (ROOT, os.path.join(ROOT, "tests", "golden", "v1_cascade_%s.npz"))
    r = subprocess.run([sys.executable, "-c", code, name], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout[-1500:], r.stderr[-1500:])
