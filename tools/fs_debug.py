"""development: integer search vs the oracle on a small picture; prints which partitions / items differ"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from h264_b200 import api, synth
W, H, R, NR = 64, 48, int(os.environ.get("R", 16)), 2
fr = synth.luma_sequence(W, H, NR + 1, seed=1)
cur, refs = fr[NR], fr[[1, 0]]
s = api.Searcher(W, H, NR, R)
s.set_cur(cur)
for r in range(NR):
    s.set_ref(r, refs[r])
pred, cen = synth.predictors(W, H, NR, seed=2, spread=int(os.environ.get("SPREAD", 0)), rmax=6)
lam = (187, 187, 187)
got = s.search_frame(pred, cen, api.make_params(lam, do_subpel=False))
exp = oracle.OrcFrame(cur, refs, R).search_frame(pred, cen, lam)
bad = np.argwhere(got[1] != exp[1])
print("stats", s.search_stats())
print("cost mismatches", len(bad), "of", got[1].size)
import collections
print("by partition", sorted(collections.Counter(bad[:, 2].tolist()).items()))
for mb, r, p in bad[:12]:
    print(mb, r, p, "got", got[0][mb, r, p], got[1][mb, r, p], "exp", exp[0][mb, r, p], exp[1][mb, r, p], "pred", pred[mb, r, p], "cen", cen[mb, r, p])
