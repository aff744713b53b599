"""Profiling target: one pool search (1080p ranges, ND domains), small enough for ncu replays."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h264_b200 import api, synth
W, H, ND = 1920, 1080, int(os.environ.get("ND", 16384))
fr = synth.luma_sequence(W, H, 2, seed=3)
s = api.PoolSearcher(W, H, W, H, ND)
s.set_planes(fr[1], fr[0])
res = s.search()
print("ok", int(res[0].astype("int64").sum()))
