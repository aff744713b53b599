"""k_sad_fs on bench.py's robustness cases (development probe, GPU box): kernel time, fraction of the measured
VABSDIFF4 peak, survivors and centre groups per item."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from h264_b200 import api
peak = api.ubench(0, 4000) * 4 / 1e3
out = bench.robustness_block(0, peak)
print(f"VABSDIFF4 peak {peak:.2f} Tpel-sp/s")
for k, v in out.items():
    print(k, json.dumps(v))
