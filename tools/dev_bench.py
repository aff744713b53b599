"""Developer probe (GPU box): instruction micro-benchmarks + per-kernel times at 1080p."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from h264_b200 import api, synth

out = {}
names = ["vabsdiff4", "vabsdiff4+imad", "vabsdiff4+iadd3", "vabsdiff4+lop3", "iadd3", "imad", "vimnmx16x2", "vabsdiff4+lds", "v+shf", "v+prmt", "v+lop3i", "v+vimnmx", "v+lds32", "v+iadd", "v+imad", "v+viaddmnmx", "v+idp2a", "v+lds64", "v+imadpack", "viaddmnmx", "v+vimnmx3_16x2", "v+viadd16x2", "v+mad1", "v+vimnmx_s16x2", "vabsdiff4_nullified", "vabsdiff4+nullified"]
for k, n in (enumerate(names) if os.environ.get("UBENCH") else []):
    if os.environ.get("UBENCH") != "1" and n not in os.environ["UBENCH"].split(","):
        continue
    out[n] = api.ubench(k, 4000)
    print(f"ubench {n}: {out[n]:.1f} G lane-ops/s", flush=True)

W, H, R, NR = 1920, 1088, int(os.environ.get("R", 32)), int(os.environ.get("NR", 4))
fr = synth.luma_sequence(W, H, NR + 1, seed=1)
s = api.Searcher(W, H, NR, R)
s.set_cur(fr[NR])
for r in range(NR):
    s.set_ref(r, fr[NR - 1 - r])
base = np.zeros((s.nmb, NR, 1, 2), np.int64)
for r in range(NR):
    base[:, r, 0, 0] = 4 * 2 * (r + 1)
    base[:, r, 0, 1] = 4 * 1 * (r + 1)
for spread in (0, 3):
    pred, cen = synth.predictors(W, H, NR, seed=1, spread=spread, base=base)
    p = api.make_params((187, 187, 187))
    s.kernel_timing(True); s.search_stats()
    for it in range(3):
        t = time.time(); res = s.search_frame(pred, cen, p); dt = time.time() - t
    ms0, n0 = s.kernel_time_ms(0); ms2, n2 = s.kernel_time_ms(2)
    pel = s.nmb * NR * (2 * R + 1) ** 2 * 256
    print(f"spread={spread}: int search {ms0/n0:.3f} ms/launch -> {pel/(ms0/n0*1e-3)/1e12:.2f} Tpel-sp/s; subpel {ms2/n2:.3f} ms; host call {dt*1e3:.1f} ms", flush=True)
    print('  stats', s.search_stats(), flush=True)
    mv = res[0]
    print("  mv_int mode:", np.unique(mv.reshape(-1, 2), axis=0, return_counts=True)[0][:3], flush=True)
    s.kernel_timing(False)
json.dump(out, open(os.path.join("gpurun_out", "ubench.json"), "w"))
