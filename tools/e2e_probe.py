"""Host-pointer call times of the bench step (development): b2me_set_cur / b2me_set_ref / b2me_search_frame, pinned buffers."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from h264_b200 import api
fr, pred, cen = bench.workload(seed=1)
s = api.Searcher(bench.W, bench.H, bench.NREFS, bench.R)
L = s.L
pin = lambda a: torch.from_numpy(a).pin_memory()
h_fr = [pin(fr[j]) for j in range(bench.NREFS + 1)]
h_pred, h_cen = pin(pred), pin(cen)
n = s.nmb * bench.NREFS * 41
h_mvs = torch.zeros((n, 2), dtype=torch.int16).pin_memory(); h_cs = torch.zeros((n,), dtype=torch.int64).pin_memory()
p = api.make_params(bench.LAMBDA)
for r in range(bench.NREFS):
    L.b2me_set_ref(s.h, r, C.c_void_p(h_fr[bench.NREFS - 1 - r].data_ptr()), bench.W)
t = [0.0, 0.0, 0.0]
N = 20
for i in range(-3, N):
    t0 = time.perf_counter()
    L.b2me_set_cur(s.h, C.c_void_p(h_fr[bench.NREFS].data_ptr()), bench.W)
    t1 = time.perf_counter()
    L.b2me_set_ref(s.h, i % bench.NREFS, C.c_void_p(h_fr[bench.NREFS - 1 - (i % bench.NREFS)].data_ptr()), bench.W)
    t2 = time.perf_counter()
    r = L.b2me_search_frame(s.h, C.c_void_p(h_pred.data_ptr()), C.c_void_p(h_cen.data_ptr()), C.byref(p), None, None,
                            C.c_void_p(h_mvs.data_ptr()), C.c_void_p(h_cs.data_ptr()))
    t3 = time.perf_counter()
    assert r == 0
    if i >= 0:
        t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2
print(f"BANDS={os.environ.get('B2ME_BANDS', 'default')}  set_cur {1e3 * t[0] / N:.3f}  set_ref {1e3 * t[1] / N:.3f}  search_frame {1e3 * t[2] / N:.3f} ms", flush=True)
