#!/usr/bin/env python
"""bench.py -- headline benchmark of the block-matching hot path (BASELINE.json metric).

Workload (config.workload): BASELINE config "1080p 4:2:0, JM full-search +-32 integer + quarter-pel
SATD refinement, 4 reference frames".  One step = the motion search of one P frame:
  * build the 16 quarter-pel planes of the newest reference (getSubImagesLuma),
  * integer full search +-32 (SAD) for all 8160 MBs x 4 refs x 41 partitions,
  * half/quarter-pel SATD refinement of every partition.
metric = Mpel-search-points/s = MBs x refs x (2R+1)^2 x 256 / time (the
SAD-tree-algorithmic unit); MB/s is reported beside it.

  value : inputs already resident in HBM (device pointers through the C ABI's *_dev calls)
  e2e   : the same step through the host-pointer C ABI (b2me_set_cur / b2me_set_ref /
          b2me_search_frame_best) with pinned HOST buffers, H2D + D2H inside the timed region; two independent
          segment streams per GPU (own context + host thread each), the single-stream figure beside it.

Multi-GPU (torchrun, one rank per GPU): independent closed-GOP segments, i.e. every rank runs the
same per-frame step on its own frames -- no data-path collective (weak scaling).

--impl reference: times the reference's own CPU implementation (the unmodified JM objects in
oracle/_ref/libjmref.so, else the restated oracle) on a bounded MB sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W, H, R, NREFS = 1920, 1088, 32, 4
LAMBDA = (187, 187, 187)        # LAMBDA_FACTOR(lambda_me) at QP 28 (JM lambda.c:30, defines.h:130)
UNIT = "Mpel-search-points/s"
# one string for both arms (the driver compares config.workload of the product arm and of --impl reference)
WORKLOAD = (f"1080p {W}x{H} luma, JM full search +-{R} integer (SAD) + half/quarter-pel SATD refinement, {NREFS} refs, "
            f"41 partitions/MB")


def workload(seed=1):
    from h264_b200 import synth
    fr = synth.luma_sequence(W, H, NREFS + 2, seed=seed)
    nmb = (W // 16) * (H // 16)
    # predictor = the clip's true pan motion per reference distance (what JM's median predictor
    # converges to on a panning clip), shared by the 41 partitions of an MB
    base = np.zeros((nmb, NREFS, 1, 2), np.int64)
    for r in range(NREFS):
        base[:, r, 0, 0] = 4 * 2 * (r + 1)
        base[:, r, 0, 1] = 4 * 1 * (r + 1)
    pred, cen = synth.predictors(W, H, NREFS, seed=seed, spread=0, base=base)
    return fr, pred, cen


def pel_sp(nmb):
    return nmb * NREFS * (2 * R + 1) ** 2 * 256


def robustness_cases(seed=1):
    """The headline frame under less friendly inputs: what full_search_motion_estimation really
    receives.  Each case = (name, frames, pred, centre).  pan: the headline's own predictors (true pan motion, shared
    by the 41 partitions); perpart: one predictor per PARTITION (true motion + up to +-3 quarter-pel, the spread JM's
    median prediction shows on jm_wrap_foreman.npz) so that centres differ inside an MB; off8: predictor wrong by up
    to +-8 pel per (MB, ref); noise: uniform random pictures (no candidate is good, every bound converges slowly)."""
    from h264_b200 import synth
    fr, pred, cen = workload(seed)
    nmb = (W // 16) * (H // 16)
    rng = np.random.default_rng(77)
    cases = [("pan", fr, pred, cen)]
    pp = (pred.astype(np.int64) + rng.integers(-3, 4, pred.shape)).astype(np.int16)
    cases.append(("perpart", fr, pp, (((pp.astype(np.int32) + 2) >> 2) * 4).astype(np.int16)))
    po = (pred.astype(np.int64) + 4 * rng.integers(-8, 9, (nmb, NREFS, 1, 2))).astype(np.int16)
    cases.append(("off8", fr, po, (((po.astype(np.int32) + 2) >> 2) * 4).astype(np.int16)))
    cases.append(("noise", rng.integers(0, 256, fr.shape).astype(np.uint8), pred, cen))
    return cases


def robustness_block(device, sad_peak_tpel, iters=3):
    """k_sad_fs alone (integer stage, resident inputs, CUDA events) on every robustness case."""
    import torch
    from h264_b200 import api
    dev = torch.device("cuda", device)
    nmb = (W // 16) * (H // 16)
    s = api.Searcher(W, H, NREFS, R, device=device)
    p = api.make_params(LAMBDA, do_subpel=False)
    mvi = torch.zeros((nmb, NREFS, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
    ci = torch.zeros((nmb, NREFS, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
    out = {}
    for name, fr, pred, cen in robustness_cases():
        d_fr = torch.from_numpy(fr).to(dev)
        s.set_cur_dev(d_fr[NREFS])
        for r in range(NREFS):
            s.set_ref_dev(r, d_fr[NREFS - 1 - r])
        dp, dc = torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev)
        s.search_frame_dev(dp, dc, p, mvi, ci, mvs, cs)
        torch.cuda.synchronize()
        s.search_stats(); s.kernel_timing(True)
        for _ in range(iters):
            s.search_frame_dev(dp, dc, p, mvi, ci, mvs, cs)
        torch.cuda.synchronize()
        ms, n = s.kernel_time_ms(0)
        st = s.search_stats(); s.kernel_timing(False)
        t = pel_sp(nmb) / (ms / n * 1e-3) / 1e12
        out[name] = {"kernel_ms": ms / n, "frac": t / sad_peak_tpel, "survivors_per_item": st["exact_evals"] / max(st["items"], 1),
                     "centre_groups_per_item": st["window_passes"] / max(st["items"], 1),
                     "mv_checksum": int(mvi.to(torch.int64).sum().item()), "cost_checksum": int(ci.sum().item())}
    s.close()
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML (nvidia_ml_py) is polled
    from a thread every ~2 ms (nvidia-smi -lms cannot resolve a region of tens of milliseconds)."""

    def __init__(self, gpu):
        self.sm, self.reasons, self.stop, self.mx, self.err = [], set(), False, None, None
        try:
            import pynvml as N
            N.nvmlInit()
            # NVML indices are not remapped by CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu]) if vis and vis.split(",")[gpu].isdigit() else gpu
            self.N, self.h = N, N.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err, self.t = repr(e), None

    def _poll(self):
        N = self.N
        names = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown if hasattr(N, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.002)

    def finish(self):
        self.stop = True
        if self.t:
            self.t.join(timeout=1.0)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def pool_leg(device, peaks, nd=16384):
    """BASELINE config 5 (secondary line): 1080p range plane (32 400 8x8 ranges x 8 isometries) against a pool of
    `nd` 2:1-averaged 16x16 domain blocks; cross terms on tcgen05.mma kind::i8.  Tensor roofline: algorithmic
    ops 2*8*Nr*Nd*64 over the kernel's CUDA-event time, against 2x the measured dense bf16 rate (the nominal
    int8:bf16 ratio; MEASURED_PEAKS.json has no int8 figure) and against a tensor-only pass of the same kernel."""
    from h264_b200 import api, synth
    Wp, Hp = 1920, 1080
    fr = synth.luma_sequence(Wp, Hp, 2, seed=3)
    s = api.PoolSearcher(Wp, Hp, Wp, Hp, nd, device=device)
    s.set_planes(fr[1], fr[0])
    s.search(); s.kernel_time_ms(); s.stats()
    for _ in range(5):
        res = s.search()
    ms, n = s.kernel_time_ms()
    st = s.stats()
    probe = min(s.probe_ms() for _ in range(3))
    ops = 2.0 * 8 * s.nr * nd * 64
    peak = api.ubench_i8(4000, device=device)            # measured: dense tcgen05.mma kind::i8, resident operands, no epilogue
    t = ms / n * 1e-3
    out = {"workload": f"{Wp}x{Hp} range plane: {s.nr} 8x8 ranges x 8 isometries vs {nd} pooled domain blocks (K = 64 useful of 160 executed)",
           "kernel": "k_frac_pool (tcgen05.mma kind::i8 + fused least-squares fit / argmin epilogue)",
           "kernel_ms": ms / n, "pairs_per_s": 8.0 * s.nr * nd / t,
           "roofline": {"bound": "tensor", "achieved": ops / t / 1e12, "peak": peak, "unit": "Tops (int8 dense)",
                        "frac": ops / t / 1e12 / peak, "peak_source": "measured live: b2fp_ubench_i8 (tcgen05.mma kind::i8 M128 N256 K32 back to back on every SM)",
                        "executed_frac": ops * 160 / 64 / t / 1e12 / peak,
                        "tensor_only_ms": probe, "tensor_only_Tops": ops / (probe * 1e-3) / 1e12,
                        "note": "useful ops = 2*64 per (range-isometry, domain) pair; the contraction executes K = 160 (66 correction columns move the "
                                "-Sr*Sd/64 term of the fit into the MMA so that the epilogue filters raw accumulators), executed_frac counts those; "
                                "the per-output epilogue on the ALU pipe (32 three-input min/max per 32 outputs) and the re-examinations, not the MMA, bound the kernel (DESIGN.md 4.2)"},
           "exact_fits_per_row": st["exact_evals"] / (5 * 8 * s.nr), "best_dom_checksum": int(res[0].astype(np.int64).sum())}
    s.close()
    return out


def fractal_leg(device, with_cpu):
    """BASELINE config 2 (secondary line): one version1 P frame, CIF 352x288 4:2:0, Search_Range 7, the four plane sets
    searched for every range block of every level, the partition cascade and the prediction of all three components --
    b2fr_set_domain / b2fr_set_range / b2fr_encode_plane / b2fr_decode_plane through the host-pointer C ABI (copies
    inside the timed region).  with_cpu: the unmodified version1 code (oracle/_ref/libv1ref.so: compute_domain_Sum,
    compute_range_Sum, encode_one_macroblock, decode_one_macroblock) on ONE host core on the same frame, and a byte
    comparison of the two reconstructions (the cpu_baseline leg is the one place bench.py may run oracle/)."""
    from h264_b200 import api, synth
    Wc, Hc, Rc, tol = 352, 288, 7, (3.5, 4.5, 2.0)
    ref, cur = synth.yuv_pair(Wc, Hc, seed=3, shift=(0, 0), gain=1.0, offset=0.0, noise=4.0)
    f = api.FractalSearcher(Wc, Hc, Rc, device=device)

    def frame():
        f.set_domain(0, *ref, build_sums=True)
        f.set_range(*cur)
        return [(f.encode_plane(con, tol), f.decode_plane(con)) for con in (1, 2, 3)]
    for _ in range(3):
        out = frame()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        out = frame()
    ms = (time.perf_counter() - t0) * 1e3 / n
    nmb = (Wc // 16) * (Hc // 16) + 2 * (Wc // 32) * (Hc // 32)
    res = {"workload": f"{Wc}x{Hc} 4:2:0, Search_Range {Rc}, 41 range blocks/MB x 4 plane sets, cascade + prediction, Y+U+V",
           "kernels": "k_frac_domain_sums, k_frac_range_sums, k_frac_window, k_frac_decide, k_frac_predict",
           "ms_per_frame_host_api": ms, "mb_per_s": nmb / (ms * 1e-3), "split_macroblocks": int((out[0][0][:, 0]["partition"] != 0).sum()),
           "rec_checksum": int(sum(int(o[1].astype(np.int64).sum()) for o in out)), "launches_per_frame": None}
    l0 = f.launch_count(); frame(); res["launches_per_frame"] = int(f.launch_count() - l0)
    if with_cpu:
        import oracle
        if oracle.have_v1ref():
            v = oracle.V1Ref(Wc, Hc, Rc, tol=tol)
            t0 = time.perf_counter()
            v.set_ref(0, *ref, build_sums=True)
            v.set_cur(*cur)
            same = True
            for con in (1, 2, 3):
                v.reset_trans()
                for mb in range((Wc // 16) * (Hc // 16) if con == 1 else (Wc // 32) * (Hc // 32)):
                    v.encode_mb(mb, con)
                same = same and bool((v.decode_plane(con) == out[con - 1][1]).all())
            cpu_ms = (time.perf_counter() - t0) * 1e3
            res["cpu_reference"] = {"ms_per_frame": cpu_ms, "cores": 1, "kind": "reference",
                                    "note": "unmodified version1 compute.c / block_enc.c / block_dec.c, gcc -O2; it searches the "
                                            "plane sets lazily (only the blocks its cascade reaches), the GPU path all of them",
                                    "reconstruction_identical": same}
    f.close()
    return res


def recon_leg(device, steps=10):
    """The reconstruction side of a 1080p 4:2:0 picture on the device (secondary line): prediction of every macroblock from
    two lists (k_mc_mb), residual transform + quantisation of all luma 4x4 blocks and of both chroma planes (k_tq4x4, k_tq_chroma),
    the in-loop deblocking filter (k_deblock, 2:1 macroblock wavefront).  Device-resident synthetic inputs; CUDA events."""
    import torch
    from h264_b200 import api, synth
    dev = torch.device("cuda", device)
    Wr, Hr, NRr = 1920, 1088, 2
    rng = np.random.default_rng(21)
    fr = synth.luma_sequence(Wr, Hr, NRr + 1, seed=5)
    s = api.Searcher(Wr, Hr, NRr, 16)
    s.set_cur(fr[NRr])
    cu = rng.integers(90, 160, (Hr // 2, Wr // 2)).astype(np.uint8)
    s.set_cur_chroma(cu, cu)
    for r in range(NRr):
        s.set_ref(r, fr[NRr - 1 - r]); s.set_ref_chroma(r, np.roll(cu, r + 1, axis=1), np.roll(cu, r + 1, axis=0))
    nmb = s.nmb
    t = lambda a: torch.from_numpy(a).to(dev)
    mb_mode = t(rng.choice([1, 2, 3, 8], nmb).astype(np.uint8)); b8mode = t(rng.integers(4, 8, (nmb, 4)).astype(np.uint8))
    pdir = t(rng.integers(0, 3, (nmb, 4)).astype(np.uint8)); ref8 = t(rng.integers(0, NRr, (nmb, 2, 4)).astype(np.int8))
    mv = t(rng.integers(-24, 25, (nmb, NRr, 41, 2)).astype(np.int16))
    oy = torch.zeros((nmb * 16, 16), dtype=torch.uint8, device=dev); py = torch.zeros_like(oy)
    oc = torch.zeros((nmb, 2, 4, 16), dtype=torch.uint8, device=dev); pc = torch.zeros_like(oc)
    p4 = api.tq_default_params(4, 28, 0)
    lv = torch.zeros((nmb * 16, 16), dtype=torch.int16, device=dev); rn = torch.zeros((nmb * 16, 16), dtype=torch.uint8, device=dev)
    rec = torch.zeros_like(oy); cst = torch.zeros(nmb * 16, dtype=torch.int32, device=dev); nz = torch.zeros(nmb * 16, dtype=torch.uint8, device=dev)
    L = api.lib()
    c_o = torch.zeros((nmb * 2, 64), dtype=torch.uint8, device=dev); c_p = torch.zeros_like(c_o); c_r = torch.zeros_like(c_o)
    c_dl = torch.zeros((nmb * 2, 4), dtype=torch.int16, device=dev); c_dr = torch.zeros((nmb * 2, 4), dtype=torch.uint8, device=dev)
    c_al = torch.zeros((nmb * 2, 4, 16), dtype=torch.int16, device=dev); c_ar = torch.zeros((nmb * 2, 4, 16), dtype=torch.uint8, device=dev)
    c_cbp = torch.zeros(nmb * 2, dtype=torch.uint8, device=dev)
    yd = t(fr[NRr].copy()); ud = t(cu.copy()); vd = t(cu.copy())
    mbs = np.zeros(nmb, synth.DBK_MB); mbs["qp"] = rng.integers(24, 40, nmb); mbs["qpc_u"] = mbs["qpc_v"] = 30; mbs["intra"] = rng.random(nmb) < 0.1
    mbs["cbp_blk"] = np.where(rng.random(nmb) < 0.5, rng.integers(0, 65536, nmb), 0)
    blks = np.zeros((Wr // 4) * (Hr // 4), synth.DBK_BLK); blks["mv"] = rng.integers(-6, 7, (len(blks), 2, 2)); blks["ref"] = rng.integers(-1, 2, (len(blks), 2))
    d_mbs = t(mbs.view(np.uint8)); d_blks = t(blks.view(np.uint8)); prog = torch.zeros(Hr // 16, dtype=torch.int32, device=dev)
    import ctypes as C

    def run():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        s.mc_mb_dev(mb_mode, b8mode, pdir, ref8, mv, mv, oy, py, oc, pc)
        ev[1].record()
        api.tq_dev(p4, oy, py, 4, lv, rn, rec, cst, nz)
        r = L.b2tq_chroma_dev(C.byref(p4), C.c_int(nmb * 2), C.c_void_p(oc.data_ptr()), C.c_void_p(pc.data_ptr()), C.c_void_p(c_dl.data_ptr()), C.c_void_p(c_dr.data_ptr()),
                              C.c_void_p(c_al.data_ptr()), C.c_void_p(c_ar.data_ptr()), C.c_void_p(c_r.data_ptr()), C.c_void_p(c_cbp.data_ptr()), C.c_void_p(0))
        assert r == 0
        ev[2].record()
        api.deblock_frame_dev(yd, ud, vd, d_mbs, d_blks, prog)
        ev[3].record()
        return ev
    # warm-up by TIME, not by count: k_deblock is latency-bound, i.e. proportional to the SM clock, and this leg may follow a host-only
    # phase (the CPU halves of other legs) during which the GPU has clocked down; 3 passes of ~1 ms do not bring it back
    t_w = time.time()
    while time.time() - t_w < 0.25:
        run()
        torch.cuda.synchronize()
    evs = [run() for _ in range(steps)]
    torch.cuda.synchronize()
    ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / steps for i in range(3)]
    s.close()
    px = Wr * Hr * 3 // 2
    return {"workload": "1920x1088 4:2:0: prediction of every macroblock from two lists (luma + chroma), transform / quantisation of all luma 4x4 blocks and both chroma "
                        "planes, in-loop deblocking; synthetic modes / vectors, inputs resident",
            "ms_prediction": ms[0], "ms_transform_quant": ms[1], "ms_deblock": ms[2],
            "prediction_gbs": 4 * px / (ms[0] * 1e-3) / 1e9, "transform_quant_gbs": 3.3 * px / (ms[1] * 1e-3) / 1e9,
            "deblock_note": "k_deblock is latency-bound by construction: the macroblock order is a 2:1 wavefront (120 + 67 dependent steps at 1080p); one CTA per macroblock row (luma warp + chroma warp), boundary strengths by a picture-wide pre-pass (k_dbk_prep)",
            "kernels": "k_mc_mb, k_tq4x4, k_tq_chroma, k_dbk_prep, k_deblock"}


def bands_leg(local, rank, world, steps=5):
    """BASELINE config 4 (secondary line, every rank takes part): ONE 3840x2160 picture, full search +-64, 1 reference, split into
    `world` MB-row bands (h264_b200/bands.py).  Per picture, inside the timed region: halo exchange of the reconstructed
    reference rows with the neighbouring bands (NCCL point-to-point), the sub-pel / search planes of the rows the band reads,
    the search of the band (integer + sub-pel).  Strong scaling: the picture is fixed, the driver compares the N values."""
    import torch
    import torch.distributed as dist
    from h264_b200 import api, bands, synth
    dev = torch.device("cuda", local)
    Wb, Hb, Rb = 3840, 2160, 64
    fr = synth.luma_sequence(Wb, Hb, 2, seed=9)
    nmb = (Wb // 16) * (Hb // 16)
    pred, cen = synth.predictors(Wb, Hb, 1, seed=2, spread=0, base=np.tile(np.array([[[[8, 4]]]], np.int64), (nmb, 1, 1, 1)))
    b = bands.BandSearcher(Wb, Hb, 1, Rb, rank, world, device=local, max_center_pel=16)
    d_cur = torch.from_numpy(fr[1]).to(dev)
    own = torch.from_numpy(fr[0][16 * b.first_row:16 * b.last_row]).to(dev).contiguous()
    d_pred, d_cen = torch.from_numpy(pred).to(dev), torch.from_numpy(cen).to(dev)
    mvi = torch.zeros((nmb, 1, 41, 2), dtype=torch.int16, device=dev); mvs = torch.zeros_like(mvi)
    ci = torch.zeros((nmb, 1, 41), dtype=torch.int64, device=dev); cs = torch.zeros_like(ci)
    params = api.make_params(LAMBDA)
    stream = torch.cuda.current_stream().cuda_stream
    b.set_cur_dev(d_cur, stream)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        b.set_ref_from_band(0, own, stream)          # halo exchange (NCCL p2p) + planes of the rows this band reads
        e[1].record()
        b.search(d_pred, d_cen, params, mvi, ci, mvs, cs, stream, check=False)
        e[2].record()
        return e

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    b.search(d_pred, d_cen, params, mvi, ci, mvs, cs, stream, check=True)       # validates the centres against the halo once
    for _ in range(3):
        step()
    barrier()
    evs = []
    for i in range(steps):
        flush.fill_(i)
        evs.append(step())
    barrier()
    t = torch.tensor([sum(e[0].elapsed_time(e[2]) for e in evs) / steps, sum(e[0].elapsed_time(e[1]) for e in evs) / steps],
                     dtype=torch.float64, device=dev)
    # ---- config 4's EPZS half: the EPZS integer search (b2me_epzs_search_dev, one warp per block) of every partition of the
    #      band's macroblocks on the same planes.  Predictors = the encoder's usual sources taken from the motion field the full
    #      search above left (zero, the block's own predictor, the left / top / top-right macroblocks' vectors); extended-diamond
    #      refinement with the dual round.  Reported: time per picture, search points per block, agreement with the full search. ----
    mvi_h = mvi[:, 0].cpu().numpy()
    mbw = Wb // 16
    mbs = np.arange(b.mb_first, b.mb_first + b.mb_count)
    geom = np.array(synth.PART_GEOM, np.int64)
    nj = len(mbs) * 41
    jobs = np.zeros(nj, synth.EPZS_JOB)
    mbx, mby = np.repeat(mbs % mbw, 41), np.repeat(mbs // mbw, 41)
    pp = np.tile(np.arange(41), len(mbs))
    jobs["pos_x"] = 16 * mbx + geom[pp, 1]; jobs["pos_y"] = 16 * mby + geom[pp, 2]; jobs["blocktype"] = geom[pp, 0]
    jobs["mv"] = (8, 4); jobs["pred"] = (8, 4); jobs["range"] = (4 * Rb, 4 * Rb)
    jobs["mv_range"] = np.where(geom[pp, 0] < 5, 10, 12)
    jobs["flags"] = 4 | 8 | np.where(geom[pp, 0] > 4, 2, 0)
    jobs["lambda_factor"] = LAMBDA[0]
    jobs["medthres"] = (geom[pp, 3] * geom[pp, 4]) << 5
    jobs["stop0"] = jobs["medthres"] + 2 * LAMBDA[0]; jobs["stop"] = 2 * jobs["medthres"] + 2 * LAMBDA[0]
    jobs["pred_first"] = 5 * np.arange(nj); jobs["npred"][:, 0] = 5; jobs["cond_host"][:, 0] = 1
    jobs["pat_init"] = 2; jobs["pat_sd"] = 0; jobs["pat_sq"] = 1; jobs["pat_else"] = 2; jobs["pat_dual"] = 2
    nb = lambda dx, dy: mvi_h[np.clip(mby + dy, 0, Hb // 16 - 1) * mbw + np.clip(mbx + dx, 0, mbw - 1), pp]
    preds = np.stack([np.zeros((nj, 2), np.int16), np.tile(np.array([8, 4], np.int16), (nj, 1)), nb(-1, 0), nb(0, -1), nb(1, -1)], axis=1).astype(np.int16)
    d_jobs = torch.from_numpy(jobs.view(np.uint8)).to(dev); d_preds = torch.from_numpy(preds.reshape(-1, 2)).to(dev)
    d_pats = torch.from_numpy(synth.epzs_patterns().view(np.uint8)).to(dev)
    d_out = torch.zeros(nj * synth.EPZS_RESULT.itemsize, dtype=torch.uint8, device=dev)
    for _ in range(2):
        b.s.epzs_search_dev(nj, d_jobs, d_preds, 3, d_pats, d_out, stream)
    barrier()
    ee = []
    for i in range(steps):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.s.epzs_search_dev(nj, d_jobs, d_preds, 3, d_pats, d_out, stream); e1.record()
        ee.append((e0, e1))
    barrier()
    res = d_out.cpu().numpy().view(synth.EPZS_RESULT)
    same = int((res["mv"] == mvi_h[mbs].reshape(-1, 2)).all(axis=1).sum())
    et = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in ee) / steps], dtype=torch.float64, device=dev)
    ec = torch.tensor([nj, int(res["npoints"].sum()), same], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX); dist.all_reduce(ec)
    epzs = {"kernel": "k_epzs (b2me_epzs_search_dev: EPZS_motion_estimation / EPZS_subMB_motion_estimation on the device, one warp per block)",
            "blocks_per_picture": int(ec[0]), "ms_per_picture": float(et[0]), "search_points_per_block": float(ec[1]) / float(ec[0]),
            "blocks_per_s": float(ec[0]) / (float(et[0]) * 1e-3), "same_vector_as_full_search": float(ec[2]) / float(ec[0]),
            "predictors": "zero, the block's predictor, left / top / top-right macroblocks' vectors; extended diamond + dual refinement"}
    halo = sum((hi - lo) * Wb for s_, d_, lo, hi in bands.exchange_plan(world, Hb // 16, Rb) if d_ == rank)
    hb = torch.tensor([halo], dtype=torch.int64, device=dev)
    chk = mvs[b.mb_first:b.mb_first + b.mb_count].to(torch.int64).sum().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(hb); dist.all_reduce(chk)
    b.s.close()
    pel = nmb * (2 * Rb + 1) ** 2 * 256
    return {"workload": f"{Wb}x{Hb} single picture, full search +-{Rb} (SAD) + sub-pel SATD, 1 ref, {world} MB-row band(s), "
                        f"halo exchange of the reconstructed reference per picture (NCCL point-to-point)",
            "scaling": "strong", "ms_per_picture": float(t[0]), "value": pel / (float(t[0]) * 1e-3) / 1e6, "unit": UNIT,
            "ms_halo_exchange_and_planes": float(t[1]), "halo_bytes_all_ranks": int(hb.item()),
            "limiter": "the band's search; the exchange + plane rebuild share is ms_halo_exchange_and_planes / ms_per_picture",
            "mv_checksum_all_bands": int(chk.item()), "l2": "192 MB L2 flush between timed pictures", "epzs": epzs}


def pool_multi_leg(local, rank, world, nds=(16384, 65536), steps=4):
    """BASELINE config 5 across ranks (secondary line, every rank takes part): the 1080p range plane split into `world` bands
    of range-block rows, the domain plane replicated by ONE NCCL broadcast per picture (inside the timed region), each rank
    building the pool operands and searching its band (h264_b200/pool_bands.py).  Strong scaling."""
    import torch
    import torch.distributed as dist
    from h264_b200 import pool_bands, synth
    dev = torch.device("cuda", local)
    Wp, Hp = 1920, 1080
    fr = synth.luma_sequence(Wp, Hp, 2, seed=3)
    rp = torch.from_numpy(fr[1]).to(dev)
    dp = torch.from_numpy(fr[0]).to(dev) if rank == 0 else torch.zeros((Hp, Wp), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    for nd in nds:
        b = pool_bands.PoolBandSearcher(Wp, Hp, Wp, Hp, nd, rank, world, device=local)
        res = (torch.zeros(b.nr, dtype=torch.int32, device=dev), torch.zeros(b.nr, dtype=torch.uint8, device=dev),
               torch.zeros(b.nr, dtype=torch.int16, device=dev), torch.zeros(b.nr, dtype=torch.int16, device=dev),
               torch.zeros(b.nr, dtype=torch.int64, device=dev))
        for _ in range(3):
            b.search_dev(rp, dp, res, stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); b.search_dev(rp, dp, res, stream); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(c) for a, c in evs) / steps], dtype=torch.float64, device=dev)
        chk = res[0].to(torch.int64).sum().reshape(1)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(chk)
        ops = 2.0 * 8 * (Wp // 8) * (Hp // 8) * nd * 64
        out[str(nd)] = {"ms_per_picture": float(t[0]), "Tops_useful": ops / (float(t[0]) * 1e-3) / 1e12, "best_dom_checksum": int(chk.item())}
        b.close()
    return {"workload": f"{Wp}x{Hp} range plane ({(Wp // 8) * (Hp // 8)} 8x8 ranges x 8 isometries) in {world} band(s) of range rows, domain "
                        f"plane broadcast ({Wp * Hp} B, NCCL) + pool build + k_frac_pool per picture", "scaling": "strong", "pools": out}


def dropin_leg(frames=10):
    """The drop-in on the reference's own clock (part of the cpu_baseline leg: it runs the prebuilt reference binaries of oracle/_ref):
    BASELINE config 1 (QCIF 176x144 IPPP, +-16, 1 reference, QP 28) encoded by the stock JM lencod and by the same objects with the
    motion search served by libb2me.so -- lencod_b2 (SearchMode -1: one GPU call per full_search / sub_pel call) and lencod_b2f
    (SearchMode 0: one GPU call per (macroblock, reference) for the SAD tables).  JM's own `Total ME time` is quoted; the
    bitstreams must be identical.  The per-call drop-in is a CORRECTNESS boundary: every call is a host round trip to the GPU."""
    import re
    import tempfile
    from h264_b200 import synth
    from oracle import jm_run
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not all(os.path.exists(os.path.join(ref, f)) for f in ("lencod", "lencod_b2", "lencod_b2f", "encoder.cfg")):
        return {"unavailable": "oracle/_ref/lencod{,_b2,_b2f} not built"}
    Wq, Hq = 176, 144
    out = {"workload": f"JM 18.5 lencod, synthetic QCIF {Wq}x{Hq}, {frames} frames IPPP, +-16, 1 ref, QP 28 (BASELINE config 1)"}
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(Wq, Hq, frames, seed=20261018))
        for mode, exes in ((-1, ("lencod", "lencod_b2")), (0, ("lencod", "lencod_b2f"))):
            res = {}
            for exe in exes:
                od = os.path.join(d, f"{exe}_{mode}")
                log = jm_run.run_lencod(yuv, Wq, Hq, frames, od, exe=exe, nrefs=1, search_range=16, qp=28, search_mode=mode)
                m = re.search(r"Total ME time for sequence\s*:\s*([0-9.]+) sec", log)
                t = re.search(r"Total encoding time for the seq\.\s*:\s*([0-9.]+) sec", log)
                res[exe] = {"me_time_s": float(m.group(1)) if m else None, "total_s": float(t.group(1)) if t else None,
                            "bitstream": open(os.path.join(od, "out.264"), "rb").read()}
            a, b = res[exes[0]], res[exes[1]]
            out["full_search" if mode == -1 else "fast_full_search"] = {
                "stock_me_time_s": a["me_time_s"], "dropin_me_time_s": b["me_time_s"], "stock_total_s": a["total_s"], "dropin_total_s": b["total_s"],
                "dropin": exes[1], "bitstreams_identical": a["bitstream"] == b["bitstream"]}
    return out


def _cpu_worker(job):
    """One process of the CPU reference arm: its own copy of the reference state, its own MB range."""
    first, cnt, trial = job
    import oracle
    fr, pred, cen = workload()
    cur, refs = fr[NREFS], fr[[NREFS - 1 - r for r in range(NREFS)]]
    if oracle.have_jmref():
        eng = oracle.JMRef(W, H, R, NREFS)
        for r in range(NREFS):
            eng.set_ref(r, refs[r])
        eng.set_cur(cur)
        lam = np.array(LAMBDA, np.int32)
        run = lambda a, b: eng.search_frame(pred, cen, lam, mb_first=a, mb_count=b)
    else:
        eng = oracle.OrcFrame(cur, refs, R)
        run = lambda a, b: eng.search_frame(pred, cen, LAMBDA, mb_first=a, mb_count=b)
    t0 = time.time(); run(first, trial); t_trial = time.time() - t0
    t0 = time.time(); run(first, cnt); dt = time.time() - t0
    return t_trial, dt


def cpu_reference(seconds_budget=20.0, procs=None):
    """The reference's CPU implementation (unmodified JM objects, oracle/_ref/libjmref.so; else the
    restated oracle) on a bounded MB sample of the same workload.  The reference is single-threaded;
    the only parallelism it admits is independent processes, so `procs` processes each search their
    own MB range and the aggregate is reported (cores = procs)."""
    import multiprocessing as mp
    import oracle
    kind = "reference" if oracle.have_jmref() else "port"
    nmb = (W // 16) * (H // 16)
    procs = procs or max(1, min(os.cpu_count() or 1, 32))
    ctx = mp.get_context("spawn")
    # calibrate on one process, then give every process an equal MB range sized to the budget
    t_trial, _ = _cpu_worker((nmb // 2, 2, 8))
    per_mb = t_trial / 8
    cnt = int(max(2, min(nmb // 2, seconds_budget / per_mb)))       # ranges may overlap across processes
    jobs = [((nmb // 2 + i * cnt) % (nmb - cnt), cnt, 1) for i in range(procs)]
    t0 = time.time()
    if procs > 1:
        with ctx.Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    else:
        res = [_cpu_worker(jobs[0])]
    wall = max(r[1] for r in res)            # slowest process = wall time of the parallel search
    v = procs * cnt * NREFS * (2 * R + 1) ** 2 * 256 / wall / 1e6
    return {"value": v, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": f"{procs} processes x {cnt} MBs of {nmb} x {NREFS} refs x 41 partitions, integer +-{R} SAD + sub-pel SATD; "
                      f"slowest process {wall:.1f} s (launch-to-finish {time.time() - t0:.1f} s incl. per-process plane build)",
            "mb_per_s": procs * cnt / wall, "one_core_value": cnt * NREFS * (2 * R + 1) ** 2 * 256 / (sum(r[1] for r in res) / len(res)) / 1e6}


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    per_step = min(30.0, 90.0 / (steps + args.warmup))
    res = None
    t0 = time.time()
    for i in range(args.warmup + steps):
        if i == args.warmup:
            t0 = time.time()
        res = cpu_reference(per_step)
    ms = (time.time() - t0) * 1e3 / steps
    line = {"metric": UNIT, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD,
                       "note": "reference CPU path: unmodified JM objects, one process per host core (the reference is single-threaded), bounded MB sample per step"},
            "cpu_baseline": res, "gpu_launches": 0,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 on their own (NCCL prints its
    # version line there at the first communicator) go to stderr for the duration of the run
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    from h264_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)

    fr, pred_np, cen_np = workload(seed=1 + rank)
    nmb = (W // 16) * (H // 16)
    s = api.Searcher(W, H, NREFS, R, device=local)
    params = api.make_params(LAMBDA)
    n = nmb * NREFS * 41
    stream = torch.cuda.current_stream().cuda_stream

    # ---- resident inputs -------------------------------------------------------------------
    d_frames = torch.from_numpy(fr).to(dev)                     # [NREFS+2, H, W] u8
    d_pred = torch.from_numpy(pred_np).to(dev)
    d_cen = torch.from_numpy(cen_np).to(dev)
    d_mvi = torch.zeros((nmb, NREFS, 41, 2), dtype=torch.int16, device=dev)
    d_mvs = torch.zeros_like(d_mvi)
    d_ci = torch.zeros((nmb, NREFS, 41), dtype=torch.int64, device=dev)
    d_cs = torch.zeros_like(d_ci)
    for r in range(NREFS):
        s.set_ref_dev(r, d_frames[NREFS - 1 - r], stream)
    # L2 flush buffer (larger than the 126 MB L2), written between timed iterations
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_resident(i):
        s.set_cur_dev(d_frames[NREFS], stream)
        # newest reference: rebuild its 16 quarter-pel planes (slot content unchanged so that the
        # predictors stay consistent with the clip's motion)
        s.set_ref_dev(i % NREFS, d_frames[NREFS - 1 - (i % NREFS)], stream)
        s.search_frame_dev(d_pred, d_cen, params, d_mvi, d_ci, d_mvs, d_cs, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step_resident(i)
    barrier()
    l0 = s.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                                     # L2 flush (untimed)
        ev[i][0].record()
        step_resident(i)
        ev[i][1].record()
    barrier()
    launches = s.launch_count() - l0
    ms_dev = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    clk = clocks.finish() if clocks else None
    value = world * pel_sp(nmb) / (ms_step * 1e-3) / 1e6

    # ---- dominant kernel (integer search): live CUDA-event time per launch ------------------
    s.kernel_timing(True)
    for i in range(max(3, min(args.steps, 10))):
        flush.fill_(i)
        step_resident(i)
    torch.cuda.synchronize()
    k_ms, k_n = s.kernel_time_ms(0)
    p_ms, p_n = s.kernel_time_ms(1)
    q_ms, q_n = s.kernel_time_ms(2)
    s.kernel_timing(False)
    k_ms_launch = k_ms / max(k_n, 1)

    # ---- e2e: host buffers through the host-pointer C ABI -----------------------------------
    # Every step: b2me_set_cur + b2me_set_ref + b2me_search_frame_best on pinned HOST buffers: the current picture, the newest
    # reference and one predictor per (MB, ref) go up; what the mode decision consumes -- best reference and cost per (mode, block),
    # the chosen refined vector of every partition -- is on the host on return (synchronous calls).  Measured twice: ONE frame stream per GPU, and E2E_STREAMS independent
    # closed-GOP segment streams per GPU (the north star's own partitioning: segments share no state), each with its own
    # context and host thread, so that one stream's copies run under the other's kernels.  The headline is the latter.
    import ctypes as C
    import threading
    L = s.L
    # three streams per GPU (two leave the result sensitive to how the streams' phases lock) unless the ranks of
    # this box would then outnumber its host cores (one synchronising host thread per stream)
    E2E_STREAMS = int(os.environ.get("B2ME_E2E_STREAMS", "3" if (os.cpu_count() or 1) >= 3 * world + 2 else "2"))

    pred_mb_np = np.ascontiguousarray(pred_np[:, :, 0, :])         # the headline's predictors ARE one per (MB, ref)

    class Stream:
        def __init__(self, ctx):
            self.ctx = ctx
            self.h_cur = torch.from_numpy(fr[NREFS]).pin_memory()
            self.h_ref = [torch.from_numpy(fr[j]).pin_memory() for j in range(NREFS)]
            self.h_pred = torch.from_numpy(pred_mb_np).pin_memory()
            self.h_br = torch.zeros((nmb, 21), dtype=torch.int8).pin_memory()
            self.h_bc = torch.zeros((nmb, 21), dtype=torch.int32).pin_memory()
            self.h_mvs = torch.zeros((nmb, 41, 2), dtype=torch.int16).pin_memory()

        def step(self, i):
            h = self.ctx.h
            r = L.b2me_set_cur(h, C.c_void_p(self.h_cur.data_ptr()), C.c_int(W))
            r |= L.b2me_set_ref(h, C.c_int(i % NREFS), C.c_void_p(self.h_ref[NREFS - 1 - (i % NREFS)].data_ptr()), C.c_int(W))
            # one predictor per (MB, ref) up; best reference / cost per (mode, block) and the chosen vectors down
            r |= L.b2me_search_frame_best(h, C.c_void_p(self.h_pred.data_ptr()), C.byref(params), C.c_int(LAMBDA[2]),
                                          C.c_void_p(self.h_br.data_ptr()), C.c_void_p(self.h_bc.data_ptr()), C.c_void_p(self.h_mvs.data_ptr()))
            if r:
                raise RuntimeError(f"C ABI call failed: {L.b2me_last_error(h)}")

    streams = [Stream(s)]
    for _ in range(E2E_STREAMS - 1):
        s2 = api.Searcher(W, H, NREFS, R, device=local)
        for r_ in range(NREFS):
            s2.set_ref(r_, fr[NREFS - 1 - r_])
        streams.append(Stream(s2))

    def run_streams(active, steps_):
        err = []
        gate = threading.Barrier(len(active) + 1)

        def body(st):
            try:
                torch.cuda.set_device(local)
                gate.wait()
                for i in range(steps_):
                    st.step(i)
            except Exception as e:                               # noqa: BLE001
                err.append(e)
        th = [threading.Thread(target=body, args=(st,)) for st in active]
        for t_ in th:
            t_.start()
        gate.wait()
        t0 = time.perf_counter()
        for t_ in th:
            t_.join()
        torch.cuda.synchronize()
        if err:
            raise err[0]
        return (time.perf_counter() - t0) / (steps_ * len(active))      # wall time per frame

    def timed(active):
        run_streams(active, warmup)
        barrier()
        dt = run_streams(active, args.steps)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    dt1 = timed(streams[:1])
    dtS = timed(streams)
    e2e_value = world * pel_sp(nmb) / dtS / 1e6
    h2d = 2 * W * H + nmb * NREFS * 2 * 2
    d2h = nmb * 21 * (1 + 4) + nmb * 41 * 2 * 2
    checksum = int(streams[0].h_mvs.to(torch.int64).sum().item())           # the result really is on the host
    assert all(int(st.h_mvs.to(torch.int64).sum().item()) == checksum for st in streams)
    for st in streams[1:]:
        st.ctx.close()

    multi = {"bands_4k": bands_leg(local, rank, world), "pool_multi": pool_multi_leg(local, rank, world)}     # every rank takes part
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peaks, peak_src = measured_peaks()
    sad_peak_glane = api.ubench(0, 4000, device=local)           # measured VABSDIFF4.U8.ACC issue rate (G lane-ops/s)
    sad_peak_tpel = sad_peak_glane * 4 / 1e3                      # 4 pixel pairs per lane-op
    achieved_tpel = pel_sp(nmb) / (k_ms_launch * 1e-3) / 1e12
    alg_bytes = W * H + NREFS * (W + 64) * (H + 40) + 2 * n * 4 + n * (4 + 8)   # cur + int planes + pred/centre in, mv+cost out
    roofline = {"bound": "int-alu", "kernel": "k_sad_fs (integer full search, VABSDIFF4.U8.ACC)",
                "achieved": achieved_tpel, "peak": sad_peak_tpel, "unit": "Tpel-sp/s", "frac": achieved_tpel / sad_peak_tpel,
                "peak_source": "measured live: b2me_ubench(kind=0) VABSDIFF4.U8.ACC lane-ops/s x 4 pixels (MEASURED_PEAKS.json has no integer figure)",
                "kernel_ms": k_ms_launch,
                "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (k_ms_launch * 1e-3) / 1e9,
                        "peak_gbs": peaks["hbm_gbs"], "peak_source": peak_src,
                        "frac": alg_bytes / (k_ms_launch * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "traffic": None,
                "share_of_step": {"k_sad_fs": k_ms / max(k_n, 1), "subpel_planes": p_ms / max(p_n, 1), "subpel_refine": q_ms / max(q_n, 1)}}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr_path):
        roofline["traffic"] = json.load(open(tr_path)).get("k_sad_fs")
    secondary = {"fractal_pool": pool_leg(local, peaks), "reconstruction_1080p": recon_leg(local), "fractal_window": fractal_leg(local, not args.no_cpu)}
    secondary.update(multi)
    # the integer search alone under less friendly predictors / content (k_sad_fs is data-dependent)
    roofline["robustness"] = robustness_block(local, sad_peak_tpel)
    cpu = None if args.no_cpu else cpu_reference(15.0)
    if cpu is not None:
        cpu["dropin_jm"] = dropin_leg()
    line = {"metric": UNIT, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "step": "the 16 quarter-pel planes of the newest reference are rebuilt, then integer search, then sub-pel refinement",
                       "predictors": "per-(MB,ref) predictor = clip pan motion, shared by the 41 partitions (search centre = rounded predictor)",
                       "lambda_factor": LAMBDA[0], "sharding": "independent closed-GOP segments (one frame stream per GPU), no collective",
                       "l2": "192 MB L2 flush written between timed iterations"},
            "mb_per_s": world * nmb / (ms_step * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": dtS * 1e3, "streams_per_gpu": E2E_STREAMS,
                    "note": "independent closed-GOP segment streams per GPU, one context + host thread each; a step = one frame of one stream (wall time / frames of all streams); kernels of different streams overlap (one stream's sub-pel stage runs under another's search tail, copies under kernels), so the time per frame can fall below the single-stream resident step that `value` measures",
                    "single_stream": {"value": world * pel_sp(nmb) / dt1 / 1e6, "ms_per_step": dt1 * 1e3},
                    "result_checksum": checksum},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
