"""TEST INFRASTRUCTURE ONLY: run the stock JM 18.5 `lencod` built by oracle/Makefile.jm.

Every config-relevant key is overridden on the command line because the author's
JM/bin/encoder.cfg differs from the BASELINE configs (SURVEY Q-J8)."""
import os
import subprocess

REF_JM = "/root/reference/4.对比程序/jm18.5/JM"
HERE = os.path.dirname(os.path.abspath(__file__))


def run_lencod(yuv, W, H, frames, outdir, exe="lencod", search_mode=-1, search_range=16, nrefs=1, qp=28,
               subpel=True, restrict=2, rdo=1, extra=(), env=None, cfg=None):
    os.makedirs(outdir, exist_ok=True)
    cfg = cfg or os.path.join(HERE, "_ref", "encoder.cfg")   # copied from JM/bin/encoder.cfg by Makefile.jm
    keys = {
        "InputFile": yuv, "SourceWidth": W, "SourceHeight": H, "OutputWidth": W, "OutputHeight": H,
        "FramesToBeEncoded": frames, "SearchMode": search_mode, "SearchRange": search_range,
        "NumberReferenceFrames": nrefs, "DisableSubpelME": 0 if subpel else 1, "OnTheFlyFractMCP": 0,
        "QPISlice": qp, "QPPSlice": qp, "RestrictSearchRange": restrict, "RDOptimization": rdo,
        "MEDistortionFPel": 0, "MEDistortionHPel": 2, "MEDistortionQPel": 2, "NumberBFrames": 0,
        "IntraPeriod": 0, "LevelIDC": 51, "ProfileIDC": 100, "Transform8x8Mode": 0, "SymbolMode": 0,
        "ReconFile": os.path.join(outdir, "rec.yuv"), "OutputFile": os.path.join(outdir, "out.264"),
        "StatsFile": os.path.join(outdir, "stats.dat"), "TraceFile": os.path.join(outdir, "trace.txt"),
        "RateControlEnable": 0, "UseWeightedReferenceME": 0, "WeightedPrediction": 0, "ChromaMEEnable": 0,
        "PList0References": 0, "BiPredMotionEstimation": 0, "SearchMode": search_mode, "FrameSkip": 0,
        "NumberOfViews": 1,
    }
    cmd = [os.path.join(HERE, "_ref", exe), "-d", cfg]
    for k, v in keys.items():
        cmd += ["-p", f"{k}={v}"]
    for kv in extra:
        cmd += ["-p", kv]
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(cmd, cwd=outdir, env=e, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"lencod failed: {r.stdout[-2000:]} {r.stderr[-2000:]}")
    return r.stdout + r.stderr
