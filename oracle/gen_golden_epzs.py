#!/usr/bin/env python
"""Golden vectors of the EPZS integer search, from the UNMODIFIED reference:
  tests/golden/jm_epzs.npz : every EPZS_motion_estimation / EPZS_subMB_motion_estimation call (list 0) of stock `lencod` runs
      (oracle/_ref/lencod_wrap_epzs = all reference objects + the boundary logger oracle/jm_wrap_epzs.c) -- the job record the
      drop-in hands to the GPU (built by integration/jm/b2me_jm_epzs_job.h), its predictors, the pattern tables serialised
      from the reference's EPZSStructure objects, the luma planes, and what the real function returned (cost, vector).
Two runs: the default EPZS configuration of encoder.cfg (extended diamond, dual refinement, all predictor kinds) with two
references, and the PMVFAST pattern chain at +-32.  The oracle restatement (orc_epzs_search) is checked against every captured
call here before the file is written; the committed file keeps every KEEP-th call.
Needs /root/reference (build container only)."""
import os, struct, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from oracle import jm_run  # noqa: E402
from h264_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
KEEP = 5


def parse(path):
    data = open(path, "rb").read()
    o, frames, pats, calls = 0, {}, None, []
    js, rs = synth.EPZS_JOB.itemsize, synth.EPZS_PATTERN.itemsize
    while o < len(data):
        tag = struct.unpack_from("<i", data, o)[0]
        if tag == 0x46:
            _, poc, ref, W, H = struct.unpack_from("<5i", data, o); o += 20
            cur = np.frombuffer(data, np.uint8, W * H, o).reshape(H, W); o += W * H
            rf = np.frombuffer(data, np.uint8, W * H, o).reshape(H, W); o += W * H
            frames[(poc, ref)] = (cur, rf)
        elif tag == 0x50:
            n = struct.unpack_from("<i", data, o + 4)[0]; o += 8
            pats = np.frombuffer(data, synth.EPZS_PATTERN, n, o).copy(); o += n * rs
        elif tag == 0x45:
            _, poc, ref, submb, npred = struct.unpack_from("<5i", data, o); o += 20
            job = np.frombuffer(data, synth.EPZS_JOB, 1, o).copy(); o += js
            pv = np.frombuffer(data, np.int16, 2 * npred, o).reshape(npred, 2).copy(); o += 4 * npred
            cost, mx, my = struct.unpack_from("<q2i", data, o); o += 16
            calls.append((poc, ref, submb, job, pv, cost, mx, my))
        else:
            raise ValueError(f"bad tag {tag:x} at {o}")
    return frames, pats, calls


def run(tag, frames_n, nrefs, sr, extra, seed):
    W, H = 176, 144
    tmp = tempfile.mkdtemp()
    yuv, log = os.path.join(tmp, "in.yuv"), os.path.join(tmp, "wrap.log")
    open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames_n, seed=seed))
    jm_run.run_lencod(yuv, W, H, frames_n, tmp, exe="lencod_wrap_epzs", search_mode=3, search_range=sr, nrefs=nrefs, qp=28,
                      extra=("EPZSSubPelGrid=0",) + tuple(extra), env={"B2_WRAP_LOG": log})
    frames, pats, calls = parse(log)
    pocs = sorted({k[0] for k in frames})
    out = dict(W=W, H=H, R=sr, pocs=np.array(pocs), patterns=pats)
    jobs, preds, res, jpoc = [], [], [], []
    bad = 0
    for poc in pocs:
        refs = [frames[(poc, r)][1] for r in range(nrefs) if (poc, r) in frames]
        fr = oracle.OrcFrame(frames[(poc, 0)][0], refs, sr)
        mine = [c for c in calls if c[0] == poc]
        for (_, ref, submb, job, pv, cost, mx, my) in mine:
            j = job.copy(); j["pred_first"] = 0
            r = fr.epzs_search(j, pv, pats)[0]
            if (int(r["cost"]), int(r["mv"][0]), int(r["mv"][1])) != (cost, mx, my):
                bad += 1
                if bad < 5:
                    print("MISMATCH", tag, poc, ref, submb, j, (cost, mx, my), r)
        for i, (_, ref, submb, job, pv, cost, mx, my) in enumerate(mine):
            if i % KEEP:
                continue
            j = job.copy(); j["pred_first"] = sum(len(p) for p in preds)
            jobs.append(j); preds.append(pv); res.append((cost, mx, my)); jpoc.append(poc)
        out[f"cur_{poc}"] = frames[(poc, 0)][0]
        for r, rf in enumerate(refs):
            out[f"ref_{poc}_{r}"] = rf
    print(tag, len(calls), "calls,", bad, "mismatches against the oracle;", len(jobs), "kept")
    assert bad == 0
    out.update(jobs=np.concatenate(jobs), preds=np.concatenate(preds), results=np.array(res, np.int64), job_poc=np.array(jpoc), nrefs=nrefs)
    return {f"{tag}_{k}": v for k, v in out.items()}


def main():
    d = {}
    d.update(run("a", 4, 2, 16, (), 31))
    d.update(run("b", 3, 1, 32, ("EPZSPattern=5", "EPZSDualRefinement=6"), 32))
    np.savez_compressed(os.path.join(GOLD, "jm_epzs.npz"), **d)
    print("written", os.path.getsize(os.path.join(GOLD, "jm_epzs.npz")), "bytes")


if __name__ == "__main__":
    main()
