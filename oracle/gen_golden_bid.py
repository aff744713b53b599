#!/usr/bin/env python
"""Golden vectors of BIDPartitionCost (the bi-predictive direction of list_prediction_cost), from the UNMODIFIED reference:
  tests/golden/jm_bid.npz : calls of stock `lencod` runs (oracle/_ref/lencod_wrap_bid = all reference objects + the logger
      oracle/jm_wrap_bid.c): per coded B picture the current luma and the reference lumas it read, the b2me_bid_job records
      built by integration/jm/b2me_jm_bid_job.h and the costs the real function returned.  Runs: SATD metric (default) with two
      references per list; the 8x8 transform on; implicit weighted bi-prediction (WeightedBiprediction=2) on a fading clip;
      bi-predictive motion estimation on (the twin BPredPartitionCost on the bipred_mv vectors, same record).
The oracle restatement (orc_bid_partition_cost) is checked against every kept call before the file is written.
Needs /root/reference (build container only)."""
import os, struct, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from oracle import jm_run  # noqa: E402
from h264_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def parse(path):
    """-> list of pictures: dict(cur, refs [list of planes], jobs, cost, metric, t8, wp, denom)"""
    data = open(path, "rb").read()
    o, pics = 0, []
    while o < len(data):
        tag, = struct.unpack_from("<i", data, o)
        if tag == 0x43:
            _, poc, W, H = struct.unpack_from("<4i", data, o); o += 16
            pics.append(dict(poc=poc, cur=np.frombuffer(data, np.uint8, W * H, o).reshape(H, W).copy(), refs=[], jobs=[], cost=[], par=[]))
            o += W * H
        elif tag == 0x52:
            _, slot, W, H = struct.unpack_from("<4i", data, o); o += 16
            assert slot == len(pics[-1]["refs"])
            pics[-1]["refs"].append(np.frombuffer(data, np.uint8, W * H, o).reshape(H, W).copy()); o += W * H
        elif tag == 0x42:
            _, metric, t8, wp, denom = struct.unpack_from("<5i", data, o); o += 20
            pics[-1]["jobs"].append(np.frombuffer(data, synth.BID_JOB, 1, o)[0].copy()); o += synth.BID_JOB.itemsize
            pics[-1]["cost"].append(struct.unpack_from("<q", data, o)[0]); o += 8
            pics[-1]["par"].append((metric, t8, wp & 1, denom, wp >> 1))       # last: 1 = the call was BPredPartitionCost
        else:
            raise ValueError(hex(tag))
    return pics


def check(pic):
    of = oracle.OrcFrame(pic["cur"], np.stack(pic["refs"]), 8)
    jobs, cost, par = np.array(pic["jobs"], synth.BID_JOB), np.array(pic["cost"], np.int64), np.array(pic["par"], np.int32)
    got = np.zeros_like(cost)
    for p in np.unique(par, axis=0):
        m = (par == p).all(axis=1)
        got[m] = of.bid_partition_cost(jobs[m], int(p[0]), bool(p[1]), bool(p[2]), int(p[3]))
    return jobs, cost, par, got


def run(tag, frames, extra, seed, stride, fade=False):
    W, H = 176, 144
    tmp = tempfile.mkdtemp()
    yuv, log = os.path.join(tmp, "in.yuv"), os.path.join(tmp, "wrap.log")
    seq = np.frombuffer(synth.yuv420_sequence(W, H, frames, seed=seed), np.uint8).reshape(frames, -1).copy()
    if fade:                                   # brightness ramp: implicit / explicit weights differ from (32, 32)
        for i in range(frames):
            seq[i, :W * H] = np.clip(seq[i, :W * H].astype(np.int32) * (100 - 9 * i) // 100 + 3 * i, 0, 255).astype(np.uint8)
    open(yuv, "wb").write(seq.tobytes())
    base = ("NumberBFrames=1", "HierarchicalCoding=0", "BReferencePictures=0", "QPBSlice=32", "DirectModeType=1")
    if not any(e.startswith("BiPredMotionEstimation") for e in extra):
        base += ("BiPredMotionEstimation=0",)
    jm_run.run_lencod(yuv, W, H, frames, tmp, exe="lencod_wrap_bid", search_mode=-1, search_range=8, nrefs=2, qp=30,
                      extra=base + tuple(extra), env={"B2_WRAP_LOG": log, "B2_WRAP_STRIDE": str(stride)})
    out, n = {}, 0
    for i, pic in enumerate(parse(log)):
        if not pic["jobs"]:
            continue
        jobs, cost, par, got = check(pic)
        bad = int((got != cost).sum())
        print(tag, "poc", pic["poc"], "calls", len(cost), "refs", len(pic["refs"]), "blocktypes", sorted(set(jobs["blocktype"].tolist())),
              "params", np.unique(par, axis=0).tolist(), "BPredPartitionCost calls", int(par[:, 4].sum()), "oracle mismatches", bad)
        assert bad == 0, (jobs[got != cost][:3], cost[got != cost][:3], got[got != cost][:3])
        out[f"{tag}{n}_cur"] = pic["cur"]; out[f"{tag}{n}_refs"] = np.stack(pic["refs"])
        out[f"{tag}{n}_jobs"] = jobs; out[f"{tag}{n}_cost"] = cost; out[f"{tag}{n}_par"] = par
        n += 1
    out[f"{tag}_n"] = n
    return out


def main():
    d = {}
    d.update(run("s", 3, ("BList0References=2", "BList1References=1"), 51, 7))
    d.update(run("t", 3, ("Transform8x8Mode=1", "ModeDecisionMetric=0"), 52, 7))
    d.update(run("w", 5, ("WeightedBiprediction=2",), 53, 9, fade=True))
    # bi-predictive motion estimation on: the mode decision also calls the twin BPredPartitionCost (mv_search.c:589-700) on the
    # vectors of the bi-predictive search (bipred_mv[list]); same record, same device function
    d.update(run("p", 3, ("BiPredMotionEstimation=1", "BiPredMERefinements=1", "BiPredMESearchRange=8", "BiPredMESubPel=2"), 54, 5))
    np.savez_compressed(os.path.join(GOLD, "jm_bid.npz"), **d)
    print("written", os.path.getsize(os.path.join(GOLD, "jm_bid.npz")), "bytes")


if __name__ == "__main__":
    main()
