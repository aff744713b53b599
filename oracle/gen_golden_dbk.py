#!/usr/bin/env python
"""Golden vectors of the deblocking filter, from the UNMODIFIED reference:
  tests/golden/jm_deblock.npz : every DeblockFrame call of stock `lencod` runs (oracle/_ref/lencod_wrap_dbk = all reference
      objects + the logger oracle/jm_wrap_dbk.c): the unfiltered reconstruction, the b2dbk_mb / b2dbk_blk records filled from
      mb_data[] / mv_info, and the planes the real function left.  Runs: IPPP with two references at QP 36 (intra and inter
      macroblocks, skipped ones, coefficient edges), IBPBP with bi-prediction at QP 30 (two lists, list-swapped reference
      comparison), 8x8 transform on, and filter offsets +2 / -2.
The oracle restatement (orc_deblock_frame) is checked against every captured picture before the file is written.
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
    data = open(path, "rb").read()
    o, pics = 0, []
    while o < len(data):
        tag, W, H, st = struct.unpack_from("<4i", data, o); o += 16
        assert tag == 0x44
        def planes(o):
            y = np.frombuffer(data, np.uint8, W * H, o).reshape(H, W); o += W * H
            u = np.frombuffer(data, np.uint8, W * H // 4, o).reshape(H // 2, W // 2); o += W * H // 4
            v = np.frombuffer(data, np.uint8, W * H // 4, o).reshape(H // 2, W // 2); o += W * H // 4
            return (y, u, v), o
        before, o = planes(o)
        nmb = (W // 16) * (H // 16)
        mbs = np.frombuffer(data, synth.DBK_MB, nmb, o).copy(); o += 12 * nmb
        nb = (W // 4) * (H // 4)
        blks = np.frombuffer(data, synth.DBK_BLK, nb, o).copy(); o += 12 * nb
        after, o = planes(o)
        pics.append((st, before, mbs, blks, after))
    return pics


def run(tag, frames, extra, qp, seed, nrefs=2):
    W, H = 176, 144
    tmp = tempfile.mkdtemp()
    yuv, log = os.path.join(tmp, "in.yuv"), os.path.join(tmp, "wrap.log")
    open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=seed))
    jm_run.run_lencod(yuv, W, H, frames, tmp, exe="lencod_wrap_dbk", search_mode=3, search_range=16, nrefs=nrefs, qp=qp,
                      extra=("EPZSSubPelGrid=0", "DFParametersFlag=1") + tuple(extra), env={"B2_WRAP_LOG": log})
    pics = parse(log)
    out = {}
    for i, (st, before, mbs, blks, after) in enumerate(pics):
        got = oracle.deblock_frame(*before, mbs, blks)
        ok = all((g == a).all() for g, a in zip(got, after))
        changed = sum(int((b != a).sum()) for b, a in zip(before, after))
        print(tag, i, "slice", st, "intra MBs", int(mbs["intra"].sum()), "changed samples", changed, "oracle ==", ok)
        assert ok
        for n, b, a in zip("yuv", before, after):
            out[f"{tag}{i}_{n}0"] = b; out[f"{tag}{i}_{n}1"] = a
        out[f"{tag}{i}_mbs"] = mbs; out[f"{tag}{i}_blks"] = blks
    out[f"{tag}_n"] = len(pics)
    return out


def main():
    d = {}
    d.update(run("p", 4, (), 36, 41))
    d.update(run("b", 5, ("NumberBFrames=1", "BiPredMotionEstimation=1", "HierarchicalCoding=0", "BReferencePictures=0", "QPBSlice=32",
                          "BList1References=1", "DirectModeType=1"), 30, 42))
    d.update(run("t", 3, ("Transform8x8Mode=1", "DFAlphaRefPSlice=2", "DFBetaRefPSlice=-2", "DFAlphaRefISlice=-1", "DFBetaRefISlice=3"), 32, 43))
    np.savez_compressed(os.path.join(GOLD, "jm_deblock.npz"), **d)
    print("written", os.path.getsize(os.path.join(GOLD, "jm_deblock.npz")), "bytes")


if __name__ == "__main__":
    main()
