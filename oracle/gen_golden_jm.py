"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/jm_*.npz from the UNMODIFIED reference:

  jm_wrap_foreman.npz : inputs/outputs of every full_search_motion_estimation and
      sub_pel_motion_estimation call of a stock `lencod` run (oracle/_ref/lencod_wrap, the
      boundary logger oracle/jm_wrap.c) on the only clip the reference ships,
      JM/bin/foreman_part_qcif.yuv, 3 frames, FS +-16, 1 ref, QP28, SAD int-pel + SATD sub-pel
      (= BASELINE config 1 on real content), together with the luma planes those calls read.
  jm_harness_qcif.npz : outputs of the reference objects driven through oracle/jm_harness.c on
      seeded synthetic QCIF frames with per-partition predictors (borders, 2 refs, restricted
      ranges), in the product's batch layout.

Run in the build container (needs /root/reference):  python oracle/gen_golden_jm.py
"""
import os
import struct
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from oracle import jm_run  # noqa: E402
from h264_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def parse_wrap_log(path):
    data = open(path, "rb").read()
    o, frames, ints, subs = 0, {}, [], []
    while o < len(data):
        tag = struct.unpack_from("<i", data, o)[0]
        if tag == 0x46:
            _, poc, ref, W, H = struct.unpack_from("<5i", data, o); o += 20
            cur = np.frombuffer(data, np.uint8, W * H, o).reshape(H, W); o += W * H
            rf = np.frombuffer(data, np.uint8, W * H, o).reshape(H, W); o += W * H
            frames[(poc, ref)] = (cur, rf)
        elif tag == 0x49:
            v = struct.unpack_from("<12iq2iq", data, o); o += 48 + 8 + 8 + 8
            ints.append(v[1:])
        elif tag == 0x53:
            v = struct.unpack_from("<12iq2iq", data, o); o += 48 + 8 + 8 + 8
            subs.append(v[1:])
        else:
            raise ValueError(f"bad tag {tag:x} at {o}")
    return frames, np.array(ints, np.int64), np.array(subs, np.int64)


def gen_wrap():
    tmp = tempfile.mkdtemp()
    log = os.path.join(tmp, "wrap.log")
    jm_run.run_lencod(jm_run.REF_JM + "/bin/foreman_part_qcif.yuv", 176, 144, 3, tmp, exe="lencod_wrap",
                      env={"B2_WRAP_LOG": log})
    frames, ints, subs = parse_wrap_log(log)
    pocs = sorted({k[0] for k in frames})
    cur = np.stack([frames[(p, 0)][0] for p in pocs])
    ref = np.stack([frames[(p, 0)][1] for p in pocs])
    # columns: poc pos_x pos_y blocktype ref pred_x pred_y cen_x cen_y sr lambda min_in out_x out_y cost
    np.savez_compressed(os.path.join(GOLD, "jm_wrap_foreman.npz"), pocs=np.array(pocs), cur=cur, ref=ref,
                        int_calls=ints.astype(np.int64), sub_calls=subs.astype(np.int64))
    print("wrap:", len(ints), "int calls", len(subs), "sub calls", "pocs", pocs)


def gen_harness():
    W, H, R, NR = 176, 144, 16, 2
    fr = synth.luma_sequence(W, H, 3, seed=7)
    cur, refs = fr[2], fr[[1, 0]]
    out = {}
    for name, restrict, spread, rmax, lam in (("a", 2, 3, 6, (187, 187, 187)), ("b", 0, 9, 40, (60, 47, 47))):
        jm = oracle.JMRef(W, H, R, NR, restrict_mode=restrict)
        for r in range(NR):
            jm.set_ref(r, refs[r])
        jm.set_cur(cur)
        pred, cen = synth.predictors(W, H, NR, seed=11, spread=spread, rmax=rmax)
        mi, ci, ms, cs = jm.search_frame(pred, cen, np.array(lam, np.int32))
        out.update({f"{name}_mv_int": mi, f"{name}_cost_int": ci, f"{name}_mv_sub": ms, f"{name}_cost_sub": cs,
                    f"{name}_cfg": np.array([restrict, spread, rmax, *lam])})
    np.savez_compressed(os.path.join(GOLD, "jm_harness_qcif.npz"), **out)
    print("harness golden written")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_wrap()
    gen_harness()
