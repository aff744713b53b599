#!/usr/bin/env python
"""profiles/r2_sass_excerpt.txt: per kernel of h264_b200/libb2me.so, the count of the SASS instructions that prove the hardware
path (tcgen05.mma / TMEM loads, TMA, packed-byte abs-diff, dp4a ...).  usage: python tools/sass_excerpt.py > profiles/<tag>_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "h264_b200", "libb2me.so")], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|VABSDIFF4|VABSDIFF|IDP\.2A|IDP\.4A|SYNCS|UTCBAR|REDUX|VIMNMX3|VOTE|ATOMS|NANOSLEEP)")
print("# SASS excerpt of h264_b200/libb2me.so (cuobjdump -sass, sm_100a): per kernel, the count of the instructions that prove the hardware path --")
print("# UTCIMMA = tcgen05.mma kind::i8, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load (cp.async.bulk.tensor),")
print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier, VABSDIFF4 = packed-byte abs-diff-accumulate, IDP.4A / IDP.2A = dp4a / dp2a, VIMNMX3 = 3-input min/max,")
print("# REDUX = warp reduction, ATOMS = shared-memory atomic.\n")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter(m.group(1) for m in pat.finditer(f))
    if not c:
        continue
    name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    print(f"{name}  ({len(re.findall(r'/[*][0-9a-f]{4,6}[*]/', f))} instructions)")
    print("    " + ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
