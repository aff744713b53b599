"""Per-source-line instruction and stall-sample totals of an .ncu-rep (source page, cuda+sass view).
usage: python tools/ncu_lines.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = "?"; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or r[0] == "": continue
    if len(r) > len(hdr): r = [r[0], ",".join(r[1:len(r) - len(hdr) + 2])] + r[len(r) - len(hdr) + 2:]
    d = {k: (v if v not in ("-", "") else "0") for k, v in zip(hdr[4:], r[4:])}
    out.append((fname, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"]), int(d.get("stall_long_sb", 0)), int(d.get("stall_barrier", 0)), int(d.get("stall_wait", 0)), int(d.get("stall_math", 0)), int(d.get("stall_short_sb", 0))))
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total inst {ti}  samples {ts}")
print(f"{'file:line':22s} {'inst%':>6s} {'smp%':>6s} {'long':>5s} {'bar':>5s} {'wait':>5s} {'math':>5s} {'short':>5s}  source")
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print(f"{o[0][:14]+':'+str(o[1]):22s} {100*o[3]/ti:6.2f} {100*o[4]/ts:6.2f} {o[5]:5d} {o[6]:5d} {o[7]:5d} {o[8]:5d} {o[9]:5d}  {o[2]}")
