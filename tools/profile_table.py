"""profiles/<tag>_kernels.md: one line per captured kernel from the ncu summaries of a round (tools/ncu_summary.py output):
duration, DRAM bytes and GB/s against the measured HBM peak, ALU / FMA / tensor pipe and issue utilisation.
usage: python tools/profile_table.py r1p"""
import glob, json, os, re, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6554.2
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tmul = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
rows = []
for f in sorted(glob.glob(os.path.join(root, "profiles", f"{tag}_*.ncu.txt"))):
    for blk in open(f).read().split("\nkernel: ")[1:]:
        name = blk.split("(")[0].replace("void ", "").strip()
        grid = re.search(r"grid \(([^)]*)\) block \(([^)]*)\)", blk)
        def val(k, units=None):
            m = re.search(r"^\s+" + re.escape(k) + r"\s+([\d.eE+-]+)\s*(\S*)", blk, re.M)
            if not m:
                return None
            v = float(m.group(1))
            return v * units.get(m.group(2), 1) if units else v
        us = val("gpu__time_duration.sum", tmul)
        rd, wr = val("dram__bytes_read.sum", mul) or 0, val("dram__bytes_write.sum", mul) or 0
        rows.append((os.path.basename(f), name, grid.group(1) + " x " + grid.group(2) if grid else "", us, rd + wr, (rd + wr) / us / 1e3 if us else 0,
                     val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                     val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or val("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
                     val("smsp__issue_active.avg.pct_of_peak_sustained_active"), val("launch__registers_per_thread")))
out = [f"# Kernels captured in round {tag} (`ncu --set full --clock-control none`, one launch each; cold-cache, serialised)", "",
       f"HBM peak used for the fraction: {peak} GB/s (MEASURED_PEAKS.json).  Times under ncu are longer than in `bench.py` (no overlap, cold caches):",
       "the bench line's CUDA-event times are the ones quoted in DESIGN.md.", "",
       "| summary file | kernel | grid x block | us | DRAM MB | GB/s | % HBM peak | ALU % | FMA % | tensor % | issue % | regs |", "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
f2 = lambda v: "" if v is None else f"{v:.1f}"
for r in rows:
    out.append(f"| {r[0]} | `{r[1]}` | {r[2]} | {r[3]:.1f} | {r[4] / 1e6:.2f} | {r[5]:.0f} | {100 * r[5] / peak:.1f} | {f2(r[6])} | {f2(r[7])} | {f2(r[8])} | {f2(r[9])} | {f2(r[10])} |")
open(os.path.join(root, "profiles", f"{tag}_kernels.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
