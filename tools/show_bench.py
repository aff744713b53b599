"""prints the interesting parts of a bench.py JSON line"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"] / 1e6, 2), "T; ms", round(d["ms_per_step"], 3), "; e2e", round(d["e2e"]["value"] / 1e6, 2), "T", round(d["e2e"]["ms_per_step"], 3), "ms; frac", round(d["roofline"]["frac"], 4))
print("robustness", {k: (round(v["kernel_ms"], 3), round(v["frac"], 3)) for k, v in d["roofline"].get("robustness", {}).items()})
for k, v in d.get("secondary", {}).items():
    print(k, json.dumps(v)[:700])
