"""Reads one bench.py JSON line on stdin and prints a compact summary (label = argv[1])."""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
txt = sys.stdin.read().strip().splitlines()
d = json.loads(txt[-1])
s = d.get("sweep", {})
print(label, "value", round(d["value"], 1), d["unit"], "| e2e", round(d["e2e"]["value"], 1), "| frac", round(s.get("frac_of_hbm_peak", 0), 3),
      "|", s.get("accounting", "")[:6], "| clocks", d.get("clocks"))
for k in s.get("kernels", []):
    print("   ", k["name"], "ms", round(k["ms"], 4), "GB/s", round(k["gbs"]))
print("    paths", s.get("kernel_paths"))
g = d.get("getk")
if g:
    print("    getk contraction", [round(x, 4) for x in g.get("contraction_s_all", [])], "TF", round(g["contraction_tflops"], 2), "frac", round(g["frac_of_dgemm_burst"], 3),
          "quantile+exp", round(g["quantile_exp_s"], 4), "expand", round(g["expansion_s"], 4), "setDEC", round(d["eigen"]["setdec_s"], 2), "s for", d["eigen"]["problems"], "blocks")
if d.get("cpu_baseline"):
    print("    cpu", d["cpu_baseline"])
