"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu -i rep --page source --csv | python tools/ncu_top_stalls.py <kernel-substring> [N]"""
import csv, sys
pat = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(sys.stdin))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = b["hdr"]; si = h.index("# Samples"); src = h.index("Source")
    st = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print(b["name"][:90], "instructions", len(b["rows"]), "samples", tot)
    agg = {h[i]: sum(int(r[i] or 0) for r in b["rows"]) for i in st}
    print("  by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si] or 0))[:N]
    for i in sorted(order):
        r = b["rows"][i]
        why = {h[k][6:]: int(r[k]) for k in st if int(r[k] or 0) > 0}
        top = sorted(why.items(), key=lambda kv: -kv[1])[:3]
        print(f"  #{i:5d} {int(r[si]):6d}  {r[src].strip()[:70]:70s} {top}")
