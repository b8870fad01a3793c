cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --workload c4 --chains 40 --steps 1 --warmup 3 --iters 20 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/launches_c4_40_plain.json 2>gpurun_out/launches_c4_40_plain.err || { echo plain failed; tail -5 gpurun_out/launches_c4_40_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"b_|dgemm" -c 400 --csv --log-file gpurun_out/r02_launches_c4_40.csv $CMD > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/r02_launches_c4_40.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; ki=H.index("Kernel Name"); vi=H.index("Metric Value")
seq=[(r[ki].split("(")[0].split("::")[-1][:40], float(r[vi].replace(",",""))/1e3) for r in rows[hdr+2:] if len(r)>vi]
# last complete iteration: find last 'b_chain_mu_kernel'
names=[s[0] for s in seq]
ends=[i for i,n in enumerate(names) if n.startswith("b_chain_mu")]
a,b=ends[-2]+1,ends[-1]+1
tot=0
for n,t in seq[a:b]:
    print(f"{n:42s} {t:8.2f} us"); tot+=t
print("sum", round(tot,1), "us")
PY
