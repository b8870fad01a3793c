cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r02_t11.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_t11.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 900 python tools/sanitize_cases.py > gpurun_out/r02_checks_build.log 2>&1; echo "checks build rc=$?"; tail -2 gpurun_out/r02_checks_build.log
for CH in 40 80 160 320; do
timeout 600 python bench.py --workload c4 --chains $CH --steps 2 --warmup 3 --iters 50 --no-cpu-baseline > gpurun_out/r02_c4_$CH.json 2> gpurun_out/r02_c4_$CH.err; echo "c4-$CH rc=$?"
python - $CH <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r02_c4_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("c4 chains", sys.argv[1], "value", round(d["value"]), "us/it", round(1000*d["ms_per_step"]/50,1), "TF", round(d["roofline"]["achieved"],2), d["run"]["padded_chains_per_gpu"])
PY
done
for W in c3 c2 c1; do
timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-dense-ref > gpurun_out/r02_${W}_d.json 2>/dev/null; echo "$W rc=$?"
python - $W <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r02_{sys.argv[1]}_d.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
done
BGGE_MRHS_TMEM=0 timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-dense-ref --no-cpu-baseline > gpurun_out/r02_c3_d_tm0.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_c3_d_tm0.json").read().strip().splitlines()[-1]); print("c3 tmem=0 value", round(d["value"],1))
PY
CMD="python bench.py --workload c4 --chains 40 --steps 1 --warmup 3 --iters 20 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 20 --csv --log-file gpurun_out/r02_c4_40_launches_c.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_c4_40_launches_c.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    agg.setdefault(r[ki].split('(')[0][-40:],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(f"{k:45s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.1f} us")
PY
