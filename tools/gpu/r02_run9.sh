cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_benchshapes.py -q > gpurun_out/r02_t9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_t9.log
for L in 1 2; do
BGGE_MRHS_TMEM=1 BGGE_MRHS_LAG=$L timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref --no-c4 > gpurun_out/r02_c5_tmem_lag$L.json 2>/dev/null; echo "bench tmem lag=$L rc=$?"
python - $L <<'PY'
import json,sys
T=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02_c5_tmem_lag{T}.json").read().strip().splitlines()[-1])
print("tmem lag",T,"value",d["value"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"])
PY
done
for CH in 40 80 320; do
CMD="python bench.py --workload c4 --chains $CH --steps 2 --warmup 3 --iters 50 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/r02_c4_$CH.json 2> gpurun_out/r02_c4_$CH.err; echo "c4-$CH rc=$?"
python - $CH <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r02_c4_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("c4 chains", sys.argv[1], "value", d["value"], "ms/it", d["ms_per_step"]/50, "TF", d["roofline"]["achieved"], d["run"]["padded_chains_per_gpu"])
PY
done
CMD="python bench.py --workload c4 --chains 40 --steps 2 --warmup 3 --iters 50 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 40 --csv --log-file gpurun_out/r02_c4_40_launches_b.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_c4_40_launches_b.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    agg.setdefault(r[ki].split('(')[0][-40:],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(f"{k:45s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.1f} us")
PY
