cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_abi.py -q > gpurun_out/r02_t12.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_t12.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 900 python tools/sanitize_cases.py > gpurun_out/r02_checks_build.log 2>&1; echo "checks build rc=$?"; tail -2 gpurun_out/r02_checks_build.log
for CH in 40 80 320; do
timeout 600 python bench.py --workload c4 --chains $CH --steps 2 --warmup 3 --iters 50 --no-cpu-baseline > gpurun_out/r02_c4_$CH.json 2> gpurun_out/r02_c4_$CH.err; echo "c4-$CH rc=$?"
python - $CH <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r02_c4_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("c4 chains", sys.argv[1], "value", round(d["value"]), "us/it", round(1000*d["ms_per_step"]/50,1), "TF", round(d["roofline"]["achieved"],2), d["run"]["padded_chains_per_gpu"])
PY
done
true
python - <<'PY'
import json
pass
PY
CMD="python bench.py --workload c4 --chains 40 --steps 1 --warmup 3 --iters 20 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 16 --csv --log-file gpurun_out/r02_c4_40_launches_d.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_c4_40_launches_d.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
for r in rows[hi+1:hi+9]:
    if len(r)<=vi: continue
    print(f"{r[ki].split('(')[0][-36:]:38s} grid={r[gi]:16s} {float(r[vi].replace(',',''))/1000:8.1f} us")
PY
