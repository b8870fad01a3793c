cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for T in 0 1 0 1; do
BGGE_MRHS_TMEM=$T timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref --no-c4 > gpurun_out/r02_ab_tmem$T.json 2>/dev/null
python - $T <<'PY'
import json,sys
T=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02_ab_tmem{T}.json").read().strip().splitlines()[-1])
print("tmem",T,"value",round(d["value"],1),"GE ms",round(d["roofline"]["ms_per_launch"],4),round(d["roofline"]["frac"],4),[(k["name"][:12],round(k["ms"],4)) for k in d["sweep"]["kernels"]])
PY
done
CMD="python bench.py --workload c4 --chains 40 --steps 1 --warmup 3 --iters 20 --no-cpu-baseline"
$CMD > gpurun_out/r02_c4_40_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 40 -c 2 -o gpurun_out/r02_dgemm40 $CMD > gpurun_out/r02_ncu_dgemm40.log 2>&1
echo "ncu rc=$?"
