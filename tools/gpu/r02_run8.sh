cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_benchshapes.py -x -q > gpurun_out/r02_t8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_t8.log
for T in 0 1; do
BGGE_MRHS_TMEM=$T BGGE_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref --no-c4 > gpurun_out/r02_c5_tmem$T.json 2> gpurun_out/r02_c5_tmem$T.err; echo "bench tmem=$T rc=$?"; grep "engine=1" gpurun_out/r02_c5_tmem$T.err | head -1
python - $T <<'PY'
import json,sys
T=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02_c5_tmem{T}.json").read().strip().splitlines()[-1])
print("tmem",T,"value",d["value"],"roofline",d["roofline"]["kernel"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],[(k["name"],round(k["ms"],4)) for k in d["sweep"]["kernels"]])
PY
done
CMD="python bench.py --workload c4 --chains 40 --steps 2 --warmup 3 --iters 50 --no-cpu-baseline"
$CMD > gpurun_out/r02_c4_40.json 2> gpurun_out/r02_c4_40.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 60 --csv --log-file gpurun_out/r02_c4_40_launches.csv $CMD > /dev/null 2>&1
echo "c4-40 rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_c4_40.json").read().strip().splitlines()[-1])
print("c4 40 chains 1 gpu:", d["value"], d["ms_per_step"], d["roofline"]["achieved"])
PY
tail -40 gpurun_out/r02_c4_40_launches.csv | cut -d, -f5,9- | cut -c1-160
