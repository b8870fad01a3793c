cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_benchshapes.py tests/test_gpu_factored.py -x -q > gpurun_out/r02_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t2.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/r02_c5_b.json 2> gpurun_out/r02_c5_b.err; echo "bench c5 rc=$?" >> gpurun_out/r02_t2.log
BGGE_MRHS_LAG=3 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/r02_c5_b_lag3.json 2>/dev/null; echo "bench c5 lag3 rc=$?" >> gpurun_out/r02_t2.log
BGGE_MRHS_LAG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/r02_c5_b_lag1.json 2>/dev/null; echo "bench c5 lag1 rc=$?" >> gpurun_out/r02_t2.log
timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/r02_c3_b.json 2> gpurun_out/r02_c3_b.err; echo "bench c3 rc=$?" >> gpurun_out/r02_t2.log
BGGE_MRHS=0 timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/r02_c3_b_old.json 2>/dev/null; echo "bench c3 old rc=$?" >> gpurun_out/r02_t2.log
tail -8 gpurun_out/r02_t2.log
for f in r02_c5_b r02_c5_b_lag3 r02_c5_b_lag1 r02_c3_b r02_c3_b_old; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["roofline"]["kernel"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], [ (k["name"],round(k["ms"],4)) for k in d["sweep"]["kernels"]])
except Exception as e: print(f, "ERR", e)
PY
done
