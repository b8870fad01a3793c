cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4 )
timeout 900 python tools/fuzz_parity.py 150 33 batch > gpurun_out/r02_fuzz_batch_33.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/r02_fuzz_batch_33.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 600 python tools/sanitize_cases.py 2>&1 | tail -2
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline ${CH} > gpurun_out/r02_v19_$name.json 2>gpurun_out/r02_v19_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_v19_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "us/iter", round(1e3*d["ms_per_step"]/d["run"]["iters_per_step"],1), "launches", d["gpu_launches"])
except Exception as e:
    print("$name fail", e); print(open("gpurun_out/r02_v19_$name.err").read()[-600:])
PY
}
CH="--chains 40"; run c40 A=1
CH="--chains 80"; run c80 A=1
CH=""; run c320 A=1
