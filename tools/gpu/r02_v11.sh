cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -4 )
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 600 python tools/sanitize_cases.py 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline ${CH} > gpurun_out/r02_v11_$name.json 2>gpurun_out/r02_v11_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_v11_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "iters", d["run"]["iters_per_step"] if "run" in d else "", "us/iter", round(1e3*d["ms_per_step"]/d["run"]["iters_per_step"],1) if "run" in d else "")
except Exception as e:
    print("$name fail", e); print(open("gpurun_out/r02_v11_$name.err").read()[-600:])
PY
}
CH=""; run c320_pdl A=1; run c320_nopdl BGGE_BATCH_PDL=0
CH="--chains 40"; run c40_pdl A=1; run c40_nopdl BGGE_BATCH_PDL=0
CH="--chains 80"; run c80_pdl A=1; run c80_nopdl BGGE_BATCH_PDL=0
