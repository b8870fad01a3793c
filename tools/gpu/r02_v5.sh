cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r02_v5_tests.log; cat gpurun_out/r02_v5_tests.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 600 python tools/sanitize_cases.py > gpurun_out/r02_v5_checks.log 2>&1; echo "checks rc=$?"; tail -3 gpurun_out/r02_v5_checks.log
for w in c5 c3 c2; do
  for nd in 0 1; do
    if [ $nd = 1 ]; then export BGGE_DIRECT_RESIDUAL=1; else unset BGGE_DIRECT_RESIDUAL; fi
    timeout 600 python bench.py --workload $w --no-c4 --no-getk-host --no-cpu-baseline > gpurun_out/r02_v5_${w}_nd${nd}.json 2>gpurun_out/r02_v5_${w}_nd${nd}.err; echo "$w nd=$nd rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_v5_${w}_nd${nd}.json").read().strip().splitlines()[-1])
    print("$w nd=$nd value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("fail", e)
PY
  done
done
