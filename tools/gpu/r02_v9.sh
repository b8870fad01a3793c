cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --workload c3 --no-c4 --no-getk-host --no-cpu-baseline --no-dense-ref --steps 3 --warmup 3 > gpurun_out/r02_v9_$name.json 2>gpurun_out/r02_v9_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_v9_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name fail", e)
PY
}
run cs2 BGGE_MRHS_CS=2
run cs4 BGGE_MRHS_CS=4
run fcs2 BGGE_FUSED_CS=2
run cs2fcs2 BGGE_MRHS_CS=2 BGGE_FUSED_CS=2
run cs4fcs2 BGGE_MRHS_CS=4 BGGE_FUSED_CS=2
run cs2fcs4 BGGE_MRHS_CS=2 BGGE_FUSED_CS=4
