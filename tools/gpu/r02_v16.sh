cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
BGGE_TIMING=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_v16.json 2> gpurun_out/r02_v16.err; echo rc=$?
grep "bgge timing" gpurun_out/r02_v16.err | grep "setDEC of one kernel  \|eig m=10000" | grep -v "   4\.\| 3\.9" | tail -40
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v16.json").read().strip().splitlines()[-1])
print(d["e2e"]["call_s"])
PY
