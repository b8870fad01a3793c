cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
BGGE_TIMING=1 timeout 900 python bench.py --no-c4 --no-getk-host --no-cpu-baseline --no-dense-ref --steps 3 --warmup 3 > gpurun_out/r02_v15.json 2> gpurun_out/r02_v15.err; echo rc=$?
grep "bgge timing" gpurun_out/r02_v15.err | tail -80
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v15.json").read().strip().splitlines()[-1])
print(d["e2e"]["call_s"])
PY
