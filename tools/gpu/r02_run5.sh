cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r02_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t5.log
tail -40 gpurun_out/r02_t5.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 900 python tools/sanitize_cases.py > gpurun_out/r02_checks_build.log 2>&1; echo "checks build rc=$?"; tail -3 gpurun_out/r02_checks_build.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c5_c.json 2> gpurun_out/r02_c5_c.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_c5_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_c5_c.json").read().strip().splitlines()[-1])
for k in ("value","e2e","e2e_sweep","roofline","c4","cpu_baseline","eigen","clocks"): print(k, json.dumps(d.get(k))[:600])
print(json.dumps(d["getk"])[:500])
PY
