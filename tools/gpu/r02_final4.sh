cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; echo "bench c5 rc=$?"; tail -4 gpurun_out/bench_r02_c5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02_c5.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), "c4", round(d["c4"]["value"]))
print("getk host", d["getk"]["host_buffers"]); print("eigen", d["eigen"])
PY
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_r02_c4.json 2>/dev/null; echo "c4 rc=$?"
