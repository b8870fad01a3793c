cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 )
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 900 python tools/sanitize_cases.py 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1500 python bench.py ) > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; echo "bench c5 rc=$?"; tail -4 gpurun_out/bench_r02_c5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02_c5.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), "clk", d["clocks"], "c4", round(d["c4"]["value"]))
print(d["getk"]["host_buffers"])
PY
