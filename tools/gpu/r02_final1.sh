cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tfinal.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_tfinal.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 1200 python bench.py > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; echo "bench c5 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_c5_reference.json 2>/dev/null; echo "ref rc=$?"
for W in c1 c2 c3 c4; do timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_r02_$W.json 2>/dev/null; echo "$W rc=$?"; done
python - <<'PY'
import json
for w in ("c5","c1","c2","c3","c4","c5_reference"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r02_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],1))
    except Exception as e: print(w, "ERR", e)
PY
CMD="python bench.py --steps 2 --warmup 3 --iters 5 --prof-iters 2 --no-cpu-baseline --no-c4 --no-dense-ref"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c5.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 12 -c 4 -o gpurun_out/r02_sweeps $CMD > gpurun_out/r02_ncu_sweeps.log 2>&1; echo "ncu full rc=$?"
