cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r02_v10_tests.log; cat gpurun_out/r02_v10_tests.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 600 python tools/sanitize_cases.py > gpurun_out/r02_v10_checks.log 2>&1; echo "checks rc=$?"; tail -2 gpurun_out/r02_v10_checks.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1500 python bench.py ) > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; echo "bench c5 rc=$?"; tail -4 gpurun_out/bench_r02_c5.err
for w in c1 c2 c3 c4; do
  timeout 900 python bench.py --workload $w > gpurun_out/bench_r02_$w.json 2>gpurun_out/bench_r02_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json
for w in ("c5","c1","c2","c3","c4"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r02_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), d["roofline"]["kernel"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        if w=="c5":
            print("  c4 sub", round(d["c4"]["value"]), "getk", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["getk"].items() if k in("contraction_s","expansion_s","expansion_gbs","write_only_fill_gbs","expansion_frac_of_fill")}, "call_s", d["e2e"]["call_s"])
    except Exception as e:
        print(w, "fail", e)
PY
