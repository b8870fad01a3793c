cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
run() { N=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run 8 --steps 3 --warmup 3 > gpurun_out/bench_r02_c5_8gpu.json 2> gpurun_out/bench_r02_c5_8gpu.err; echo "8gpu rc=$?"
run 4 --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/bench_r02_c5_4gpu.json 2>/dev/null; echo "4gpu rc=$?"
run 4 --steps 3 --warmup 3 --shard-chain --no-cpu-baseline --no-dense-ref > gpurun_out/bench_r02_c5_shard4.json 2> gpurun_out/bench_r02_c5_shard4.err; echo "shard4 rc=$?"; tail -3 gpurun_out/bench_r02_c5_shard4.err | cut -c1-300
run 2 --steps 3 --warmup 3 --shard-chain --no-cpu-baseline --no-dense-ref > gpurun_out/bench_r02_c5_shard2.json 2>/dev/null; echo "shard2 rc=$?"
python - <<'PY'
import json
for w in ("c5_8gpu","c5_4gpu","c5_shard4","c5_shard2"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r02_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "c4", d.get("c4") and (round(d["c4"]["value"]), d["c4"]["chains_per_gpu"], round(d["c4"]["tflops_per_gpu"],1)), "contr", round(d["getk"]["contraction_s"],4), d["getk"].get("row_panel_compute_exchange_s"))
    except Exception as e: print(w, "ERR", e)
PY
