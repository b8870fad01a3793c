cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_r02_c5_${N}gpu.json 2> gpurun_out/bench_r02_c5_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c5_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "c4", round(d["c4"]["value"]), "chains/gpu", d["c4"]["chains_per_gpu"], "getk contraction", d["getk"]["contraction_s"], d["getk"]["row_panel_compute_exchange_s"])
PY
