# usage: bash tools/gpu/bench_scale.sh N   (under gpurun --gpus N)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_r02_c5_${N}gpu.json 2> gpurun_out/bench_r02_c5_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02_c5_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c5_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "c4", round(d["c4"]["value"]), "chains/gpu", d["c4"]["chains_per_gpu"], "getk contraction", d["getk"]["contraction_s"], d["getk"]["row_panel_compute_exchange_s"])
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
