cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
( cd tools/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && timeout 60 ./tmem_probe ) > gpurun_out/r02_tmem_probe.log 2>&1; echo "tmem probe rc=$?"; cat gpurun_out/r02_tmem_probe.log
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_t7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_t7.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_c5_${N}gpu.json 2> gpurun_out/r02_c5_${N}gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_c5_${N}gpu.err | cut -c1-300
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02_c5_{N}gpu.json").read().strip().splitlines()[-1])
for k in ("value","n_gpus","e2e","c4","clocks"): print(k, json.dumps(d.get(k))[:600])
print(json.dumps(d["getk"])[:900])
PY
