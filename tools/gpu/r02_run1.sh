cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_benchshapes.py tests/test_gpu_factored.py -x -q > gpurun_out/r02_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t1.log
BGGE_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense-ref > gpurun_out/r02_c5_a.json 2> gpurun_out/r02_c5_a.err; echo "bench rc=$?" >> gpurun_out/r02_t1.log
tail -5 gpurun_out/r02_t1.log; tail -c 1500 gpurun_out/r02_c5_a.json
