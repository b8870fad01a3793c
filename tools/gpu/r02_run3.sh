cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_t3.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_getk.py -x -q >> gpurun_out/r02_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t3.log
tail -15 gpurun_out/r02_t3.log
