cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_getk.py tests/test_gpu_abi.py tests/test_gpu_factored.py -x -q 2>&1 | tail -8 )
timeout 600 python tools/microbench/expand_bw.py 2>&1 | tail -8
