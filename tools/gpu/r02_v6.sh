cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02_v6_tests.log; cat gpurun_out/r02_v6_tests.log
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 600 python tools/sanitize_cases.py > gpurun_out/r02_v6_checks.log 2>&1; echo "checks rc=$?"; tail -4 gpurun_out/r02_v6_checks.log
