cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
BGGE_LIB=$PWD/bgge_b200/libbgge_b200_checks.so timeout 900 python tools/sanitize_cases.py > gpurun_out/r02_v12_checks.log 2>&1; echo "checks rc=$?"; tail -22 gpurun_out/r02_v12_checks.log
timeout 900 python tools/sanitize_cases.py > gpurun_out/r02_v12_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r02_v12_plain.log
