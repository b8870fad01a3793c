cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t4.log
tail -30 gpurun_out/r02_t4.log
bash tools/sanitize.sh memcheck
