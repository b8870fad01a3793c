cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --iters 5 --prof-iters 2 --no-cpu-baseline --no-dense-ref"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_mrhs -s 6 -c 2 -o gpurun_out/r02_mrhs $CMD > gpurun_out/r02_ncu_mrhs.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r02_ncu_mrhs.log
