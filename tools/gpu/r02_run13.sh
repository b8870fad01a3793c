cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --workload c4 --chains 40 --steps 1 --warmup 3 --iters 20 --no-cpu-baseline"
$CMD > gpurun_out/r02_c4_40_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 40 -c 2 -o gpurun_out/r02_dgemm40b $CMD > gpurun_out/r02_ncu_dgemm40b.log 2>&1
echo "ncu rc=$?"
