cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --iters 5 --prof-iters 2 --no-cpu-baseline --no-c4 --no-dense-ref"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|fused_finalize|scalar_kernel|virt_gather" -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
