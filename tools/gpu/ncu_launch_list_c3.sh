cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --workload c3 --no-c4 --no-getk-host --no-cpu-baseline --no-dense-ref --steps 1 --warmup 3"
timeout 600 $CMD > gpurun_out/launches_c3_plain.json 2>gpurun_out/launches_c3_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|fused_finalize|scalar_kernel|virt_gather" -c 200 --csv --log-file gpurun_out/r02_launches_c3.csv $CMD > gpurun_out/launches_c3_ncu.log 2>&1; echo "ncu rc=$?"
