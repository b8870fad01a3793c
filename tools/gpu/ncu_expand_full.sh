cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python tools/microbench/expand_bw.py > gpurun_out/r02_expand_bw.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"expand2_kernel|expand_kernel" -c 4 -o gpurun_out/r02_expand -f python tools/microbench/expand_bw.py > gpurun_out/r02_ncu_expand.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_expand.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread > gpurun_out/r02_ncu_expand_summary.csv 2>/dev/null; cat gpurun_out/r02_ncu_expand_summary.csv | cut -c1-400
