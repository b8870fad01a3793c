cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for CH in 40 320; do
CMD="python bench.py --workload c4 --chains $CH --steps 1 --warmup 3 --iters 10 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/r02_ncu_dgemm_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 20 -c 2 -f -o gpurun_out/r02_dgemm_final_$CH $CMD > gpurun_out/r02_ncu_dgemm_$CH.log 2>&1; echo "ncu $CH rc=$?"
done
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_warps_issue_stalled_math_pipe_throttle,smsp__pcsamp_warps_issue_stalled_short_scoreboard,smsp__pcsamp_warps_issue_stalled_wait,smsp__pcsamp_warps_issue_stalled_selected,smsp__pcsamp_warps_issue_stalled_barrier
for CH in 40 320; do ncu -i gpurun_out/r02_dgemm_final_$CH.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_dgemm_final_${CH}_summary.csv 2>/dev/null; done
cut -d, -f5,12- gpurun_out/r02_ncu_dgemm_final_40_summary.csv | cut -c1-700
cut -d, -f5,12- gpurun_out/r02_ncu_dgemm_final_320_summary.csv | tail -2 | cut -c1-700
