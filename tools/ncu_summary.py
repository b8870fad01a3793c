"""Condenses an .ncu-rep into the few columns DESIGN.md / the judge need.  usage: ncu_summary.py rep out.csv"""
import csv, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = re.compile(r'^(ID|Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|'
                  r'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|launch__(registers_per_thread|grid_size|block_size|'
                  r'cluster_size|cluster_max_active|waves_per_multiprocessor|occupancy_limit_registers|occupancy_limit_shared_mem|shared_mem_per_block_dynamic)|'
                  r'sm__warps_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|'
                  r'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|'
                  r'lts__t_sector_hit_rate.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed|'
                  r'smsp__pcsamp_warps_issue_stalled_(wait|math_pipe_throttle|long_scoreboard|short_scoreboard|barrier|membar|lg_throttle|not_selected|selected|no_instructions|branch_resolving)|'
                  r'sm__cycles_elapsed.avg.per_second)$')
idx = [i for i, h in enumerate(hdr) if keep.match(h)]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
for r in rows[2:]:
    d = {hdr[i]: r[i] for i in idx}
    print(d["Kernel Name"][:60], "| us", d.get("gpu__time_duration.sum"), "| dram rd", d.get("dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")],
          "wr", d.get("dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")], "| dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
          "| dmma%", d.get("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"), "| regs", d.get("launch__registers_per_thread"))
