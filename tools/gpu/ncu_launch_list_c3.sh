cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python tools/microbench/expand_bw.py > gpurun_out/r02_expand_bw.log 2>&1; echo "expand rc=$?"; cat gpurun_out/r02_expand_bw.log | tail -12
( timeout 600 python -m pytest tests/test_gpu_getk.py -x -q 2>&1 | tail -3 )
timeout 600 python bench.py --workload c3 --no-c4 --no-getk-host --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r02_v7_c3.json 2>gpurun_out/r02_v7_c3.err; echo "c3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|fused_finalize|scalar_kernel|virt_gather" -c 200 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --workload c3 --no-c4 --no-getk-host --no-cpu-baseline --no-dense-ref --steps 1 --warmup 3 > gpurun_out/r02_v7_ncu.log 2>&1; echo "ncu rc=$?"
tail -40 gpurun_out/r02_launches_c3.csv | cut -d, -f5,12- | tail -30
