#!/bin/bash
# usage: tools/run_bench_variants.sh  (on the GPU box)
for cs in 0 4 8; do
  BGGE_FUSED_CS=$cs python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py "c5 cs>=$cs"
done
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python tools/bench_summary.py c3
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python tools/bench_summary.py c2
python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python tools/bench_summary.py c1
