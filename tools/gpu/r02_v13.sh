cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_getk.py tests/test_gpu_abi.py -x -q 2>&1 | tail -12 )
for th in 8; do
  if [ $th = 0 ]; then export BGGE_NO_STAGED_COPY=1; else unset BGGE_NO_STAGED_COPY; export BGGE_COPY_THREADS=$th; fi
  timeout 600 python bench.py --no-c4 --no-cpu-baseline --no-dense-ref --steps 2 --warmup 3 > gpurun_out/r02_v13_t$th.json 2>gpurun_out/r02_v13_t$th.err; echo "threads $th rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v13_t$th.json").read().strip().splitlines()[-1])
h=d["getk"]["host_buffers"]
print("threads $th base_call_s", round(h["base_call_s"],3), "expand_call_s", round(h["expand_one_kernel_call_s"],3), "diff", h["max_abs_diff_vs_device_pipeline"], "e2e", round(d["e2e"]["value"],1), "calls", [round(x,3) for x in d["e2e"]["call_s"]])
PY
done
