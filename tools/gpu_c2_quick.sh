#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_gemm_tc.py tests/test_gpu_sweep.py -m gpu -x -q > gpurun_out/pytest_c2q.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_c2q.log
for wl in c2 c1 c5; do
timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_q.json 2> gpurun_out/bench_${wl}_q.err; echo "bench rc=$?"
python - $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_%s_q.json"%sys.argv[1]))
print(sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["achieved"], d["final_loss"])
for k,v in d["kernels"].items():
    if v["share"]>0.01: print("  %-40s %6.1f %8.3f ms %.3f"%(k,v["launches_per_step"],v["ms_per_step"],v["share"]))
PY
done
