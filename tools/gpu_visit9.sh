#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -x -q > gpurun_out/pytest_v9.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_v9.log
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_v2.json 2> gpurun_out/bench_c3_v2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c3_v2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c3_v2.json"))
print(d["value"], d["ms_per_step"], d["final_loss"], d["path"], d["config"]["per_gpu_batch"])
for k,v in d["kernels"].items(): print("  %-40s %6.1f %8.3f ms %.3f"%(k,v["launches_per_step"],v["ms_per_step"],v["share"]))
PY
