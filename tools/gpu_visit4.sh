#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/pytest_v4.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_v4.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_v5.json 2> gpurun_out/bench_c2_v5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2_v5.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2_v5.json"))
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["achieved"], d["final_loss"])
for k,v in d["kernels"].items(): print("  %-40s %6.1f %8.3f ms %.3f"%(k,v["launches_per_step"],v["ms_per_step"],v["share"]))
PY
timeout 300 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c1.json"))
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["achieved"], d["final_loss"])
for k,v in d["kernels"].items(): print("  %-40s %6.1f %8.3f ms %.3f"%(k,v["launches_per_step"],v["ms_per_step"],v["share"]))
PY
