#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_evaluator.py -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_v2.log
timeout 200 python tools/gemm_tc_bench.py 256 512 4096 2>&1 | tee gpurun_out/gemm_tc_bench_c3.txt
timeout 200 python tools/gemm_tc_bench.py 32 512 8192 2>&1 | tee gpurun_out/gemm_tc_bench_c5.txt
timeout 300 python bench.py --workload c3 --precision bf16 --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_bf16_v2.json 2> gpurun_out/bench_c3_bf16_v2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c3_bf16_v2.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"])
for k,v in d["kernels"].items(): print("  %-20s %6.1f %8.3f ms %.3f"%(k,v["launches_per_step"],v["ms_per_step"],v["share"]))
PY
