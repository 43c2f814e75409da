#!/bin/bash
# d_model = 256 iteration loop: parity tests, C4 bench line with the per-class breakdown, backward timeline of CTA 0
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_d256.py -m gpu -x -q > gpurun_out/pytest_c4q.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_c4q.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_q.json 2> gpurun_out/bench_c4_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c4_q.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["final_loss"])
for k,v in d["kernels"].items(): print("  %-40s %6.1f %8.3f ms %.3f %s"%(k,v["launches_per_step"],v["ms_per_step"],v["share"], "%.0f TF"%v["tflops"] if "tflops" in v else ""))
PY
GT_T256_DBG=6 timeout 200 python bench.py --workload c4 --no-cpu-baseline --steps 1 --warmup 3 2>&1 >/dev/null | grep -A2 "timeline" | head -24 > gpurun_out/t256_timeline_q.txt; cat gpurun_out/t256_timeline_q.txt
