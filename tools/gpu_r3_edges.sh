#!/bin/bash
# edge32 kernels (stem / tail of the d_model = 32 path): parity tests + the per-class breakdown of the headline step
mkdir -p gpurun_out/r3
TAG=${1:-v}
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_bf16_exact.py tests/test_tc_engine.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -4
for w in c2 c1 c5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-eager-baseline --no-extras --steps 10 > gpurun_out/r3/bench_${w}_edges_$TAG.json 2> gpurun_out/r3/bench_${w}_edges_$TAG.err
  python - <<PY
import json
l = json.load(open("gpurun_out/r3/bench_${w}_edges_$TAG.json"))
print("$w", round(l["value"]), "seq/s", round(l["ms_per_step"], 3), "ms", {k.split(" ")[0]: round(v["ms_per_step"], 3) for k, v in l["kernels"].items()})
PY
done
