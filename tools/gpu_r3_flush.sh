#!/bin/bash
# coalesced weight-gradient flush of the d_model = 32 backward kernel: parity tests, single member at the yaml batch, headline batch
mkdir -p gpurun_out/r3
TAG=${1:-v}
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_bf16_exact.py tests/test_tc_engine.py tests/test_gpu_gemm_tc.py tests/test_gpu_sweep.py -q -m gpu -x 2>&1 | tail -4
for drive in fused graph; do
  timeout 300 python tools/sweep_bench.py --members 1,8 --steps 400 --batch 32 --drive $drive > gpurun_out/r3/sweep_b32_${drive}_$TAG.jsonl 2> gpurun_out/r3/sweep_b32_${drive}_$TAG.err
  python - <<PY
import json
for l in open("gpurun_out/r3/sweep_b32_${drive}_$TAG.jsonl"):
    l = json.loads(l)
    print("$drive members", l["members"], "ms/step seq", round(l["ms_per_step_sequential"], 4), "packed seq/s", round(l["packed"]), "loss", l["final_losses"][:2])
PY
done
for w in c2 c5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-eager-baseline --no-extras --steps 10 > gpurun_out/r3/bench_${w}_$TAG.json 2> gpurun_out/r3/bench_${w}_$TAG.err
  python - <<PY
import json
l = json.load(open("gpurun_out/r3/bench_${w}_$TAG.json"))
print("$w", round(l["value"]), "seq/s", round(l["ms_per_step"], 3), "ms", {k.split(" ")[0]: round(v["ms_per_step"], 3) for k, v in l["kernels"].items()})
PY
done
