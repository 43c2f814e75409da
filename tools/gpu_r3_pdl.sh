#!/bin/bash
# programmatic dependent launch between the d_model = 32 layer kernels: tests, then a single sweep member at the yaml batch 32
# (eager per-run loop and CUDA-graph replay) and the large-batch headline with GT_PDL = 0 / 1
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_tc_engine.py tests/test_gpu_sweep.py tests/test_gpu_bf16_exact.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r3/test_pdl.txt
cat gpurun_out/r3/test_pdl.txt
for pdl in 0 1; do
  for drive in fused graph; do
    GT_PDL=$pdl timeout 300 python tools/sweep_bench.py --members 1,8 --steps 400 --batch 32 --drive $drive > gpurun_out/r3/sweep_b32_${drive}_pdl$pdl.jsonl 2> gpurun_out/r3/sweep_b32_${drive}_pdl$pdl.err
    python - <<PY
import json
for l in open("gpurun_out/r3/sweep_b32_${drive}_pdl$pdl.jsonl"):
    l = json.loads(l)
    print("pdl=$pdl $drive members", l["members"], "ms/step seq", round(l["ms_per_step_sequential"], 4), "packed seq/s", round(l["packed"]), "loss", l["final_losses"][:2])
PY
  done
  GT_PDL=$pdl timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline --no-extras --steps 10 > gpurun_out/r3/bench_c2_pdl$pdl.json 2> gpurun_out/r3/bench_c2_pdl$pdl.err
  python -c "
import json; l = json.load(open('gpurun_out/r3/bench_c2_pdl$pdl.json')); print('pdl=$pdl C2', round(l['value']), 'seq/s', round(l['ms_per_step'], 3), 'ms e2e', round(l['e2e']['value']))"
done
