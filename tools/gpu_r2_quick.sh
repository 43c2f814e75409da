#!/bin/bash
# quick GPU regression: exact bf16 tests, the whole GPU suite (stop at 8 failures), C2 bench without baselines.  TAG names the outputs.
TAG=${1:-q}
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_bf16_exact.py -q -rf --no-header -p no:cacheprovider > gpurun_out/r2/test_exact_$TAG.txt 2>&1
grep -n "passed\|failed\|^E   .*assert\|mismatch" gpurun_out/r2/test_exact_$TAG.txt | head -30
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider --maxfail=8 --deselect tests/test_gpu_bf16_exact.py > gpurun_out/r2/test_all_$TAG.txt 2>&1
tail -12 gpurun_out/r2/test_all_$TAG.txt
python bench.py --no-cpu-baseline --no-eager-baseline --no-extras --steps 10 > gpurun_out/r2/bench_c2_$TAG.json 2> gpurun_out/r2/bench_c2_$TAG.err
python - <<PY
import json
l=json.load(open('gpurun_out/r2/bench_c2_$TAG.json'))
print('C2 train seq/s', round(l['value']), 'ms/step', round(l['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in l['kernels'].items()})
PY
