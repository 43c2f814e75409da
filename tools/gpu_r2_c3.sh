#!/bin/bash
# C3 (head_dim 128) on the fused d_model = 256 kernels: targeted parity tests (bounded by timeout: a wrong mbarrier phase hangs), then the C3 bench
TAG=${1:-c3}
mkdir -p gpurun_out/r2
timeout -s KILL 600 python -m pytest tests/test_gpu_bf16_d256.py tests/test_gpu_bf16_exact.py tests/test_gpu_gemm_tc.py -k "c3 or h2_f64 or path_kind or same_masks" -q -rf --no-header -p no:cacheprovider -x > gpurun_out/r2/test_c3_$TAG.txt 2>&1
echo "pytest rc $?"
grep -n "passed\|failed\|^E   .*assert\|mismatch\|Error" gpurun_out/r2/test_c3_$TAG.txt | head -30
tail -5 gpurun_out/r2/test_c3_$TAG.txt
timeout -s KILL 600 python bench.py --workload c3 --no-cpu-baseline --no-eager-baseline --no-extras --steps 6 > gpurun_out/r2/bench_c3_$TAG.json 2> gpurun_out/r2/bench_c3_$TAG.err
echo "bench rc $?"
python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r2/bench_c3_$TAG.json'))
    print('C3 train seq/s', round(l['value']), 'ms/step', round(l['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in l.get('kernels',{}).items()})
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/r2/bench_c3_$TAG.err').read()[-2000:])
PY
