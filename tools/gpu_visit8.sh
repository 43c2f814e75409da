#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_v8.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_v8.log
for prec in bf16 fp32; do
timeout 300 python bench.py --workload c5 --mode infer --precision $prec --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_infer_$prec.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_infer_$prec.json')); print('infer $prec', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])"
done
timeout 300 python bench.py --workload c5 --mode infer --batch 16384 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_infer_b16k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_infer_b16k.json')); print('infer b16384', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])"
