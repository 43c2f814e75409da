#!/bin/bash
# round 2, GPU call 1: bf16-oracle diagnostics, full GPU suite, default bench (new extras), first sanitizer pass
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/gpu.txt
python tools/diag_bf16_exact.py > gpurun_out/r2/diag_bf16_exact.txt 2>&1
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2/test_all.txt 2>&1
tail -5 gpurun_out/r2/test_all.txt
python bench.py > gpurun_out/r2/bench_c2.json 2> gpurun_out/r2/bench_c2.err
cat gpurun_out/r2/bench_c2.json | head -c 3000
SAN_PREC=bf16 timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py c1 c4 > gpurun_out/r2/san_memcheck_bf16_c1_c4.txt 2>&1
tail -3 gpurun_out/r2/san_memcheck_bf16_c1_c4.txt
cat gpurun_out/r2/diag_bf16_exact.txt
