#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_bf16_exact.py -q -rf --no-header -p no:cacheprovider --durations=8 > gpurun_out/r2/test_exact_v2.txt 2>&1
grep -n "passed\|failed\|AssertionError\|^E  " gpurun_out/r2/test_exact_v2.txt | head -60
tail -15 gpurun_out/r2/test_exact_v2.txt
