#!/bin/bash
# 2-GPU box: the NCCL-path equivalence test (skipped on the driver's 1-GPU box) + the whole GPU suite + a dp2 bench line
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_multi.py -q -rs --no-header -p no:cacheprovider > gpurun_out/r2/test_multi_2gpu.txt 2>&1; tail -5 gpurun_out/r2/test_multi_2gpu.txt
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider --maxfail=10 > gpurun_out/r2/test_all_2gpu.txt 2>&1; tail -8 gpurun_out/r2/test_all_2gpu.txt
WL=c4 NS="2" bash tools/gpu_r2_scale.sh
