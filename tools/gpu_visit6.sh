#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_sweep.log
timeout 600 python tools/sweep_bench.py --members 1,4,8,16,32 --batch 32 2>gpurun_out/sweep_c2.err | tee gpurun_out/sweep_c2_b32.jsonl; tail -3 gpurun_out/sweep_c2.err
timeout 600 python tools/sweep_bench.py --members 8,16 --spec closedhh --steps 30 2>gpurun_out/sweep_hh.err | tee gpurun_out/sweep_closedhh.jsonl; tail -3 gpurun_out/sweep_hh.err
