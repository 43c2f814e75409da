#!/bin/bash
# quick visit: bf16 d_model=32 parity tests + pipeline tests + C2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_quick.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-2600 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
