#!/bin/bash
# generic tcgen05 GEMM path: unit + model parity tests, then C3 / C5 bench lines (bf16 via gemm_tc vs the fp32 SIMT path)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_bf16.py -m gpu -x -q > gpurun_out/pytest_gemm_tc.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gemm_tc.log
for wl in c3 c5; do
  for prec in bf16 fp32; do
    timeout 300 python bench.py --workload $wl --precision $prec --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_${prec}.json 2> gpurun_out/bench_${wl}_${prec}.err; echo "bench $wl $prec rc=$?"
    cut -c1-3000 gpurun_out/bench_${wl}_${prec}.json; tail -2 gpurun_out/bench_${wl}_${prec}.err
  done
done
