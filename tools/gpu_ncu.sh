#!/bin/bash
# ncu --set full captures (with source) of the fused layer kernels: C2 (d_model 32) and C4 (d_model 256).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_layer_(fwd|bwd)" -s 17 -c 2 -o gpurun_out/tc32_c2 -f python bench.py --workload c2 --steps 1 --warmup 3 --batch 8192 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:t256_(layer|wgrad)" -s 43 -c 3 -o gpurun_out/t256_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2.json | cut -c1-1800; tail -3 gpurun_out/bench_c2.err
ls -la gpurun_out
