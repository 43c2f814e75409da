#!/bin/bash
# final evidence of the round for the headline command: launch list + --set full capture of the d_model = 32 layer kernels
mkdir -p gpurun_out/r3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3/launches_bf16_c2_b65536.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r3/ncu_ll.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_layer_(fwd|bwd)" -s 24 -c 2 -o gpurun_out/r3/tc32_c2_b65536 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r3/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r3/*.ncu-rep gpurun_out/r3/launches_bf16_c2_b65536.csv
