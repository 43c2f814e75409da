#!/bin/bash
# launch list of the headline bench command + one --set full capture of the dominant kernel at the headline batch
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2_b32768.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_layer_bwd" -s 8 -c 1 -o gpurun_out/tc32_bwd_b32768 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k "regex:tc_layer_fwd" -s 8 -c 1 -o gpurun_out/tc32_fwd_b32768 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu full fwd rc=$?"
timeout 600 ncu --set full --clock-control none -k "regex:t256_layer_bwd" -s 12 -c 1 -o gpurun_out/t256_bwd_b8192 -f python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1; echo "ncu full c4 rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_c2_b32768.csv
