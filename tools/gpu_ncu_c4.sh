#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:t256_layer_bwd" -s 5 -c 1 -o gpurun_out/t256_bwd_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
