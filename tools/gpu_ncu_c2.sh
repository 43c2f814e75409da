#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_layer_(fwd|bwd)" -s 17 -c 2 -o gpurun_out/tc32_c2 -f python bench.py --workload c2 --steps 1 --warmup 3 --batch 8192 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
