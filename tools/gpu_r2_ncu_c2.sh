#!/bin/bash
# ncu --set full (+ source) of one tc_layer_fwd and one tc_layer_bwd launch of a C2 step; TAG names the variant
TAG=${1:-base}
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_layer_(fwd|bwd)" -s 17 -c 2 -o gpurun_out/r2/tc32_c2_$TAG -f \
  python bench.py --workload c2 --steps 1 --warmup 3 --batch 8192 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/ncu_c2_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2/*.ncu-rep
