#!/bin/bash
TAG=${1:-base}
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:t256_(layer|wgrad)" -s 43 -c 3 -o gpurun_out/r2/t256_c4_$TAG -f \
  python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/ncu_c4_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2/t256_c4_$TAG.ncu-rep
