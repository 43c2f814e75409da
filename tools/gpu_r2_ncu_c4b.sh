#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:t256_layer_bwd" -s 33 -c 1 -o gpurun_out/r2/t256_c4_bwd -f \
  python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/ncu_c4_bwd.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:t256_wgrad" -s 15 -c 1 -o gpurun_out/r2/t256_c4_wgrad -f \
  python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/ncu_c4_wgrad.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2/t256_c4_*.ncu-rep
