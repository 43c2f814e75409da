#!/bin/bash
# ncu --set full captures (one launch each, -lineinfo source import) of the three d_model = 256 kernels on C4 and C3, batch 4096
# usage: gpu_r2b_ncu.sh TAG
TAG=${1:-s3}
mkdir -p gpurun_out/r2
for W in c4 c3; do
  for K in layer_fwd layer_bwd wgrad; do
    SKIP=14; [ "$K" = "wgrad" ] && SKIP=18
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:t256_${K}" -s $SKIP -c 1 -o gpurun_out/r2/t256_${W}_${K}_$TAG -f \
      python bench.py --workload $W --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/ncu_${W}_${K}_$TAG.log 2>&1
    echo "$W $K ncu rc=$?"
  done
done
ls -la gpurun_out/r2/t256_*_$TAG.ncu-rep
