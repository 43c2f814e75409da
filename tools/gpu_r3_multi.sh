#!/bin/bash
# 2-GPU box at the end of the round: the NCCL-path equivalence test (skipped on the driver's 1-GPU box) + dp1 / dp2 bench lines of C2 and C4
mkdir -p gpurun_out/r3
python -m pytest tests/test_gpu_multi.py -q -rs --no-header -p no:cacheprovider > gpurun_out/r3/test_multi_2gpu.txt 2>&1; tail -4 gpurun_out/r3/test_multi_2gpu.txt
for w in ${WL:-c2 c4}; do
  for n in 1 2; do
    out=gpurun_out/r3/scale_${w}_dp${n}${TAG}.json
    if [ "$n" = "1" ]; then
      python bench.py --workload $w --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > $out 2> gpurun_out/r3/scale_${w}_dp${n}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --workload $w --gpus $n --exchange ${EXCH:-auto} --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > $out 2> gpurun_out/r3/scale_${w}_dp${n}.err
    fi
    python - <<PY
import json
try:
    l=json.load(open('$out')); print('$w N=$n', round(l['value']), 'seq/s', round(l['ms_per_step'],2), 'ms/step  e2e', round(l['e2e']['value']), l['clocks'])
except Exception as e:
    print('$w N=$n FAILED', e); print(open('gpurun_out/r3/scale_${w}_dp${n}.err').read()[-1500:])
PY
  done
done
