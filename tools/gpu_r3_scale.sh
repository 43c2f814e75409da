#!/bin/bash
# weak-scaling lines (per-GPU batch fixed) on ONE 8-GPU box, same session: N = 1 and N = 8 (and the N given in $NS) per workload
mkdir -p gpurun_out/r3
NS=${NS:-"1 8"}
for w in ${WL:-c4 c3 c2}; do
  for n in $NS; do
    out=gpurun_out/r3/scale_${w}_dp${n}.json
    if [ "$n" = "1" ]; then
      python bench.py --workload $w --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > $out 2> gpurun_out/r3/scale_${w}_dp${n}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --workload $w --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > $out 2> gpurun_out/r3/scale_${w}_dp${n}.err
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
