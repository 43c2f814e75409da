#!/bin/bash
# peer-memory gradient exchange (csrc/peer_opt.cu) on an N-GPU box: equivalence check, then bench lines with exchange = nccl / p2p
N=${N:-2}
mkdir -p gpurun_out/r3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $(( N < 4 ? N : 4 )) --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > gpurun_out/r3/dp_check_${N}gpu.txt 2>&1
echo "dp_check rc=$?"; grep "DP_CHECK\|Error\|error" gpurun_out/r3/dp_check_${N}gpu.txt | tail -12
for w in ${WL:-c4 c3 c2}; do
  for ex in nccl p2p; do
    out=gpurun_out/r3/scale_${w}_dp${N}_${ex}.json
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --workload $w --gpus $N --exchange $ex --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > $out 2> gpurun_out/r3/scale_${w}_dp${N}_${ex}.err
    python - <<PY
import json
try:
    l=json.load(open('$out')); print('$w N=$N $ex', round(l['value']), 'seq/s', round(l['ms_per_step'],3), 'ms/step  e2e', round(l['e2e']['value']), l['config'].get('exchange','')[:40], l['clocks']['reasons'])
except Exception as e:
    print('$w N=$N $ex FAILED', e); print(open('gpurun_out/r3/scale_${w}_dp${N}_${ex}.err').read()[-1500:])
PY
  done
done
