#!/bin/bash
mkdir -p gpurun_out/r2
for W in c4 c3; do
for S in 0 8000 16000; do
  GT_T256_STAGGER_FWD=$S timeout -s KILL 300 python bench.py --workload $W --no-cpu-baseline --no-eager-baseline --no-extras --steps 6 > gpurun_out/r2/bench_${W}_stagf$S.json 2> gpurun_out/r2/bench_${W}_stagf$S.err
  python - <<PY
import json
l=json.load(open('gpurun_out/r2/bench_${W}_stagf$S.json'))
print('fwd stagger $S: $W seq/s', round(l['value']), {k:round(v['ms_per_step'],2) for k,v in l.get('kernels',{}).items() if 'layer' in k})
PY
done
done
