#!/bin/bash
# A/B of an environment knob on the C3 / C4 bench lines: gpu_r2_ab.sh TAG VAR "v1 v2 ..."
TAG=$1; VAR=$2; VALS=$3
mkdir -p gpurun_out/r2
for W in c4 c3; do
for S in $VALS; do
  env $VAR=$S timeout -s KILL 300 python bench.py --workload $W --no-cpu-baseline --no-eager-baseline --no-extras --steps 8 > gpurun_out/r2/bench_${W}_${TAG}_$S.json 2> gpurun_out/r2/bench_${W}_${TAG}_$S.err
  python - <<PY
import json
l=json.load(open('gpurun_out/r2/bench_${W}_${TAG}_$S.json'))
print('$VAR=$S: $W seq/s', round(l['value']), {k:round(v['ms_per_step'],2) for k,v in l.get('kernels',{}).items() if 'layer' in k or 'wgrad' in k})
PY
done
done
