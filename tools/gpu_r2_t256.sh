#!/bin/bash
# d_model = 256 fused kernels (C3 head_dim 128 and C4 head_dim 16): parity tests bounded by timeout (a wrong mbarrier phase hangs),
# C3 / C4 bench lines with the per-class breakdown, clock64 phase time line of CTA 0 from the developer library.
# usage: gpu_r2_t256.sh TAG [notest]
TAG=${1:-v}
mkdir -p gpurun_out/r2
if [ "$2" != "notest" ]; then
  timeout -s KILL 900 python -m pytest tests/test_gpu_bf16_d256.py tests/test_gpu_bf16_exact.py -q -rf --no-header -p no:cacheprovider -x > gpurun_out/r2/test_t256_$TAG.txt 2>&1
  echo "pytest rc $?"
  grep -n "passed\|failed\|^E   .*assert\|mismatch\|Error" gpurun_out/r2/test_t256_$TAG.txt | head -30
fi
for W in c3 c4; do
  timeout -s KILL 600 python bench.py --workload $W --no-cpu-baseline --no-eager-baseline --no-extras --steps 6 > gpurun_out/r2/bench_${W}_$TAG.json 2> gpurun_out/r2/bench_${W}_$TAG.err
  echo "bench $W rc $?"
  python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r2/bench_${W}_$TAG.json'))
    print('$W train seq/s', round(l['value']), 'ms/step', round(l['ms_per_step'],2), 'loss', l.get('final_loss'), {k:round(v['ms_per_step'],2) for k,v in l.get('kernels',{}).items()})
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/r2/bench_${W}_$TAG.err').read()[-2000:])
PY
done
export GROOVE_B200_DEV_FLAGS="-DGT_T256_TIMELINE"
for W in c3 c4; do
  GT_T256_DBG=3 timeout -s KILL 300 python tools/t256_timeline.py $W 4096 > gpurun_out/r2/timeline_${W}_$TAG.txt 2>&1
  grep -A0 "bwd timeline" gpurun_out/r2/timeline_${W}_$TAG.txt | tail -1
done
