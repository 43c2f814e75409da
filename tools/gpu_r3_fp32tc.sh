#!/bin/bash
# precision = fp32_tc: parity tests + C3 / C4 / C2 step rates next to precision = fp32 (FFMA) on the same batch
mkdir -p gpurun_out/r3
python tools/diag_fp32_tc.py 2>&1 | grep -v Warning > gpurun_out/r3/diag_fp32_tc_grads.txt
grep fp32_tc gpurun_out/r3/diag_fp32_tc_grads.txt | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_fp32_tc.py -q -m gpu 2>&1 | grep -v "^E  *\[\|^E    *[0-9-]" | tail -60 > gpurun_out/r3/test_fp32_tc.txt
cat gpurun_out/r3/test_fp32_tc.txt
for w in ${WL:-c3 c4 c2}; do
  for p in ${PRECS:-fp32 fp32_tc}; do
    timeout 600 python bench.py --workload $w --precision $p --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras \
      > gpurun_out/r3/bench_${w}_${p}_b4096.json 2> gpurun_out/r3/bench_${w}_${p}_b4096.err
    python - <<PY
import json
try:
    l = json.load(open("gpurun_out/r3/bench_${w}_${p}_b4096.json"))
    print("$w $p", round(l["value"]), "seq/s", round(l["ms_per_step"], 2), "ms", l.get("path"), "loss", l.get("final_loss"))
except Exception as e:
    print("$w $p ERR", e)
PY
  done
done
