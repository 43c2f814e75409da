#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_v7.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_v7.log
timeout 300 python bench.py --workload c5 --precision bf16 --batch 4096 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_v2.json 2> gpurun_out/bench_c5_v2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5_v2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c5_v2.json"))
print(d["value"], d["ms_per_step"], d["final_loss"], d["path"])
for k,v in d["kernels"].items(): print("  %-40s %6.1f %8.3f ms %.3f"%(k,v["launches_per_step"],v["ms_per_step"],v["share"]))
PY
timeout 300 python bench.py --workload c5 --mode infer --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_infer_v2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_infer_v2.json')); print('infer', d['value'], d['ms_per_step'], d['config']['precision'])"
