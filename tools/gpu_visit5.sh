#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_d256.py -m gpu -x -q > gpurun_out/pytest_v5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_v5.log
for stg in 0 15000 30000; do
GT_T256_STAGGER=$stg timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_s$stg.json 2> gpurun_out/bench_c4_s$stg.err; echo "bench stagger=$stg rc=$?"
python - $stg <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_c4_s%s.json"%sys.argv[1]))
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["final_loss"])
for k,v in d["kernels"].items(): print("  %-40s %6.1f %8.3f ms %.3f %s"%(k,v["launches_per_step"],v["ms_per_step"],v["share"], "%.0f TF"%v["tflops"] if "tflops" in v else ""))
PY
done
GT_T256_DBG=4 timeout 200 python bench.py --workload c4 --no-cpu-baseline --steps 1 --warmup 3 2>&1 >/dev/null | grep -A2 "bwd timeline" | head -12 > gpurun_out/t256_timeline_v5.txt; cat gpurun_out/t256_timeline_v5.txt
# small-batch regime (the reference's yaml batch sizes)
for b in 16 32 256; do timeout 200 python bench.py --workload c2 --batch $b --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2_b$b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_b$b.json')); print('c2 batch $b', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
