#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
for w in c5 c2 c4; do
timeout 300 python bench.py --workload $w --mode infer --steps 5 > gpurun_out/bench_infer_$w.json 2> gpurun_out/bench_infer_$w.err; echo "infer $w rc=$?"; cat gpurun_out/bench_infer_$w.json; tail -3 gpurun_out/bench_infer_$w.err
done
GT_T256_DBG=4 timeout 300 python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > /dev/null 2> gpurun_out/t256_timeline.txt; echo "timeline rc=$?"
head -c 6000 gpurun_out/t256_timeline.txt
