#!/bin/bash
# One GPU-box visit: parity tests, bench lines (C2 = configs[1] headline, C4 = tensor-pipe target shape),
# ncu launch lists + full captures of the fused layer kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
cat gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/ncu_launch_c4.log 2>&1; echo "ncu launches c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:t256_(layer|wgrad)" -s 42 -c 4 -o gpurun_out/t256_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
ls -la gpurun_out
