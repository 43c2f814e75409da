#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): DP equivalence test + weak-scaling bench lines at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_multi.log
for w in c2 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w > gpurun_out/bench_${w}_dp$N.json 2> gpurun_out/bench_${w}_dp$N.err; echo "bench $w dp$N rc=$?"
  cat gpurun_out/bench_${w}_dp$N.json; tail -3 gpurun_out/bench_${w}_dp$N.err
done
