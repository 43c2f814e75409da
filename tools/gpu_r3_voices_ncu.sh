#!/bin/bash
# voices tests, flakiness check of the fp32 Adam trajectory of c5, and an ncu --set full capture of the split-mode generic GEMM
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_voices.py "tests/test_gpu_parity.py" "tests/test_gpu_fp32_tc.py" -q -m gpu -k "voices or widths or host_predict" 2>&1 | grep -v "^E  *\[\|^E    *[0-9-]" | tail -30
for i in 1 2 3 4; do timeout 300 python -m pytest "tests/test_gpu_parity.py::test_loss_trajectory_matches_reference" -q -m gpu -k "adam-c5" 2>&1 | grep -E "passed|failed|Max absolute" ; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 60 -c 12 -f -o gpurun_out/r3/gemm_tc_split_c3 \
  python bench.py --workload c3 --precision fp32_tc --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r3/ncu_gemm_split.log 2>&1
tail -3 gpurun_out/r3/ncu_gemm_split.log
ls -la gpurun_out/r3/*.ncu-rep
