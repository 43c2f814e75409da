#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc" -s 3 -c 1 -o gpurun_out/gemm_tc_qkv -f python tools/gemm_tc_bench.py 256 512 4096 > gpurun_out/ncu_gemm_tc.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_gemm_tc.log
ls -la gpurun_out/*.ncu-rep
