#!/bin/bash
mkdir -p gpurun_out/r2
python tools/diag_bf16_exact.py > gpurun_out/r2/diag_bf16_exact_v2.txt 2>&1
python tools/diag_hits.py > gpurun_out/r2/diag_hits_v2.txt 2>&1
cat gpurun_out/r2/diag_bf16_exact_v2.txt; echo; cat gpurun_out/r2/diag_hits_v2.txt
