#!/bin/bash
mkdir -p gpurun_out/r2
python tools/diag_bf16_grid.py > gpurun_out/r2/diag_grid_p0.txt 2>&1
DIAG_P=0.2 python tools/diag_bf16_grid.py > gpurun_out/r2/diag_grid_p02.txt 2>&1
python tools/diag_hits.py > gpurun_out/r2/diag_hits.txt 2>&1
cat gpurun_out/r2/diag_grid_p0.txt; echo; cat gpurun_out/r2/diag_grid_p02.txt; echo; cat gpurun_out/r2/diag_hits.txt
