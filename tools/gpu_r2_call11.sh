#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_sweep.py -q -rf --no-header -p no:cacheprovider -k "graph" > gpurun_out/r2/test_sweep.txt 2>&1
tail -15 gpurun_out/r2/test_sweep.txt
