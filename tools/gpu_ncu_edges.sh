#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tail32|stem32" -s 8 -c 4 -o gpurun_out/edges_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_edges.log 2>&1; echo "ncu rc=$?"
