#!/bin/bash
# final ncu evidence for the headline workload (C2 at the bench batch): launch list of one bench command + --set full of one fwd / one bwd launch
mkdir -p gpurun_out/r2
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2/launches_c2_b65536.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/launches_c2.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_layer_(fwd|bwd)" -s 17 -c 2 -o gpurun_out/r2/tc32_c2_b65536_final -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r2/ncu_c2_final.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2/tc32_c2_b65536_final.ncu-rep
