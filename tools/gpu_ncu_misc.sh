#!/bin/bash
# one metrics pass over the smaller kernels (ends of the stack, per-op attention / GEMM, decode step, weight gradients) for the kernel
# table of DESIGN.md.  CSV logs only (a --set full report of these many launches exceeds what gpurun copies back).
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size"
timeout 600 ncu --metrics $M --clock-control none -k "regex:tail32|stem32" -s 8 -c 4 --csv --log-file gpurun_out/misc_edges32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m1.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k "regex:tail256|stem256|wgrad|t256_layer" -s 40 -c 12 --csv --log-file gpurun_out/misc_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m2.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k "regex:attn_mma|gemm_tc|ln_" -s 100 -c 24 --csv --log-file gpurun_out/misc_c3.csv python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m3.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k "regex:dec32" -s 200 -c 3 --csv --log-file gpurun_out/misc_dec.csv python bench.py --workload c5 --mode infer --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m4.log 2>&1; echo "rc=$?"
ls -la gpurun_out/misc_*.csv
