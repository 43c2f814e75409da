#!/bin/bash
# compute-sanitizer on what session 4 changed: the d_model = 32 path (coalesced flush through shared memory, edge kernels, PDL) in bf16
# and the split-mode generic GEMM (precision = fp32_tc).  Logs -> gpurun_out/r3/san_*.txt
mkdir -p gpurun_out/r3
for tool in memcheck racecheck synccheck initcheck; do
  SAN_PREC=bf16 timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_smoke.py c1 c2 c5 > gpurun_out/r3/san_${tool}_bf16.txt 2>&1
  echo "$tool bf16 rc=$? $(grep -c 'sanitize_smoke .* ok' gpurun_out/r3/san_${tool}_bf16.txt) cases ok; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r3/san_${tool}_bf16.txt | tail -1)"
done
for tool in memcheck racecheck synccheck; do
  SAN_PREC=fp32_tc timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_smoke.py c1 c3 c5 > gpurun_out/r3/san_${tool}_fp32_tc.txt 2>&1
  echo "$tool fp32_tc rc=$? $(grep -c 'sanitize_smoke .* ok' gpurun_out/r3/san_${tool}_fp32_tc.txt) cases ok; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r3/san_${tool}_fp32_tc.txt | tail -1)"
done
