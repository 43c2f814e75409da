#!/bin/bash
# compute-sanitizer on smoke-sized steps (tools/sanitize_smoke.py): C1 / C2 (d_model 32 fused), C4-l2 (d_model 256 fused), C3-l1 (d_model 256 fused, head_dim 128),
# C5 encoder-decoder; fp32 and bf16.  Logs -> gpurun_out/r2/san_*.txt
mkdir -p gpurun_out/r2
for tool in memcheck racecheck synccheck initcheck; do
  for prec in ${PRECS:-bf16 fp32}; do
    SAN_PREC=$prec timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_smoke.py c1 c2 c4 c3 c5 > gpurun_out/r2/san_${tool}_${prec}.txt 2>&1
    echo "$tool $prec rc=$? $(grep -c 'sanitize_smoke .* ok' gpurun_out/r2/san_${tool}_${prec}.txt) cases ok; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2/san_${tool}_${prec}.txt | tail -1)"
  done
done
