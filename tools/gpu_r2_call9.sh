#!/bin/bash
mkdir -p gpurun_out/r2
for prec in bf16; do
  SAN_PREC=$prec timeout 420 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_smoke.py c1 c2 c4 c3 c5 > gpurun_out/r2/san_racecheck_${prec}_v2.txt 2>&1
  echo "racecheck $prec rc=$? $(grep -c 'sanitize_smoke .* ok' gpurun_out/r2/san_racecheck_${prec}_v2.txt) cases ok; $(grep 'RACECHECK SUMMARY' gpurun_out/r2/san_racecheck_${prec}_v2.txt | tail -1)"
  SAN_PREC=$prec timeout 300 compute-sanitizer --tool synccheck --print-limit 30 python tools/sanitize_smoke.py c1 c2 c5 > gpurun_out/r2/san_synccheck_${prec}_v2.txt 2>&1
  echo "synccheck $prec rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r2/san_synccheck_${prec}_v2.txt | tail -1)"
done
bash tools/gpu_r2_quick.sh s3
