#!/bin/bash
# session 9: gemm_tc fast path / vector epilogue — unit tests of the generic GEMM, the per-op path parity tests, C3 bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py tests/test_tc_engine.py -m gpu -x -q 2>&1 | tail -8
timeout 150 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/s9_c3.json 2> gpurun_out/s9_c3.err
python - <<PY
import json
d=json.load(open("gpurun_out/s9_c3.json")); print(d["value"], d["ms_per_step"], d["roofline"]["achieved"])
for k,v in d["kernels"].items(): print(k, round(v["ms_per_step"],3))
PY
