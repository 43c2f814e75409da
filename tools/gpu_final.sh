#!/bin/bash
# end-of-session record: whole GPU suite, smoke, default bench (C2) with CPU baseline, reference arm, every other workload
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo "bench c2 rc=$?"; tail -2 gpurun_out/final_c2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_ref.json 2>/dev/null; echo "ref rc=$?"
for wl in c1 c3 c4 c5; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/final_$wl.json 2> gpurun_out/final_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/final_$wl.err
done
for wl in c2 c4 c5; do
  timeout 300 python bench.py --workload $wl --mode infer --no-cpu-baseline > gpurun_out/final_infer_$wl.json 2>/dev/null; echo "infer $wl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final_*.json")):
    try:
        d=json.load(open(f))
        r=d.get("roofline") or {}
        print("%-34s %-16s value %10.0f e2e %10.0f ms %8.3f batch %6d path %s roof %.1f/%.3f traffic %s"%(f.split('/')[-1], d["metric"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["per_gpu_batch"], d.get("path"), r.get("achieved",0), r.get("frac",0), r.get("traffic")))
    except Exception as e:
        print(f, "ERR", e)
PY
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python tools/sweep_bench.py --members 1,8,32 --steps 120 --drive graph > gpurun_out/final_sweep_graph.jsonl 2> gpurun_out/final_sweep_graph.err; echo "sweep graph rc=$?"; cut -c1-330 gpurun_out/final_sweep_graph.jsonl
