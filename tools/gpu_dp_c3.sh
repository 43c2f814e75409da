#!/bin/bash
# C3 (InfillingKicksAndSnares_training.yaml, BASELINE.json configs[2]) data-parallel weak scaling, launched the way the driver launches bench.py
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_dp$N.json 2> gpurun_out/bench_c3_dp$N.err; echo "bench c3 dp$N rc=$?"
python - $N <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/bench_c3_dp%s.json"%sys.argv[1]))
    print("c3 n_gpus", d["n_gpus"], "value %.0f e2e %.0f ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "loss", d["final_loss"])
except Exception as e:
    print("no json", e)
PY
tail -3 gpurun_out/bench_c3_dp$N.err
