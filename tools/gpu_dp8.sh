#!/bin/bash
# 8-GPU weak-scaling check, launched the way the driver launches bench.py
mkdir -p gpurun_out
N=${1:-8}
for wl in c2 c4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_${wl}_dp$N.json 2> gpurun_out/bench_${wl}_dp$N.err; echo "bench $wl dp$N rc=$?"
python - $wl $N <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/bench_%s_dp%s.json"%(sys.argv[1],sys.argv[2])))
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value %.0f e2e %.0f ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "loss", d["final_loss"])
except Exception as e:
    print("no json", e)
PY
tail -3 gpurun_out/bench_${wl}_dp$N.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --impl reference > gpurun_out/bench_ref_dp$N.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_dp$N.json
