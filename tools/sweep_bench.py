"""Sweep packing measurement: aggregate training sequences/s of K sweep members on ONE B200, packed
(one stream + host thread per member) against the same members run one after the other, at the reference's batch sizes.

    python tools/sweep_bench.py [--members 1,4,8,16] [--steps 60] [--batch 32] [--spec closedhh|c2] [--drive loop|fused|graph]

--spec c2      : every member is InfillingClosedHH_training.yaml (C2) at --batch
--spec closedhh: members drawn from the parameter ranges of configs/InfillingClosedHH_sweep.yaml (restated below)
Prints one JSON line per K."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, emit, synth_batch  # noqa: E402  (bench.py points fd 1 at stderr; emit() writes to the real stdout)
from transformergrooveinfilling_b200 import SweepPacker, sample_sweep  # noqa: E402

# parameter ranges of configs/InfillingClosedHH_sweep.yaml:5-37 (the reference's random sweep)
CLOSEDHH_SWEEP = {"parameters": {
    "batch_size": {"values": [16, 32, 64, 128, 256, 512]}, "d_model": {"values": [16, 32, 64, 128, 256, 512]},
    "dim_feedforward": {"values": [16, 32, 64, 128, 256, 512]}, "dropout": {"distribution": "uniform", "min": 0.1, "max": 0.3},
    "optimizer_algorithm": {"value": "sgd"}, "learning_rate": {"distribution": "uniform", "min": 0, "max": 0.1},
    "n_heads": {"values": [1, 2, 4, 8, 16]}, "num_encoder_decoder_layers": {"distribution": "int_uniform", "min": 6, "max": 12},
    "epochs": {"value": 100}, "encoder_only": {"value": 1}, "experiment": {"value": "InfillingClosedHH"},
    "hit_loss_penalty": {"distribution": "uniform", "min": 0, "max": 1}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", default="1,4,8,16")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--spec", default="c2", choices=["c2", "closedhh"])
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--drive", default="fused", choices=["loop", "fused", "graph"],
                    help="per-step Python loop / gt_train_steps per run of batches / one CUDA-graph launch per step")
    args = ap.parse_args()
    w = WORKLOADS["c2"]
    x, y = synth_batch(w, 4096, 1234)
    for k in [int(v) for v in args.members.split(",")]:
        if args.spec == "c2":
            cfgs = [dict(batch_size=args.batch, d_model=w["d"], dim_feedforward=w["F"], dropout=w["p"], optimizer_algorithm="sgd",
                         learning_rate=w["lr"], n_heads=w["H"], num_encoder_decoder_layers=w["L"], encoder_only=1,
                         experiment="InfillingClosedHH", hit_loss_penalty=w["pen"]) for _ in range(k)]
        else:
            cfgs = sample_sweep(CLOSEDHH_SWEEP, k, seed=11)
        res = {}
        for mode in ("sequential", "packed"):
            torch.manual_seed(0)
            pk = SweepPacker(cfgs, x, y, "cuda", precision=args.precision, seed=3)
            drive = dict(fused=args.drive != "loop", graph=args.drive == "graph")
            pk.run(5, concurrent=(mode == "packed"), **drive)          # warm-up: workspaces, lazy module loads, graph capture
            pk.synchronize()
            seq0 = sum(m.sequences for m in pk.members)
            t0 = time.perf_counter()
            pk.run(args.steps, concurrent=(mode == "packed"), **drive)
            pk.synchronize()
            dt = time.perf_counter() - t0
            res[mode] = (sum(m.sequences for m in pk.members) - seq0) / dt
            res[mode + "_ms_per_member_step"] = dt / args.steps * 1e3 / (k if mode == "sequential" else 1)
            res["final_losses"] = [round(float(h[-1, 0]), 4) for h in pk.history()]
            del pk
            torch.cuda.empty_cache()
        emit(({"metric": "sweep_train_seq_per_s", "members": k, "spec": args.spec, "batch": args.batch if args.spec == "c2" else "sampled",
                          "steps_per_member": args.steps, "precision": args.precision, "drive": args.drive, "sequential": res["sequential"], "packed": res["packed"],
                          "speedup": res["packed"] / res["sequential"], "ms_per_step_sequential": res["sequential_ms_per_member_step"],
                          "ms_per_round_packed": res["packed_ms_per_member_step"], "final_losses": res["final_losses"],
                          "host_threads": k, "host_cpus": os.cpu_count()}))


if __name__ == "__main__":
    main()
