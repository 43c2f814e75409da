#!/usr/bin/env python
"""bench.py — headline benchmark of the groove_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A "step" is one full training step (forward with dropout, calculate_loss, backward, optimizer update)
over one synthetic batch of 2-bar grooves (32 steps x 27 hvo channels) of the workload's shape.
Rank 0 prints ONE JSON line (see DESIGN.md §Measurement for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the ONE JSON line (NCCL prints its version banner there)

# stdout carries exactly ONE JSON line: everything else a library prints there (NCCL banner, torchrun notices)
# is sent to stderr by pointing fd 1 at fd 2 for the whole run; emit() writes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = os.fdopen(os.dup(2), "w")


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import torch  # noqa: E402

# hyper-parameters verbatim from the reference yamls (C1..C5); `batch` is the per-GPU synthetic batch used for
# throughput, the sizes of SURVEY.md §8(d): 262 144 (C1), 65 536 (C2 / C5), 16 384 (C3 / C4)
WORKLOADS = {
    "c1": dict(yaml="InfillingClosedHH_testing_training.yaml", d=32, H=4, F=16, L=6, Ld=0, E=16, p=0.18, lr=0.094, pen=0.47, batch=262144),
    "c2": dict(yaml="InfillingClosedHH_training.yaml", d=32, H=16, F=512, L=6, Ld=0, E=16, p=0.24, lr=0.07, pen=0.38, batch=65536),
    "c3": dict(yaml="InfillingKicksAndSnares_training.yaml", d=256, H=2, F=512, L=6, Ld=0, E=16, p=0.30, lr=0.089, pen=0.73, batch=16384),
    "c4": dict(yaml="InfillingRandom_test_large.yaml", d=256, H=16, F=64, L=11, Ld=0, E=16, p=0.15, lr=0.04, pen=1.0, batch=16384),
    "c5": dict(yaml="InfillingClosedHH_Symbolic_training.yaml(encoder_only=0)", d=32, H=16, F=512, L=6, Ld=6, E=27, p=0.24, lr=0.07, pen=0.38, batch=65536),
}


def train_flops_per_seq(w):
    """Algorithmic FLOPs per sequence: 2 FLOP/MAC, backward = 2x forward, recompute
    not counted, element-wise / softmax / LN not counted."""
    d, F, L, Ld, E = w["d"], w["F"], w["L"], w["Ld"], w["E"]
    mac = 32 * E * d + L * (32 * (4 * d * d + 2 * d * F) + 2 * 32 * 32 * d) + 32 * d * 27
    if Ld:
        mac += 32 * 27 * d + Ld * (32 * (8 * d * d + 2 * d * F) + 4 * 32 * 32 * d)
    return 3 * 2 * mac


def synth_batch(w, n, seed):
    """SURVEY.md §8d synthetic inputs, generated on the CPU with a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    hits = (torch.rand(n, 32, 9, generator=g) < 0.15).float()
    y = torch.cat((hits, torch.rand(n, 32, 9, generator=g) * hits, (torch.rand(n, 32, 9, generator=g) - 0.5) * hits), 2)
    if w["E"] == 27:
        h2 = (torch.rand(n, 32, 9, generator=g) < 0.15).float()
        x = torch.cat((h2, torch.rand(n, 32, 9, generator=g) * h2, (torch.rand(n, 32, 9, generator=g) - 0.5) * h2), 2)
    else:
        m = (torch.rand(n, 32, 8, generator=g) < 0.5).float()
        x = torch.cat((torch.rand(n, 32, 8, generator=g) * m, (torch.rand(n, 32, 8, generator=g) - 0.5) * m), 2)
    return x.contiguous(), y.contiguous()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 100 ms through NVML while running."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def reference_stepper(w, batch, dropout, optimizer, device):
    """One training step of the reference, as a closure: the body of BGT/models/train.py:118-141 without wandb
    (zero_grad -> forward -> calculate_loss -> backward -> opt.step()).

    kind "reference": the UNMODIFIED reference modules from oracle/_ref (oracle/build_ref.py copies them there, git-ignored,
    at build() time): GrooveTransformerEncoder / GrooveTransformer + calculate_loss + torch.optim.SGD / Adam exactly as
    initialize_model builds them (BGT/models/train.py:43-66), nn.Dropout at the given p.
    kind "port": oracle/groove_oracle.py (fused ATen ops, torch-native dropout) — only when oracle/_ref is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ref, _why = ref_loader.load_reference()
    x, y = synth_batch(w, batch, 1234)
    x, y = x.to(device), y.to(device)
    if ref is not None:
        import importlib
        tr = importlib.import_module(ref_loader.NAME + ".models.transformer")
        torch.manual_seed(0)
        if w["Ld"]:
            model = tr.GrooveTransformer(w["d"], w["E"], 27, w["H"], w["F"], dropout, w["L"], w["Ld"], 32, device)
        else:
            model = tr.GrooveTransformerEncoder(w["d"], w["E"], 27, w["H"], w["F"], dropout, w["L"], 32, device)
        model.to(device)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3) if optimizer == "adam" else torch.optim.SGD(model.parameters(), lr=w["lr"])
        bce, mse = torch.nn.BCEWithLogitsLoss(reduction="none"), torch.nn.MSELoss(reduction="none")
        y_s = torch.cat((torch.zeros_like(y[:, :1]), y[:, :-1]), dim=1) if w["Ld"] else None       # train.py:130-131

        def step():
            opt.zero_grad()
            pred = model(x, y_s) if w["Ld"] else model(x)
            loss = ref.calculate_loss(pred, y, bce, mse, w["pen"])[0]
            loss.backward()
            opt.step()
            return loss
        return step, "reference"
    import groove_oracle as G
    G.FAST_BASELINE = True
    cfg = G.GrooveCfg(w["d"], w["H"], w["F"], w["L"], w["Ld"], w["E"], 27, dropout)
    state = {"P": {k: v.to(device) for k, v in G.det_params(cfg).items()}, "t": 0}
    state["m"] = {k: torch.zeros_like(v) for k, v in state["P"].items()}
    state["v"] = {k: torch.zeros_like(v) for k, v in state["P"].items()}

    def step():
        P = state["P"]
        _, grads, _ = G.train_step_oracle(P, cfg, x, y, w["pen"], G.DropCtx(dropout, train=True, native=True))
        if optimizer == "adam":
            state["t"] += 1
            for k in P:
                P[k], state["m"][k], state["v"][k] = G.adam_step(P[k], grads[k], state["m"][k], state["v"][k], state["t"], 1e-3)
        else:
            state["P"] = {k: G.sgd_step(v, grads[k], w["lr"]) for k, v in P.items()}
    return step, "port"


def cpu_baseline(w, steps, warmup, batch, dropout, optimizer):
    """The reference's own CPU implementation of the step, timed on this box's host cores (all of them)."""
    torch.set_num_threads(os.cpu_count())
    step, kind = reference_stepper(w, batch, dropout, optimizer, "cpu")
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tot = sum(times)
    return batch * len(times) / tot, tot / len(times), kind


def cuda_eager_baseline(w, steps, warmup, batch, dropout, optimizer, dev):
    """SURVEY.md §2 / §8(d) "second baseline": the same reference code with device="cuda" (train.py:133) — PyTorch eager on this
    B200, fp32 (TF32 off, torch defaults), timed with CUDA events.  calculate_loss's five .item() calls sync every step, as
    they do in the reference."""
    step, kind = reference_stepper(w, batch, dropout, optimizer, dev)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return batch * steps / (ms / 1e3), ms / steps, kind


def run_reference(args, w):
    """--impl reference: the reference's own CPU path on this box's host cores, same workload (model, dropout, optimizer);
    each step is a bounded sample (batch 512 of the per-GPU batch the GPU arm runs) so that K + W steps end within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 512 if args.steps + args.warmup <= 30 else 256
    v, sec, kind = cpu_baseline(w, args.steps, args.warmup, batch, w["p"], args.optimizer)
    what = "unmodified reference from oracle/_ref (BGT/models/train.py:118-141 body)" if kind == "reference" else "oracle port (oracle/_ref absent)"
    line = {
        "impl": "reference", "metric": "train_seq_per_s", "value": v, "unit": "seq/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, args, batch, 1, "f32-cpu"),
        "cpu_baseline": {"value": v, "unit": "seq/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{args.steps} steps of batch {batch} ({w['yaml']}, dropout {w['p']}, {args.optimizer}); {what}",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": v, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(w, args, per_gpu_batch, world, precision):
    return {"workload": f"{w['yaml']} full train step (fwd+loss+bwd+{args.optimizer})", "d_model": w["d"], "nhead": w["H"],
            "dim_feedforward": w["F"], "num_encoder_layers": w["L"], "num_decoder_layers": w["Ld"], "dropout": w["p"],
            "hit_loss_penalty": w["pen"], "optimizer": args.optimizer, "per_gpu_batch": per_gpu_batch,
            "global_batch": per_gpu_batch * world, "seq_len": 32, "channels": 27, "precision": precision,
            "parallelism": f"dp{world}",
            "l2_policy": "per-step working set (inputs + saved activations) is far larger than the 126 MB L2; no flush needed"}


def run_infer(args, w, model, lib, x, xh, n, world, rank, dev, barrier, timed, precision):
    """predict() throughput (BGT/models/transformer.py:117-125 / :48-83): sequences/s through model.predict."""
    import ctypes as C
    out_h = torch.empty(n, 32, 27, dtype=torch.float32).pin_memory()
    model.eval()

    def step_resident():
        with torch.no_grad():
            model._predict_hvo(x, 0.5)

    # e2e: the public host-array path (pipeline.HostPredictor: what the evaluator does with predict(), evaluator.py:171-175).
    # Every step copies the pinned host inputs to the device and the [N, 32, 27] result back to pinned host memory, both
    # inside the timed region, chunked so the copies of neighbouring chunks overlap gt_predict.
    from transformergrooveinfilling_b200 import HostPredictor
    hp = HostPredictor(model, chunk=args.infer_chunk or None)

    def step_e2e():
        hp.predict(xh, out=out_h)

    for _ in range(args.warmup):
        step_resident()
    from transformergrooveinfilling_b200 import _lib
    path_kind = lib.gt_path_kind(C.byref(model._cfg()))
    fused = path_kind in (_lib.PATH_FUSED_D32, _lib.PATH_FUSED_D256)
    lib.gt_profile_enable({0: 1, 1: 17, 2: 17, 3: 22, 4: 22}[path_kind], 16384)
    l0 = lib.gt_launch_count(-1)
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.finish()
    launches = lib.gt_launch_count(-1) - l0
    tot_ms, cnt = C.c_double(0), C.c_int64(0)
    lib.gt_profile_collect(C.byref(tot_ms), C.byref(cnt))
    lib.gt_profile_enable(0, 0)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fl_step = train_flops_per_seq(w) / 3 * n
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    roof = None
    if cnt.value and fused and not w["Ld"]:
        layer_mac = 32 * (4 * w["d"] ** 2 + 2 * w["d"] * w["F"]) + 2 * 32 * 32 * w["d"]
        avg_ms = tot_ms.value / cnt.value
        ach = 2 * layer_mac * n / (avg_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "tc_layer_fwd (eval)", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": None, "avg_launch_ms": avg_ms, "launches_timed": cnt.value, "kernel_share_of_step": tot_ms.value / ms}
    cfg = workload_config(w, args, n, world, precision)
    cfg["workload"] = f"{w['yaml']} predict() (eval forward + hit threshold" + (", autoregressive decoder)" if w["Ld"] else ")")
    line = {"metric": "infer_seq_per_s", "value": n * world * args.steps / (ms / 1e3), "unit": "seq/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic", "config": cfg,
            "step_tflops": fl_step * world * args.steps / (ms / 1e3) / 1e12, "roofline": roof,
            "e2e": {"value": n * world * args.steps / (ms_e2e / 1e3), "unit": "seq/s", "h2d_bytes_per_step": xh.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": ms_e2e / args.steps, "chunk": hp.chunk,
                    "api": "HostPredictor.predict (H2D / gt_predict / D2H of neighbouring chunks on three streams)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "path": {0: "fp32_simt", 1: "fused_tcgen05_d32", 2: "fused_tcgen05_d256", 3: "per_op_gemm_tc", 4: "per_op_gemm_tc_split_fp32"}[path_kind]}
    if cnt.value and not fused:
        line["dominant_gemm"] = {"class": "gemm_tc" if path_kind in (3, 4) else "gemm_f32", "launches_timed": cnt.value,
                                 "share_of_step": tot_ms.value / ms}
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def quick_infer(args, w, model, x, xh, n, timed):
    """predict() on the headline workload: resident value + e2e through pipeline.HostPredictor (pinned host array in, pinned
    [N, 32, 27] host array out, copies inside the timed region) — BGT/models/transformer.py:117-125, evaluator.py:171-175."""
    from transformergrooveinfilling_b200 import HostPredictor
    steps = max(3, min(args.steps, 10))
    model.eval()
    out_h = torch.empty(n, 32, 27, dtype=torch.float32).pin_memory()
    hp = HostPredictor(model)

    def resident():
        with torch.no_grad():
            model._predict_hvo(x, 0.5)

    def e2e():
        hp.predict(xh, out=out_h)

    for _ in range(3):
        resident()
    ms = timed(resident, steps)
    for _ in range(2):
        e2e()
    ms_e = timed(e2e, steps)
    return {"metric": "infer_seq_per_s", "value": n * steps / (ms / 1e3), "unit": "seq/s", "ms_per_step": ms / steps, "steps": steps,
            "per_gpu_batch": n, "step_tflops": train_flops_per_seq(w) / 3 * n * steps / (ms / 1e3) / 1e12,
            "e2e": {"value": n * steps / (ms_e / 1e3), "unit": "seq/s", "h2d_bytes_per_step": xh.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": ms_e / steps,
                    "api": "HostPredictor.predict (H2D / gt_predict / D2H of neighbouring chunks on three streams)"}}


def quick_train(args, name, dev, lib):
    """Train value (resident inputs) of another BASELINE config on this GPU, in the same record: 3 warm-up + 6 timed steps."""
    import ctypes as C
    from transformergrooveinfilling_b200 import FusedAdam, FusedSGD, GrooveTransformer, GrooveTransformerEncoder, _lib
    w = WORKLOADS[name]
    n = w["batch"]
    torch.manual_seed(0)
    if w["Ld"]:
        m = GrooveTransformer(w["d"], w["E"], 27, w["H"], w["F"], w["p"], w["L"], w["Ld"], 32, dev)
    else:
        m = GrooveTransformerEncoder(w["d"], w["E"], 27, w["H"], w["F"], w["p"], w["L"], 32, dev)
    m.set_precision("bf16").set_seed(1234).train()
    opt = FusedAdam(m, 1e-3) if args.optimizer == "adam" else FusedSGD(m, w["lr"])
    xh, yh = synth_batch(w, n, 1234)
    x, y = xh.to(dev), yh.to(dev)

    def step():
        m.train_step(x, y, w["pen"])
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    kind = lib.gt_path_kind(C.byref(m._cfg()))
    v = n / (ms / 1e3)
    return {"metric": "train_seq_per_s", "workload": w["yaml"], "value": v, "unit": "seq/s", "ms_per_step": ms, "steps": 6,
            "per_gpu_batch": n, "optimizer": args.optimizer, "precision": "bf16", "step_tflops": train_flops_per_seq(w) * v / 1e12,
            "path": {0: "fp32_simt", 1: "fused_tcgen05_d32", 2: "fused_tcgen05_d256", 3: "per_op_gemm_tc"}[kind]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: workload table)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16", "fp32_tc"],
                    help="fp32_tc: the 1e-4 parity mode with every Linear contraction on tcgen05 (three-term bf16 split operands)")
    ap.add_argument("--optimizer", default="adam", choices=["adam", "sgd"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="train step (headline) or predict()")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-this-GPU run of the reference modules")
    ap.add_argument("--no-extras", action="store_true", help="skip the infer / other-workload keys of the N = 1 line")
    ap.add_argument("--infer-chunk", type=int, default=0, help="--mode infer: sequences per chunk of the host-array predict pipeline (e2e); 0 = HostPredictor's default")
    ap.add_argument("--dp-bucket-mb", type=float, default=4.0, help="gradient bucket size of the overlapped all-reduce (N > 1)")
    ap.add_argument("--no-overlap", action="store_true", help="one all-reduce after backward instead of per-bucket overlap")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: gradient exchange + optimizer as one kernel over NVLink peer memory (p2p; auto = when the ranks can "
                         "map each other's buffers) or bucketed NCCL all-reduce + optimizer kernel (nccl)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import ctypes as C
    import torch.distributed as dist
    from transformergrooveinfilling_b200 import FusedAdam, FusedSGD, GrooveTransformer, GrooveTransformerEncoder, _lib
    from transformergrooveinfilling_b200.dp import DataParallelStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the groove_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n = args.batch or w["batch"]

    def make(precision):
        if w["Ld"]:
            m = GrooveTransformer(w["d"], w["E"], 27, w["H"], w["F"], w["p"], w["L"], w["Ld"], 32, dev)
        else:
            m = GrooveTransformerEncoder(w["d"], w["E"], 27, w["H"], w["F"], w["p"], w["L"], 32, dev)
        m.set_precision(precision).set_seed(1234).train()
        return m

    torch.manual_seed(0)
    precision = args.precision
    model = None
    if precision in ("auto", "bf16"):
        try:
            model = make("bf16")
            model._workspace(4, 1, dev)          # raises if the tensor-core path does not cover this shape
            precision = "bf16"
        except RuntimeError:
            if args.precision == "bf16":
                raise
            model = None
    if model is None:
        torch.manual_seed(0)
        precision = "fp32_tc" if args.precision == "fp32_tc" else "fp32"
        model = make(precision)
    if world > 1:                                   # identical replicas
        dist.broadcast(model.flat_parameters().detach(), 0)
    opt = FusedAdam(model, 1e-3) if args.optimizer == "adam" else FusedSGD(model, w["lr"])
    dp = DataParallelStep(model, opt, w["pen"], overlap=False if args.no_overlap else None,
                          bucket_bytes=int(args.dp_bucket_mb * (1 << 20)), exchange=args.exchange)

    xh, yh = synth_batch(w, n, 1234 + rank)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    x, y = xh.to(dev), yh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if args.mode == "infer":
        return run_infer(args, w, model, lib, x, xh, n, world, rank, dev, barrier, timed, precision)

    def step_resident():
        dp.step(x, y)

    # e2e: the public host-buffer path (pipeline.HostBatchPrefetcher -> DataParallelStep.step).  Every step copies its
    # inputs from pinned host memory (on a copy stream, overlapping the previous step's kernels) and reads its six
    # metrics back to the host; the read of step i is consumed on the host while step i+1 is already enqueued.
    from transformergrooveinfilling_b200.pipeline import HostBatchPrefetcher
    feeder = HostBatchPrefetcher(dev, tuple(xh.shape), tuple(yh.shape))
    metrics_host = [torch.empty(6, dtype=torch.float32).pin_memory() for _ in range(2)]
    metrics_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "loss": float("nan")}

    def step_e2e():
        i = e2e_state["i"]
        if i == 0:
            feeder.submit(xh, yh)
        xd, yd = feeder.get()
        feeder.submit(xh, yh)                            # next step's inputs: H2D overlaps this step's kernels
        m = dp.step(xd, yd)
        metrics_host[i & 1].copy_(m, non_blocking=True)  # the step's result (loss + 5 metrics), read back every step
        metrics_ev[i & 1].record()
        if i > 0:
            metrics_ev[(i - 1) & 1].synchronize()
            e2e_state["loss"] = float(metrics_host[(i - 1) & 1][0])
        e2e_state["i"] = i + 1

    def e2e_drain():
        i = e2e_state["i"]
        if i > 0:
            metrics_ev[(i - 1) & 1].synchronize()
            e2e_state["loss"] = float(metrics_host[(i - 1) & 1][0])

    for _ in range(args.warmup):
        step_resident()
    # ---- timed region: kernel-resident throughput, dominant kernel class bracketed by CUDA events ----
    # fp32 SIMT GEMM / fused layer backward kernel / generic tcgen05 GEMM (shapes without fused layer kernels)
    path_kind = lib.gt_path_kind(C.byref(model._cfg()))
    dom = {_lib.PATH_FP32_SIMT: 1, _lib.PATH_FUSED_D32: 18, _lib.PATH_FUSED_D256: 18, _lib.PATH_GEMM_TC: 22, _lib.PATH_GEMM_TC_SPLIT: 22}[path_kind]
    lib.gt_profile_enable(dom, 4096)
    l0 = lib.gt_launch_count(-1)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.finish()
    launches = lib.gt_launch_count(-1) - l0
    tot_ms, cnt = C.c_double(0), C.c_int64(0)
    lib.gt_profile_collect(C.byref(tot_ms), C.byref(cnt))
    dom_launches_per_step = lib.gt_launch_count(dom)
    lib.gt_profile_enable(0, 0)
    # ---- end-to-end: pinned host inputs copied in, metrics copied out, every step ----
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_drain()
    final_loss = e2e_state["loss"]

    value = n * world * args.steps / (ms / 1e3)
    e2e = n * world * args.steps / (ms_e2e / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fl_step = train_flops_per_seq(w) * n
    roof = None
    if cnt.value > 0:
        # the dominant kernel class and the share of the step's algorithmic FLOPs its launches carry
        if path_kind in (_lib.PATH_GEMM_TC, _lib.PATH_GEMM_TC_SPLIT):
            # every Linear contraction of the step (forward, data gradient, weight gradient) except the K = 16 / 27 input
            # layers and the 27-wide head runs in gemm_tc; attention runs in the SIMT attention kernels
            d, F, L, Ld = w["d"], w["F"], w["L"], w["Ld"]
            lin = L * 32 * (4 * d * d + 2 * d * F) + Ld * 32 * (8 * d * d + 2 * d * F)
            fl_launch = 3 * 2 * lin * n * args.steps / max(cnt.value, 1)
            kname = "gemm_tc (generic tcgen05 GEMM: fp32 operands rounded to bf16 while staged, fp32 accumulate; average over all Linear fwd / dgrad / wgrad launches)"
        elif precision == "bf16":
            layer_mac = 32 * (4 * w["d"] ** 2 + 2 * w["d"] * w["F"]) + 2 * 32 * 32 * w["d"]
            if w["d"] == 256:
                # d_model = 256: the layer-backward kernel computes the data gradients (1x the layer's forward FLOPs);
                # the weight gradients run in t256_wgrad_kernel (class 20, listed under "kernels")
                fl_launch = 2 * layer_mac * n
                kname = "t256_layer_bwd (data gradients + attention backward; recomputed q|k|v and probabilities not counted)"
            elif w["Ld"]:
                # encoder-decoder: 6 whole-layer launches + 3 block launches per decoder layer; average over the launches
                dec_mac = 32 * (8 * w["d"] ** 2 + 2 * w["d"] * w["F"]) + 4 * 32 * 32 * w["d"]
                fl_launch = 2 * 2 * (w["L"] * layer_mac + w["Ld"] * dec_mac) * n * args.steps / max(cnt.value, 1)
                kname = "tc_layer_bwd (encoder layers + decoder self-attention / cross-attention / FFN blocks; average launch)"
            else:
                fl_launch = 2 * 2 * layer_mac * n       # one encoder layer backward = 2x its forward FLOPs
                kname = "tc_layer_bwd (data + weight gradients)"
        else:
            fl_launch = fl_step * args.steps / max(cnt.value, 1)   # every contraction except attention runs in gemm_f32
            kname = "gemm_f32 (fp32 FMA; tensor peak shown for reference only)"
        avg_ms = tot_ms.value / cnt.value
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = fl_launch / (avg_ms / 1e3) / 1e12
        traffic, traffic_note, tr = None, None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
            if tr and precision == "bf16":
                traffic = tr["dram_bytes_per_launch"] * n / tr["per_gpu_batch"]
                traffic_note = (f"ncu dram bytes of one {tr['kernel']} launch at batch {tr['per_gpu_batch']}"
                                + ("" if n == tr["per_gpu_batch"] else f", scaled per sequence to batch {n}"))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", peaks.get("hbm_gb_s", 6550.4))
        roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "hbm_achieved_gbs": (traffic / (avg_ms / 1e3) / 1e9) if traffic else None, "hbm_peak_gbs": hbm_peak,
                "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_note,
                "avg_launch_ms": avg_ms, "launches_timed": cnt.value,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (B200_PROFILING.md)",
                "kernel_share_of_step": tot_ms.value / ms}
        # what actually limits the kernel (ncu `--set full` capture summarised under profiles/; the contract's `bound` stays
        # the roof `achieved` / `peak` are quoted on): neither roof binds the d_model = 32 kernels — CUDA-core issue does
        if tr and precision == "bf16" and "limiter" in tr:
            roof["limiter"] = tr["limiter"]
            for k in ("issue_active", "tensor_pipe_active", "warp_insts_per_seq_layer", "source"):
                if k in tr:
                    roof["ncu_" + k if k != "source" else "ncu_source"] = tr[k]
    # ---- per-kernel-class breakdown: 3 extra steps with every class bracketed by events (outside the timed regions) ----
    kernels = {}
    names = {1: "gemm_f32", 2: "attention_fwd_f32", 3: "attention_bwd_f32", 4: "layernorm", 5: "elementwise", 6: "loss", 7: "optimizer",
             16: "tc_weight_prep", 17: "tc_layer_fwd", 18: "tc_layer_bwd", 19: "fused_tail (final LN + head + loss)", 20: "tc_wgrad",
             21: "fused_stem (input layer + pe)", 22: "gemm_tc"}
    lib.gt_profile_enable(-1, 8192)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step_resident()
    e1.record()
    torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1)
    layer_mac = 32 * (4 * w["d"] ** 2 + 2 * w["d"] * w["F"]) + 2 * 32 * 32 * w["d"]
    lin_mac = 32 * (4 * w["d"] ** 2 + 2 * w["d"] * w["F"])
    fl_cls = {17: 2 * layer_mac * n, 18: (2 if w["d"] == 256 else 4) * layer_mac * n, 20: 2 * lin_mac * n}
    for cls, nm in names.items():
        t, c = C.c_double(0), C.c_int64(0)
        lib.gt_profile_collect_class(cls, C.byref(t), C.byref(c))
        if c.value:
            kernels[nm] = {"launches_per_step": c.value / 3, "ms_per_step": t.value / 3, "share": t.value / ms3}
            if path_kind in (_lib.PATH_FUSED_D32, _lib.PATH_FUSED_D256) and cls in fl_cls and not w["Ld"]:
                # one launch per encoder layer carries fl_cls[cls] (the weight-gradient class also holds the two small
                # input-layer / head launches of edge256.cu: their time is included, their FLOPs are not)
                kernels[nm]["tflops"] = fl_cls[cls] * w["L"] / (t.value / 3 / 1e3) / 1e12
    lib.gt_profile_collect(C.byref(tot_ms), C.byref(cnt))
    lib.gt_profile_enable(0, 0)
    line = {
        "metric": "train_seq_per_s", "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": dict(workload_config(w, args, n, world, precision),
                       **({"exchange": ("one kernel: rank-ordered gradient sum over NVLink peer memory + optimizer (csrc/peer_opt.cu)"
                                        if dp.exchange == "p2p" else "bucketed NCCL all-reduce overlapped with backward + optimizer kernel")}
                          if world > 1 else {})),
        "step_tflops": fl_step * world * args.steps / (ms / 1e3) / 1e12,
        "roofline": roof,
        "e2e": {"value": e2e, "unit": "seq/s", "h2d_bytes_per_step": (xh.numel() + yh.numel()) * 4, "d2h_bytes_per_step": 24,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "final_loss": final_loss, "kernels": kernels,
        "path": {0: "fp32_simt", 1: "fused_tcgen05_d32", 2: "fused_tcgen05_d256", 3: "per_op_gemm_tc", 4: "per_op_gemm_tc_split_fp32"}[path_kind],
    }
    if rank == 0 and world == 1 and not args.no_extras:
        # ---- the other halves of BASELINE.json's metric, in the same driver-run record: predict() throughput on this
        # workload (resident + through HostPredictor with host buffers) and the train value of the other configs ----
        del dp, opt, feeder
        model._train_ws = None
        torch.cuda.empty_cache()
        line["infer"] = quick_infer(args, w, model, x, xh, n, timed)
        del model, x, y
        torch.cuda.empty_cache()
        line["other_workloads"] = {}
        for name in ("c3", "c4"):
            if name != args.workload:
                try:
                    line["other_workloads"][name] = quick_train(args, name, dev, lib)
                except Exception as e:  # noqa: BLE001 — an extra must never cost the headline line
                    line["other_workloads"][name] = {"error": repr(e)[:200]}
                torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        try:
            eb = min(n, 8192)
            v, ms_e, kind = cuda_eager_baseline(w, 5, 3, eb, w["p"], args.optimizer, dev)
            line["cuda_eager_baseline"] = {"value": v, "unit": "seq/s", "ms_per_step": ms_e, "kind": kind, "dtype": "f32",
                                           "sample": f"5 steps of batch {eb}, same workload (dropout {w['p']}, {args.optimizer}); "
                                                     "reference modules with device='cuda' (train.py:133), PyTorch eager on this GPU"}
        except Exception as e:  # noqa: BLE001
            line["cuda_eager_baseline"] = {"error": repr(e)[:200]}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = 512
        v, sec, kind = cpu_baseline(w, 4, 1, cb, w["p"], args.optimizer)
        v0, _, _ = cpu_baseline(w, 6, 1, cb, 0.0, args.optimizer)
        what = "unmodified reference from oracle/_ref (BGT/models/train.py:118-141 body)" if kind == "reference" else "oracle port (oracle/_ref absent)"
        line["cpu_baseline"] = {"value": v, "unit": "seq/s", "cores": torch.get_num_threads(), "kind": kind,
                                "sample": f"4 steps of batch {cb}, same workload (dropout {w['p']}, {args.optimizer}); {what}",
                                "value_dropout0": v0, "host_cpus": os.cpu_count()}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
