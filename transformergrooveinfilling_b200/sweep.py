"""Sweep packing: train several hyper-parameter variants of one wandb sweep concurrently on ONE GPU.

The reference is used through ``wandb agent`` random sweeps (``configs/InfillingClosedHH_sweep.yaml:1-37``,
``configs/InfillingKicksAndSnares_sweep_2.yaml``): every agent run is one ``train.py`` process with one small model and a
batch of 16..512 grooves.  At those batch sizes a step of the fused path is a chain of ~40 short kernels that occupies a
handful of the 148 SMs (a tile is 4 sequences: batch 16 = 4 CTAs), so one run leaves the GPU > 90 % idle.  ``SweepPacker``
gives every member of the sweep its own model, fused optimizer, device-resident loader and CUDA stream and drives each
member from its own host thread: the library calls release the GIL (ctypes), so the members' kernel chains are enqueued
concurrently and the hardware scheduler runs them side by side.  No member sees another member's state: results are those
of running each configuration alone (checked in ``tests/test_gpu_sweep.py``).

``sample_sweep`` draws configurations from a wandb sweep specification the way the agent's ``method: random`` does
(``values`` / ``value`` / ``distribution: uniform | int_uniform``), and ``params_from_config`` builds the ``params`` dict of
``train.py:114-143`` from one drawn configuration, so ``initialize_model`` is used unchanged.
"""
from __future__ import annotations

import random
import threading
from typing import Callable, Dict, List, Optional, Sequence

import torch

from .pipeline import DeviceResidentLoader
from .training import initialize_model


def sample_sweep(sweep: Dict, n: int, seed: int = 0) -> List[Dict]:
    """``n`` random draws from a wandb sweep spec (the ``parameters`` block of configs/*_sweep*.yaml).  Draws whose
    ``d_model`` is not divisible by ``n_heads`` are rejected, as ``torch.nn.MultiheadAttention`` rejects them in the
    reference (the agent would record a crashed run)."""
    rng = random.Random(seed)
    spec = sweep.get("parameters", sweep)
    out: List[Dict] = []
    guard = 0
    while len(out) < n:
        guard += 1
        if guard > 1000 * max(n, 1):
            raise ValueError("sweep specification admits no valid configuration")
        cfg = {}
        for name, p in spec.items():
            if "value" in p:
                cfg[name] = p["value"]
            elif "values" in p:
                cfg[name] = rng.choice(list(p["values"]))
            elif p.get("distribution") == "uniform":
                cfg[name] = rng.uniform(float(p["min"]), float(p["max"]))
            elif p.get("distribution") == "int_uniform":
                cfg[name] = rng.randint(int(p["min"]), int(p["max"]))
            else:
                raise ValueError(f"sweep parameter '{name}': unsupported specification {p}")
        if "d_model" in cfg and "n_heads" in cfg and cfg["d_model"] % cfg["n_heads"] != 0:
            continue
        out.append(cfg)
    return out


def params_from_config(cfg: Dict, device="cuda", precision: Optional[str] = None) -> Dict:
    """The ``params`` dict train.py:114-143 builds from ``wandb.config`` (same keys, same derivations)."""
    encoder_only = cfg.get("encoder_only", 1)
    layers = cfg["num_encoder_decoder_layers"]
    model = {
        "experiment": cfg.get("experiment", "InfillingClosedHH"),
        "encoder_only": encoder_only,
        "optimizer": cfg.get("optimizer_algorithm", "sgd"),
        "d_model": cfg["d_model"],
        "n_heads": cfg["n_heads"],
        "dim_feedforward": cfg["dim_feedforward"],
        "dropout": cfg["dropout"],
        "num_encoder_layers": layers,
        "num_decoder_layers": 0 if encoder_only else layers,
        "max_len": 32,
        "embedding_size_src": 16 if cfg.get("experiment", "InfillingClosedHH") != "InfillingClosedHH_Symbolic" else 27,
        "embedding_size_tgt": 27,
        "device": device,
    }
    if precision is not None:
        model["precision"] = precision
    return {"model": model,
            "training": {"learning_rate": cfg["learning_rate"], "batch_size": cfg["batch_size"],
                         "hit_loss_penalty": cfg["hit_loss_penalty"]},
            "load_model": cfg.get("load_model")}


class SweepMember:
    """One configuration of the sweep: model + optimizer + loader + stream + the metrics it has produced."""

    def __init__(self, params: Dict, inputs, outputs, device, seed: int):
        self.params = params
        self.model, self.optimizer, self.epoch = initialize_model(params)
        self.model.set_seed(seed).train()
        self.penalty = float(params["training"]["hit_loss_penalty"])
        self.encoder_only = bool(params["model"]["encoder_only"])
        self.loader = DeviceResidentLoader(inputs, outputs, params["training"]["batch_size"], device, shuffle=True, seed=seed)
        self.stream = torch.cuda.Stream(device=device)
        self.metrics: List[torch.Tensor] = []       # one [6] device tensor per step: loss, acc, ppl, bce, mse_v, mse_o
        self.steps = 0
        self.sequences = 0

    def run_steps(self, n_steps: int) -> None:
        """Enqueue ``n_steps`` optimizer steps on this member's stream (wrapping around epochs of its loader)."""
        if not self.encoder_only:
            raise NotImplementedError("SweepPacker drives encoder-only members (every shipped sweep sets encoder_only: 1)")
        with torch.cuda.stream(self.stream):
            done = 0
            while done < n_steps:
                for x, y, _idx in self.loader:
                    m, _ = self.model.train_step(x, y, self.penalty)
                    self.optimizer.step()
                    self.metrics.append(m)
                    self.sequences += x.shape[0]
                    done += 1
                    if done == n_steps:
                        break
                else:
                    self.epoch += 1
            self.steps += n_steps

    def _fused_run(self, order: torch.Tensor, start: int, bsz: int, k: int) -> None:
        """``k`` consecutive batches of ``bsz`` rows starting at row ``start`` of the permutation, in one ``gt_train_steps`` call
        on the current stream."""
        import ctypes as C
        from . import _lib
        from .training import FusedAdam
        opt, model, ld = self.optimizer, self.model, self.loader
        lib = _lib.load()
        adam = isinstance(opt, FusedAdam)
        dev = ld.device
        key = (bsz, model.precision)
        if model._train_ws is None or model._train_ws[0] != key:
            model._train_ws = (key, model._workspace(bsz, 1, dev), torch.empty(bsz, 32, model.embedding_size_tgt, dtype=torch.float32, device=dev))
        _, ws, hvo = model._train_ws
        xbuf = torch.empty((bsz,) + tuple(ld.x.shape[1:]), device=dev)
        ybuf = torch.empty((bsz,) + tuple(ld.y.shape[1:]), device=dev)
        met = torch.empty(k, 6, dtype=torch.float32, device=dev)
        g = model.flat_grad()
        cfg = model._cfg()
        if adam:
            b1, b2 = opt.param_groups[0]["betas"]
            assert (b1, b2, opt.param_groups[0]["eps"]) == (0.9, 0.999, 1e-8), "gt_train_steps applies torch's Adam defaults"
        _lib.check(lib.gt_train_steps(
            C.byref(cfg), _lib.ptr(model._flat), _lib.ptr(model._pe_flat()), _lib.ptr(ld.x), _lib.ptr(ld.y), _lib.ptr(order),
            start, bsz, k, self.penalty, _lib.ptr(g), _lib.ptr(met), _lib.ptr(hvo), _lib.ptr(xbuf), _lib.ptr(ybuf), _lib.ptr(ws),
            ws.numel(), 1 if adam else 0, opt._lr(), _lib.ptr(opt._m) if adam else None, _lib.ptr(opt._v) if adam else None,
            opt._t if adam else 0, model._seed, model._step, _lib.stream_ptr(dev)), "gt_train_steps")
        model._step += k
        if adam:
            opt._t += k
        self.metrics.extend(met.unbind(0))
        self.sequences += bsz * k
        self._keep = (order, xbuf, ybuf)                      # alive until the stream has consumed them

    def run_steps_fused(self, n_steps: int) -> None:
        """Same steps as ``run_steps`` (same permutations, batches, dropout counters, optimizer updates), but every run of
        equal-sized batches of an epoch is ONE library call (``gt_train_steps``: device-side row gather + train step +
        optimizer per step), so the host thread holds the GIL only between epochs — what lets more than a handful of members
        make progress at once."""
        from .training import FusedAdam, FusedSGD
        if not self.encoder_only:
            raise NotImplementedError("SweepPacker drives encoder-only members (every shipped sweep sets encoder_only: 1)")
        ld = self.loader
        if not isinstance(self.optimizer, (FusedSGD, FusedAdam)):
            return self.run_steps(n_steps)
        dev = ld.device
        with torch.cuda.stream(self.stream):
            done = 0
            S, B = ld.x.shape[0], ld.batch_size
            while done < n_steps:
                order = torch.randperm(S, device=dev, generator=ld.gen) if ld.shuffle else torch.arange(S, device=dev)
                runs = [(0, B, S // B)]                                    # (first row, batch, number of batches)
                if S % B and not ld.drop_last:
                    runs.append((S - S % B, S % B, 1))
                for start, bsz, count in runs:
                    k = min(count, n_steps - done)
                    if k > 0:
                        self._fused_run(order, start, bsz, k)
                        done += k
                if done < n_steps:
                    self.epoch += 1
            self.steps += n_steps

    # ---- CUDA-graph replay: one graph launch per optimizer step --------------------------------------------------------
    def graph_capable(self) -> bool:
        """True when this member's step can be replayed as a CUDA graph (``gt_graph_train_create``): a fused optimizer on any
        path (fp32, fused d_model = 32 / 256, per-op tcgen05); sweep members are encoder-only like every shipped yaml."""
        from . import _lib
        from .training import FusedAdam, FusedSGD
        import ctypes as C
        m = self.model
        if not self.encoder_only or not isinstance(self.optimizer, (FusedSGD, FusedAdam)):
            return False
        cfg = m._cfg()
        if getattr(m, "num_decoder_layers", 0) != 0:
            return False
        kind = _lib.load().gt_path_kind(C.byref(cfg))
        if kind == _lib.PATH_FUSED_D32:
            return m.embedding_size_src in (16, 27) and m.embedding_size_tgt == 27
        return kind in (_lib.PATH_FP32_SIMT, _lib.PATH_GEMM_TC, _lib.PATH_GEMM_TC_SPLIT, _lib.PATH_FUSED_D256)

    def _graph_for(self, bsz: int):
        """The captured step for full batches of ``bsz`` rows (built once; rebuilt if lr / precision change)."""
        import ctypes as C
        from . import _lib
        from .training import FusedAdam
        opt, model, ld = self.optimizer, self.model, self.loader
        adam = isinstance(opt, FusedAdam)
        key = (bsz, model.precision, opt._lr(), self.penalty)
        g = getattr(self, "_graph", None)
        if g is not None and g["key"] == key:
            return g
        if g is not None:
            self.stream.synchronize()
            _lib.check(_lib.load().gt_graph_destroy(g["handle"]), "gt_graph_destroy")
            self._graph = None
        dev = ld.device
        if adam:
            b1, b2 = opt.param_groups[0]["betas"]
            assert (b1, b2, opt.param_groups[0]["eps"]) == (0.9, 0.999, 1e-8), "the captured step applies torch's Adam defaults"
        ring_slots = 1024
        g = dict(key=key, ring_slots=ring_slots,
                 ws=model._workspace(bsz, 1, dev), hvo=torch.empty(bsz, 32, model.embedding_size_tgt, dtype=torch.float32, device=dev),
                 xbuf=torch.zeros((bsz,) + tuple(ld.x.shape[1:]), device=dev), ybuf=torch.zeros((bsz,) + tuple(ld.y.shape[1:]), device=dev),
                 met6=torch.zeros(6, dtype=torch.float32, device=dev), ring=torch.zeros(ring_slots, 6, dtype=torch.float32, device=dev),
                 counters=torch.zeros(4, dtype=torch.int64, device=dev), perm=torch.zeros(ld.x.shape[0], dtype=torch.int64, device=dev))
        handle = C.c_void_p()
        cfg = model._cfg()
        self.stream.synchronize()
        _lib.check(_lib.load().gt_graph_train_create(
            C.byref(cfg), _lib.ptr(model._flat), _lib.ptr(model._pe_flat()), _lib.ptr(g["xbuf"]), _lib.ptr(g["ybuf"]), bsz, self.penalty,
            _lib.ptr(model.flat_grad()), _lib.ptr(g["met6"]), _lib.ptr(g["hvo"]), _lib.ptr(g["ws"]), g["ws"].numel(), 1 if adam else 0,
            opt._lr(), _lib.ptr(opt._m) if adam else None, _lib.ptr(opt._v) if adam else None, model._seed, _lib.ptr(g["counters"]),
            _lib.ptr(ld.x), _lib.ptr(ld.y), _lib.ptr(g["perm"]), _lib.ptr(g["ring"]), ring_slots, _lib.stream_ptr(dev),
            C.byref(handle)), "gt_graph_train_create")
        g["handle"] = handle
        self._graph = g
        return g

    def run_steps_graph(self, n_steps: int) -> None:
        """Same steps as ``run_steps`` / ``run_steps_fused`` (same permutations, batches, dropout counters, optimizer updates),
        but every full batch is ONE ``cudaGraphLaunch`` (row gather + train step + optimizer + bookkeeping captured once by
        ``gt_graph_train_create``); the step counters, the row offset into the permutation and the metrics slot live in device
        memory and the graph advances them itself.  The ragged last batch of an epoch goes through ``gt_train_steps``.
        Members whose path cannot be captured (``graph_capable``) fall back to ``run_steps_fused``."""
        import ctypes as C
        from . import _lib
        from .training import FusedAdam
        if not self.graph_capable():
            return self.run_steps_fused(n_steps)
        opt, model, ld = self.optimizer, self.model, self.loader
        lib = _lib.load()
        adam = isinstance(opt, FusedAdam)
        dev = ld.device
        with torch.cuda.stream(self.stream):
            done = 0
            S, B = ld.x.shape[0], ld.batch_size
            while done < n_steps:
                order = torch.randperm(S, device=dev, generator=ld.gen) if ld.shuffle else torch.arange(S, device=dev)
                self._keep_order = order
                full = min(S // B, n_steps - done)
                if full > 0:
                    g = self._graph_for(B)
                    g["perm"].copy_(order)
                    sp = _lib.stream_ptr(dev)
                    at = 0
                    while at < full:
                        k = min(full - at, g["ring_slots"])
                        g["counters"].copy_(torch.tensor([model._step, opt._t if adam else 0, at * B, 0], dtype=torch.int64))
                        _lib.check(lib.gt_graph_launch(g["handle"], k, sp), "gt_graph_launch")
                        self.metrics.extend(g["ring"][:k].clone().unbind(0))
                        model._step += k
                        if adam:
                            opt._t += k
                        at += k
                    self.sequences += B * full
                    done += full
                if done < n_steps and S % B and not ld.drop_last:
                    self._fused_run(order, S - S % B, S % B, 1)
                    done += 1
                if done < n_steps:
                    self.epoch += 1
            self.steps += n_steps

    def close(self) -> None:
        """Releases the captured step (``gt_graph_destroy``) once the stream has drained; the member stays usable (the graph is
        rebuilt on the next ``run_steps_graph``)."""
        g = getattr(self, "_graph", None)
        if g is not None:
            from . import _lib
            self._graph = None
            self.stream.synchronize()
            _lib.load().gt_graph_destroy(g["handle"])

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the driver reclaims the graph with the context
            pass

    def history(self) -> torch.Tensor:
        """[steps, 6] host tensor of the per-step metrics (synchronises this member's stream)."""
        self.stream.synchronize()
        return torch.stack(self.metrics).cpu() if self.metrics else torch.empty(0, 6)


class SweepPacker:
    """Runs the members of a sweep concurrently on one device.

    ``configs``: drawn sweep configurations (``sample_sweep``) or ready ``params`` dicts; ``inputs`` / ``outputs``: the
    processed dataset (``[S, 32, E_src]`` / ``[S, 32, 27]``), copied to the device once and shared read-only by all members.
    ``concurrent=False`` runs the members one after the other on the same streams (the baseline the packing is measured
    against: what separate agent runs sharing the GPU in turn would get)."""

    def __init__(self, configs: Sequence[Dict], inputs, outputs, device="cuda", precision: Optional[str] = None, seed: int = 0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SweepPacker runs on a CUDA device — the groove_b200 path has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        x = torch.as_tensor(inputs, dtype=torch.float32).to(self.device).contiguous()
        y = torch.as_tensor(outputs, dtype=torch.float32).to(self.device).contiguous()
        self.members: List[SweepMember] = []
        for i, cfg in enumerate(configs):
            params = cfg if "model" in cfg and "training" in cfg else params_from_config(cfg, str(self.device), precision)
            self.members.append(SweepMember(params, x, y, self.device, seed + i))
        torch.cuda.synchronize(self.device)          # dataset copy and parameter initialisation ran on the default stream

    def run(self, n_steps: int, concurrent: bool = True, on_error: Optional[Callable] = None, fused: bool = True,
            graph: bool = False) -> None:
        """Every member takes ``n_steps`` optimizer steps.  Returns once all steps are ENQUEUED; ``synchronize()`` or
        ``history()`` waits for the device.  ``fused=True`` drives each member through ``gt_train_steps`` (one library call per
        run of equal batches of an epoch), ``fused=False`` through the per-step Python loop, ``graph=True`` replays each
        member's step as a CUDA graph (one ``cudaGraphLaunch`` per step; members off the fused d_model = 32 path use the fused
        call); all three take identical steps."""
        step_fn = ((lambda m: m.run_steps_graph(n_steps)) if graph else
                   (lambda m: m.run_steps_fused(n_steps)) if fused else (lambda m: m.run_steps(n_steps)))
        if not concurrent:
            for m in self.members:
                step_fn(m)
                m.stream.synchronize()           # one member at a time on the device, not just on the host
            return
        errors: List[BaseException] = []

        def work(m: SweepMember):
            try:
                torch.cuda.set_device(self.device)
                step_fn(m)
            except BaseException as e:      # surfaced on the caller's thread below
                errors.append(e)

        threads = [threading.Thread(target=work, args=(m,), daemon=True) for m in self.members]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            if on_error is not None:
                on_error(errors)
            else:
                raise errors[0]

    def synchronize(self) -> None:
        for m in self.members:
            m.stream.synchronize()

    def history(self) -> List[torch.Tensor]:
        return [m.history() for m in self.members]

    def best(self) -> int:
        """Index of the member with the lowest last loss (the sweep's ``metric: {goal: minimize, name: loss}``)."""
        last = [float(h[-1, 0]) if h.numel() else float("inf") for h in self.history()]
        return min(range(len(last)), key=last.__getitem__)
