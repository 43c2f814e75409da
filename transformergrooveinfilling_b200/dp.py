"""Data-parallel training step: one process per GPU, batch sharded across ranks, full parameter
replica per rank, ONE all-reduce(SUM) of the flat fp32 gradient per step over NCCL (NVLink 5 /
NVSwitch), 1/world folded into the fused optimizer kernel.

The reference has no distributed code; this is the only collective the path needs: sequences are
independent in forward, loss and backward, the only coupling is the batch mean (BGT/models/train.py:
17,21,25) and the shared parameters.  With equal shards mean-of-shard-means == global mean, so
all-reduce(SUM)/world reproduces the single-GPU gradient at the same global batch.  Dropout masks are
keyed by the GLOBAL sequence index (seq0 = rank * n_local), so a DP run draws the same masks as the
single-GPU run of the concatenated batch."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous equal shards; n_global must divide evenly (otherwise mean-of-means != global mean)."""
    if n_global % world != 0:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return rank * per, (rank + 1) * per


def combine_metrics(metrics: torch.Tensor, group=None) -> torch.Tensor:
    """metrics6 = [loss, acc, ppl, bce, mse_v, mse_o] of the local shard -> global values.
    Every entry except the perplexity is a mean over the shard; ppl = exp(global bce)."""
    world = dist.get_world_size(group)
    m = metrics.clone()
    dist.all_reduce(m, op=dist.ReduceOp.SUM, group=group)
    m /= world
    m[2] = torch.exp(m[3])
    return m


def merge_buckets(ranges: Sequence[Tuple[int, int]], min_floats: int) -> List[Tuple[int, int, int]]:
    """``ranges`` = (offset, size) of the library's gradient buckets in completion order — a descending
    partition of the flat vector (gt_grad_buckets).  Consecutive buckets are merged until a group
    holds at least ``min_floats`` floats (an all-reduce below ~0.5 MB is pure launch latency); a
    short tail joins the previous group.  Returns (offset, size, index of the LAST bucket of the
    group) — the group may be reduced once that bucket's event has fired."""
    end = None
    for o, sz in ranges:
        if end is not None and o + sz != end:
            raise ValueError("gradient buckets must form a descending contiguous partition")
        end = o
    groups: List[List[int]] = []
    cur = None
    for i, (o, sz) in enumerate(ranges):
        cur = [o, sz, i] if cur is None else [o, cur[1] + sz, i]
        if cur[1] >= min_floats:
            groups.append(cur)
            cur = None
    if cur is not None:                       # short tail: join the previous group
        if groups:
            groups[-1] = [cur[0], groups[-1][1] + cur[1], cur[2]]
        else:
            groups.append(cur)
    return [tuple(g) for g in groups]


class PeerExchange:
    """Gradient exchange over NVLink peer memory for one process per GPU on ONE node (csrc/peer_opt.cu): every rank owns an
    exchange buffer (two halves, alternating by step) that all peers map through CUDA IPC; ``step`` = publish the flat gradient
    -> barrier (a tiny NCCL all-reduce that also carries the step's six metrics) -> ONE kernel that sums element i over all ranks'
    buffers in rank order and applies the optimizer.  Setup is collective and agrees across ranks: if any rank cannot allocate
    or map a buffer, ``ok`` is False everywhere and the caller keeps the NCCL all-reduce."""

    def __init__(self, n_floats: int, device, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib, self.group = C, _lib.load(), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = int(n_floats)
        self.half = (self.n + 63) // 64 * 64                    # floats per half (16-byte aligned offsets)
        self.device = device
        self.local = C.c_void_p()
        self.peers = [None] * self.world
        self.parity = 0
        self.bar = torch.zeros(8, dtype=torch.float32, device=device)
        handle = C.create_string_buffer(64)
        ok, why = True, ""
        with torch.cuda.device(device):
            if self._lib.gt_peer_alloc(2 * self.half * 4, C.byref(self.local), handle) != 0:
                ok, why = False, self._lib.gt_last_error().decode()
            handles = [None] * self.world
            dist.all_gather_object(handles, (bool(ok), bytes(handle.raw)), group=group)
            ok = all(h[0] for h in handles)
            if ok:
                for r, (_, hb) in enumerate(handles):
                    if r == self.rank:
                        self.peers[r] = self.local.value
                        continue
                    q = C.c_void_p()
                    if self._lib.gt_peer_open(hb, C.byref(q)) != 0:
                        ok, why = False, self._lib.gt_last_error().decode()
                        break
                    self.peers[r] = q.value
            flags = [None] * self.world
            dist.all_gather_object(flags, (bool(ok), why), group=group)
        self.ok = all(f[0] for f in flags)
        self.why = "; ".join(f[1] for f in flags if f[1])
        if self.ok:
            self.bufs = (C.c_void_p * self.world)(*self.peers)
        else:
            self.close()

    def publish(self, grad: torch.Tensor):
        from . import _lib
        _lib.check(self._lib.gt_peer_publish(self.local, self.parity * self.half, _lib.ptr(grad), self.n,
                                             _lib.stream_ptr(self.device)), "gt_peer_publish")

    def barrier(self, metrics: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Stream-ordered rendezvous: returns the SUM over ranks of ``metrics`` (6 floats) as a by-product."""
        self.bar.zero_()
        if metrics is not None:
            self.bar[:6].copy_(metrics[:6])
        dist.all_reduce(self.bar, op=dist.ReduceOp.SUM, group=self.group)
        return self.bar

    def optimizer_step(self, opt):
        opt.step_peers(self.bufs, self.world, self.parity * self.half)
        self.parity ^= 1

    def close(self):
        for r, q in enumerate(self.peers):
            if q is not None and r != self.rank:
                self._lib.gt_peer_close(self._C.c_void_p(q))
        self.peers = [None] * self.world
        if self.local:
            self._lib.gt_peer_free(self.local)
            self.local = self._C.c_void_p()


class DataParallelStep:
    """step(x_local, y_local) = local fused fwd+loss+bwd -> all-reduce(SUM) of the flat gradient ->
    optimizer step with grad_scale = 1/world.

    With ``overlap=True`` (default on CUDA, world > 1) the gradient is reduced in BUCKETS while
    backward is still running: the library records one CUDA event per bucket as backward finishes
    it (include/groove_b200.h: gt_grad_buckets / gt_grad_bucket_wait), and each group's NCCL
    all-reduce is issued on a communication stream that waits only for that event.  ``compute`` /
    ``bucket_ranges`` / ``wait_bucket`` are injection points for the CPU (gloo) tests.

    ``exchange='p2p'`` (``'auto'``: whenever the ranks can map each other's buffers and the optimizer is fused) replaces the
    all-reduce + optimizer pair by PeerExchange: publish -> barrier -> one kernel that sums the gradient over NVLink peer memory
    in rank order and applies the optimizer.  ``self.exchange`` says which form runs."""

    def __init__(self, model, optimizer, hit_loss_penalty: float, group=None,
                 compute: Optional[Callable] = None, comm_stream: Optional["torch.cuda.Stream"] = None,
                 overlap: Optional[bool] = None, bucket_bytes: int = 512 * 1024,
                 bucket_ranges: Optional[Sequence[Tuple[int, int]]] = None,
                 wait_bucket: Optional[Callable[[int], None]] = None, exchange: str = "nccl", peer=None):
        self.model, self.opt, self.penalty, self.group = model, optimizer, hit_loss_penalty, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.compute = compute
        self.comm_stream = comm_stream
        self.wait_bucket = wait_bucket
        self.groups: Optional[List[Tuple[int, int, int]]] = None
        if hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = 1.0 / self.world
        if exchange not in ("nccl", "p2p", "auto"):
            raise ValueError("exchange must be 'nccl', 'p2p' or 'auto'")
        self.peer: Optional[PeerExchange] = None
        self.exchange = "nccl"
        if peer is not None:                               # injection point of the CPU (gloo) tests: publish / barrier / optimizer_step
            self.peer, self.exchange, overlap, exchange = peer, "p2p", False, "nccl"
        if exchange == "p2p" and compute is not None:
            raise ValueError("exchange='p2p' runs the library's own step (no injected compute)")
        if exchange != "nccl" and self.world > 1 and compute is None:
            if not hasattr(optimizer, "step_peers"):
                if exchange == "p2p":
                    raise ValueError("exchange='p2p' needs FusedSGD / FusedAdam (the exchange is part of the optimizer kernel)")
            else:
                flat = model.flat_parameters()
                px = PeerExchange(flat.numel(), flat.device, group)
                if px.ok:
                    self.peer, self.exchange, overlap = px, "p2p", False
                elif exchange == "p2p":
                    raise RuntimeError("exchange='p2p': the ranks cannot map each other's exchange buffers: " + px.why)
        if overlap is None:
            overlap = self.world > 1 and (compute is None or bucket_ranges is not None)
        if overlap and self.world > 1:
            if bucket_ranges is None:
                bucket_ranges = self._library_buckets()
            self.groups = merge_buckets(bucket_ranges, max(1, bucket_bytes // 4))

    def _library_buckets(self):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        cfg = self.model._cfg()
        cap = 160
        offs, sizes = (C.c_int64 * cap)(), (C.c_int64 * cap)()
        n = lib.gt_grad_buckets(C.byref(cfg), offs, sizes, cap)
        if n < 0:
            raise RuntimeError(lib.gt_last_error().decode())
        _lib.check(lib.gt_grad_events_enable(n), "gt_grad_events_enable")
        dev = self.model.flat_parameters().device
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(dev)
        stream = self.comm_stream
        self.wait_bucket = lambda b: _lib.check(lib.gt_grad_bucket_wait(b, stream.cuda_stream), "gt_grad_bucket_wait")
        return [(int(offs[i]), int(sizes[i])) for i in range(n)]

    def _reduce_overlapped(self, grad):
        works = []
        for off, size, last in self.groups:
            if self.comm_stream is not None:
                with torch.cuda.stream(self.comm_stream):
                    self.wait_bucket(last)            # comm stream waits for this bucket's backward only
                    works.append(dist.all_reduce(grad[off:off + size], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                if self.wait_bucket is not None:
                    self.wait_bucket(last)
                works.append(dist.all_reduce(grad[off:off + size], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()                                  # the compute stream waits for every bucket before the optimizer

    def step(self, x_local, y_local, reduce_metrics: bool = False):
        n_local = x_local.shape[0]
        if hasattr(self.model, "_seq0"):
            self.model._seq0 = self.rank * n_local
        if self.compute is not None:
            metrics, grad = self.compute(x_local, y_local)
        else:
            metrics, _ = self.model.train_step(x_local, y_local, self.penalty)
            grad = self.model.flat_grad()
        if self.peer is not None:
            self.peer.publish(grad)
            summed = self.peer.barrier(metrics if reduce_metrics else None)      # every rank's gradient is published
            self.peer.optimizer_step(self.opt)
            if reduce_metrics:
                m = summed[:6] / self.world
                m[2] = torch.exp(m[3])
                return m
            return metrics
        if self.world > 1:
            if self.groups is not None:
                self._reduce_overlapped(grad)
            else:
                dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            if not hasattr(self.opt, "grad_scale"):
                grad /= self.world
        self.opt.step()
        if reduce_metrics and self.world > 1:
            metrics = combine_metrics(metrics, self.group)
        return metrics
