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

from typing import Callable, Optional, Tuple

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


class DataParallelStep:
    """step(x_local, y_local) = local fused fwd+loss+bwd -> all-reduce(SUM) flat grad -> optimizer
    step with grad_scale = 1/world.  ``compute`` defaults to ``model.train_step``; tests inject a CPU
    stand-in to exercise the collective logic over gloo."""

    def __init__(self, model, optimizer, hit_loss_penalty: float, group=None,
                 compute: Optional[Callable] = None, comm_stream: Optional["torch.cuda.Stream"] = None):
        self.model, self.opt, self.penalty, self.group = model, optimizer, hit_loss_penalty, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.compute = compute
        self.comm_stream = comm_stream
        if hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = 1.0 / self.world

    def step(self, x_local, y_local, reduce_metrics: bool = False):
        n_local = x_local.shape[0]
        if hasattr(self.model, "_seq0"):
            self.model._seq0 = self.rank * n_local
        if self.compute is not None:
            metrics, grad = self.compute(x_local, y_local)
        else:
            metrics, _ = self.model.train_step(x_local, y_local, self.penalty)
            grad = self.model.flat_grad()
        if self.world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            if not hasattr(self.opt, "grad_scale"):
                grad /= self.world
        self.opt.step()
        if reduce_metrics and self.world > 1:
            metrics = combine_metrics(metrics, self.group)
        return metrics
