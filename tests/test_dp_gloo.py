"""World-size-2 gloo test (CPU) of the data-parallel step logic: shard -> local gradient ->
all-reduce(SUM) -> 1/world -> update equals the single-process step on the full batch.  The local
gradient is produced by the oracle (test infrastructure) because there is no GPU here."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import groove_oracle as G

CFG = G.GrooveCfg(16, 2, 8, 1, 0, 16, 27)
PEN, LR, N = 0.5, 0.1, 8


class _FlatSGD:
    def __init__(self, flat, lr):
        self.flat, self.lr, self.grad_scale, self.grad = flat, lr, 1.0, None

    def step(self):
        self.flat -= self.lr * self.grad_scale * self.grad


def _names():
    return [k for k, _ in G.param_shapes(CFG)]


def _flatten(d):
    return torch.cat([d[k].reshape(-1) for k in _names()])


def _worker(rank, world, port, out, bucketed=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from transformergrooveinfilling_b200.dp import DataParallelStep, shard_bounds
    P = G.det_params(CFG)
    flat = _flatten(P).clone()
    x, y = G.det_batch(CFG, N)
    lo, hi = shard_bounds(N, rank, world)
    opt = _FlatSGD(flat, LR)

    def compute(xl, yl):
        loss6, grads, _ = G.train_step_oracle(P, CFG, xl, yl, PEN, G.DropCtx(0.0))
        opt.grad = _flatten(grads)
        return torch.tensor(loss6, dtype=torch.float32), opt.grad

    if bucketed:
        # descending partition of the flat gradient in "completion order", like gt_grad_buckets; the wait hook
        # records the order in which groups were released
        n = flat.numel()
        cuts = [n, n - 40, n // 2, 7, 0]
        ranges = [(cuts[i + 1], cuts[i] - cuts[i + 1]) for i in range(len(cuts) - 1)]
        waited = []
        dp = DataParallelStep(None, opt, PEN, compute=compute, bucket_ranges=ranges, bucket_bytes=4 * 64,
                              wait_bucket=waited.append)
        assert dp.groups is not None and sum(g[1] for g in dp.groups) == n
    else:
        dp = DataParallelStep(None, opt, PEN, compute=compute)
    metrics = dp.step(x[lo:hi], y[lo:hi], reduce_metrics=True)
    if bucketed:
        assert waited == [g[2] for g in dp.groups] and waited[-1] == len(ranges) - 1
    if rank == 0:
        torch.save({"flat": flat, "metrics": metrics, "scale": opt.grad_scale}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bucketed", [False, True])
def test_dp_step_equals_single_process(tmp_path, bucketed):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000) + (7 if bucketed else 0)
    mp.spawn(_worker, args=(2, port, out, bucketed), nprocs=2, join=True)
    got = torch.load(out)
    P = G.det_params(CFG)
    x, y = G.det_batch(CFG, N)
    loss6, grads, _ = G.train_step_oracle(P, CFG, x, y, PEN, G.DropCtx(0.0))
    want = _flatten(P) - LR * _flatten(grads)
    assert got["scale"] == 0.5
    np.testing.assert_allclose(got["flat"].numpy(), want.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got["metrics"].numpy(), np.array(loss6), rtol=2e-5)


class _GlooPeer:
    """Stand-in for dp.PeerExchange on CPU: the exchange buffers (two halves, alternating by step) live in this process and the
    "peer reads" of the optimizer kernel are an all_gather of the half the step published — same protocol, same rank-ordered sum."""

    def __init__(self, n, opt, world):
        self.half = [torch.full((n,), float("nan")), torch.full((n,), float("nan"))]
        self.parity, self.opt, self.world, self.log = 0, opt, world, []
        self.bar = torch.zeros(8)

    def publish(self, grad):
        self.half[self.parity].copy_(grad)
        self.log.append(("publish", self.parity))

    def barrier(self, metrics=None):
        self.bar.zero_()
        if metrics is not None:
            self.bar[:6].copy_(metrics[:6])
        dist.all_reduce(self.bar)
        self.log.append(("barrier", self.parity))
        return self.bar

    def optimizer_step(self, opt):
        assert opt is self.opt
        bufs = [torch.empty_like(self.half[0]) for _ in range(self.world)]
        dist.all_gather(bufs, self.half[self.parity])
        total = bufs[0].clone()
        for b in bufs[1:]:                               # rank order
            total += b
        opt.grad = total
        opt.step()
        self.log.append(("step", self.parity))
        self.half[self.parity].fill_(float("nan"))       # a later read of a stale half would poison the parameters
        self.parity ^= 1


def _worker_p2p(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from transformergrooveinfilling_b200.dp import DataParallelStep, shard_bounds
    names = _names()
    shapes = dict(G.param_shapes(CFG))
    flat = _flatten(G.det_params(CFG)).clone()
    x, y = G.det_batch(CFG, N)
    lo, hi = shard_bounds(N, rank, world)
    opt = _FlatSGD(flat, LR)

    def unflatten(f):
        out_, o = {}, 0
        for k in names:
            n = int(np.prod(shapes[k]))
            out_[k] = f[o:o + n].reshape(shapes[k]).clone()
            o += n
        return out_

    def compute(xl, yl):
        loss6, grads, _ = G.train_step_oracle(unflatten(flat), CFG, xl, yl, PEN, G.DropCtx(0.0))
        return torch.tensor(loss6, dtype=torch.float32), _flatten(grads)

    peer = _GlooPeer(flat.numel(), opt, world)
    dp = DataParallelStep(None, opt, PEN, compute=compute, peer=peer)
    assert dp.exchange == "p2p" and dp.groups is None
    traj = [dp.step(x[lo:hi], y[lo:hi], reduce_metrics=True).clone() for _ in range(3)]
    assert peer.log == [(w, p) for p in (0, 1, 0) for w in ("publish", "barrier", "step")]
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)           # replicas bit-identical
    if rank == 0:
        torch.save({"flat": flat, "traj": torch.stack(traj)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_protocol_equals_single_process(tmp_path):
    """exchange = 'p2p' (dp.PeerExchange; on the GPU: csrc/peer_opt.cu, tests/test_gpu_multi.py): publish -> rendezvous that carries
    the metrics -> optimizer over the rank-ordered sum, with the exchange buffer double buffered by step parity — three steps of two
    gloo ranks against the single-process run of the whole batch."""
    out = str(tmp_path / "p2p.pt")
    mp.spawn(_worker_p2p, args=(2, 29500 + (os.getpid() % 2000) + 13, out), nprocs=2, join=True)
    got = torch.load(out)
    P = G.det_params(CFG)
    x, y = G.det_batch(CFG, N)
    losses = []
    for _ in range(3):
        loss6, grads, _ = G.train_step_oracle(P, CFG, x, y, PEN, G.DropCtx(0.0))
        P = {k: v - LR * grads[k] for k, v in P.items()}
        losses.append(loss6)
    np.testing.assert_allclose(got["flat"].numpy(), _flatten(P).numpy(), rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(got["traj"].numpy(), np.array(losses), rtol=5e-5)


def test_shard_bounds():
    from transformergrooveinfilling_b200.dp import shard_bounds
    assert [shard_bounds(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)


def test_merge_buckets():
    from transformergrooveinfilling_b200.dp import merge_buckets
    r = [(90, 10), (60, 30), (30, 30), (5, 25), (0, 5)]
    assert merge_buckets(r, 1) == [(90, 10, 0), (60, 30, 1), (30, 30, 2), (5, 25, 3), (0, 5, 4)]
    assert merge_buckets(r, 40) == [(60, 40, 1), (0, 60, 4)]
    assert merge_buckets(r, 1000) == [(0, 100, 4)]
    with pytest.raises(ValueError):
        merge_buckets([(90, 10), (50, 30)], 1)
