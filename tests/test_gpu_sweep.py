"""Sweep packing: members trained concurrently on one GPU (one stream + one host thread each) give
the results of training each configuration alone."""
import numpy as np
import pytest
import torch

import groove_oracle as G
from transformergrooveinfilling_b200 import SweepPacker, params_from_config

pytestmark = pytest.mark.gpu

CONFIGS = [
    dict(batch_size=16, d_model=32, dim_feedforward=64, dropout=0.2, optimizer_algorithm="sgd", learning_rate=0.05, n_heads=4,
         num_encoder_decoder_layers=2, encoder_only=1, experiment="InfillingClosedHH", hit_loss_penalty=0.4),
    dict(batch_size=32, d_model=32, dim_feedforward=512, dropout=0.24, optimizer_algorithm="adam", learning_rate=1e-3, n_heads=16,
         num_encoder_decoder_layers=3, encoder_only=1, experiment="InfillingClosedHH", hit_loss_penalty=0.38),
    dict(batch_size=8, d_model=64, dim_feedforward=32, dropout=0.1, optimizer_algorithm="sgd", learning_rate=0.02, n_heads=2,
         num_encoder_decoder_layers=2, encoder_only=1, experiment="InfillingClosedHH", hit_loss_penalty=0.9),
    dict(batch_size=16, d_model=256, dim_feedforward=64, dropout=0.15, optimizer_algorithm="sgd", learning_rate=0.04, n_heads=16,
         num_encoder_decoder_layers=2, encoder_only=1, experiment="InfillingClosedHH", hit_loss_penalty=1.0),
]


def _dataset(n=96):
    cfg = G.GrooveCfg(32, 4, 16, 1, 0, 16, 27)
    return G.det_batch(cfg, n)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_packed_members_match_solo_runs(precision):
    x, y = _dataset()
    steps = 9                                  # more than one epoch for the batch-32 member (96 / 32 = 3 steps per epoch)

    def run(concurrent):
        torch.manual_seed(0)                   # initialize_model draws the reference's random init
        pk = SweepPacker(CONFIGS, x, y, "cuda", precision=precision, seed=5)
        pk.run(steps, concurrent=concurrent)
        return [h.numpy() for h in pk.history()], pk

    packed, pk = run(True)
    solo, _ = run(False)
    assert len(packed) == len(CONFIGS)
    for a, b in zip(packed, solo):
        assert a.shape == (steps, 6) and np.isfinite(a).all()
        # same kernels, same seeds, same data order: only the order of fp32 atomics differs between the two runs.  In bf16
        # mode a last-bit difference in a master weight can flip a bf16 operand rounding, so later steps drift apart
        # (6e-3 after 9 large SGD steps); the first step has no such history and must agree tightly in both modes.
        np.testing.assert_allclose(a[0], b[0], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(a, b, rtol=2e-4 if precision == "fp32" else 3e-2, atol=1e-6)
    assert 0 <= pk.best() < len(CONFIGS)
    assert all(m.steps == steps for m in pk.members)


def test_members_learn_and_are_isolated():
    x, y = _dataset(64)
    torch.manual_seed(0)
    pk = SweepPacker(CONFIGS[:2], x, y, "cuda", precision="bf16", seed=1)
    p0 = [m.model.flat_parameters().detach().clone() for m in pk.members]
    pk.run(30)
    h = pk.history()
    for m, before, hist in zip(pk.members, p0, h):
        assert not torch.equal(before, m.model.flat_parameters().detach())
        assert float(hist[-5:, 0].mean()) < float(hist[:5, 0].mean())
    # distinct parameter buffers (no aliasing between members)
    assert pk.members[0].model.flat_parameters().data_ptr() != pk.members[1].model.flat_parameters().data_ptr()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_multi_step_call_equals_python_loop(precision):
    """gt_train_steps (device-side row gather + train step + optimizer for a whole run of batches in one library call) takes the
    same steps as the per-step Python loop: same permutations, ragged last batch, dropout counters, SGD and Adam updates."""
    x, y = _dataset(100)                        # 100 sequences: batch 16 -> 6 full batches + a ragged batch of 4 per epoch
    steps = 17                                  # two full epochs of the batch-16 member and the start of a third

    def run(fused):
        torch.manual_seed(0)
        pk = SweepPacker(CONFIGS[:3], x, y, "cuda", precision=precision, seed=9)
        pk.run(steps, concurrent=False, fused=fused)
        return [h.numpy() for h in pk.history()], pk

    a, pka = run(True)
    b, pkb = run(False)
    for ha, hb, ma, mb in zip(a, b, pka.members, pkb.members):
        assert ha.shape == (steps, 6)
        np.testing.assert_allclose(ha[0], hb[0], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(ha, hb, rtol=2e-4 if precision == "fp32" else 3e-2, atol=1e-6)
        assert ma.sequences == mb.sequences and ma.model._step == mb.model._step
        pa, pb = ma.model.flat_parameters().detach(), mb.model.flat_parameters().detach()
        # 17 optimizer steps apart, the two runs differ by the order of their fp32 gradient atomics only (1.4e-4 measured)
        assert float((pa - pb).abs().max()) / float(pb.abs().max()) < (1e-3 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("optimizer", ["sgd", "adam"])
def test_graph_replay_equals_fused_call(optimizer):
    """gt_graph_train_create / gt_graph_launch: the captured step (row gather + train step + optimizer, with the dropout step,
    the Adam step count, the row offset and the metrics slot read from device counters) takes the same steps as gt_train_steps:
    same dropout masks (keys derived on the device from the counter), same Adam bias corrections, ragged last batch, epochs."""
    x, y = _dataset(100)                        # batch 16 -> 6 graph replays + one ragged batch of 4 per epoch
    steps = 17
    cfgs = [dict(CONFIGS[0], optimizer_algorithm=optimizer, learning_rate=0.05 if optimizer == "sgd" else 1e-3),
            dict(CONFIGS[1], optimizer_algorithm=optimizer, learning_rate=0.05 if optimizer == "sgd" else 1e-3, batch_size=32)]

    def run(graph):
        torch.manual_seed(0)
        pk = SweepPacker(cfgs, x, y, "cuda", precision="bf16", seed=9)
        assert all(m.graph_capable() for m in pk.members)
        pk.run(steps, concurrent=False, graph=graph)
        return [h.numpy() for h in pk.history()], pk

    a, pka = run(True)
    b, pkb = run(False)
    for ha, hb, ma, mb in zip(a, b, pka.members, pkb.members):
        assert ha.shape == (steps, 6) and np.isfinite(ha).all()
        np.testing.assert_allclose(ha[0], hb[0], rtol=1e-5, atol=1e-7)        # first step: no history, identical masks
        np.testing.assert_allclose(ha, hb, rtol=3e-2, atol=1e-6)              # bf16 trajectory tolerance of the tests above
        assert ma.sequences == mb.sequences and ma.model._step == mb.model._step and ma.epoch == mb.epoch
        assert getattr(ma.optimizer, "_t", 0) == getattr(mb.optimizer, "_t", 0)
        pa, pb = ma.model.flat_parameters().detach(), mb.model.flat_parameters().detach()
        assert float((pa - pb).abs().max()) / float(pb.abs().max()) < 2e-2
        assert ma._graph is not None and getattr(mb, "_graph", None) is None


def test_graph_replay_dropout_masks_follow_the_device_counter():
    """With learning rate 0 the parameters never move, so step k of a replayed run must reproduce step k of the eager run
    EXACTLY up to fp32 summation order — any mismatch of a dropout key (derived on the device from counters[0]) would change the
    loss by far more than that."""
    x, y = _dataset(96)
    cfg = dict(CONFIGS[1], optimizer_algorithm="sgd", learning_rate=0.0, batch_size=16, dropout=0.3)

    def run(graph):
        torch.manual_seed(0)
        pk = SweepPacker([cfg], x, y, "cuda", precision="bf16", seed=4)
        pk.run(12, concurrent=False, graph=graph)
        return pk.history()[0].numpy()

    a, b = run(True), run(False)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7)
    assert len({round(float(v), 6) for v in a[:, 0]}) > 6      # different batches and masks per step: the losses differ


def test_graph_members_packed_concurrently():
    x, y = _dataset(96)
    cfgs = [dict(CONFIGS[1], batch_size=16), CONFIGS[2], CONFIGS[3]]      # fused d_model = 32 | per-op d_model = 64 | fused d_model = 256

    def run(concurrent):
        torch.manual_seed(0)
        pk = SweepPacker(cfgs, x, y, "cuda", precision="bf16", seed=5)
        pk.run(9, concurrent=concurrent, graph=True)
        return [h.numpy() for h in pk.history()], pk

    packed, pk = run(True)
    solo, _ = run(False)
    assert [m.graph_capable() for m in pk.members] == [True, True, True]
    for a, b in zip(packed, solo):
        np.testing.assert_allclose(a[0], b[0], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(a, b, rtol=3e-2, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_on_the_per_op_paths(precision):
    """The per-op kernels (fp32 SIMT path; bf16 path with the contractions on the generic tcgen05 GEMM and mma.sync attention)
    also derive their dropout keys from the device counter: with learning rate 0 a replayed run reproduces the eager losses."""
    x, y = _dataset(96)
    cfg = dict(CONFIGS[2], optimizer_algorithm="sgd", learning_rate=0.0, batch_size=16, dropout=0.3)      # d_model = 64, 2 heads of 32

    def run(graph):
        torch.manual_seed(0)
        pk = SweepPacker([cfg], x, y, "cuda", precision=precision, seed=4)
        assert pk.members[0].graph_capable()
        pk.run(12, concurrent=False, graph=graph)
        return pk.history()[0].numpy()

    a, b = run(True), run(False)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7)
    assert len({round(float(v), 6) for v in a[:, 0]}) > 6


def test_graph_replay_on_the_fused_d256_path():
    """The weight-streaming d_model = 256 kernels (tc256*.cu, edge256.cu) derive their dropout keys from the device counter too
    (common.cuh: DropArgsView): with learning rate 0 a replayed run reproduces the eager losses."""
    x, y = _dataset(96)
    cfg = dict(CONFIGS[3], optimizer_algorithm="sgd", learning_rate=0.0, batch_size=16, dropout=0.3)

    def run(graph):
        torch.manual_seed(0)
        pk = SweepPacker([cfg], x, y, "cuda", precision="bf16", seed=4)
        assert pk.members[0].graph_capable()
        pk.run(12, concurrent=False, graph=graph)
        return pk.history()[0].numpy()

    a, b = run(True), run(False)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7)
    assert len({round(float(v), 6) for v in a[:, 0]}) > 6


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_encoder_decoder(precision):
    """gt_graph_train_create on a GrooveTransformer (encoder-decoder, BGT/models/transformer.py:9-46): the decoder blocks of the
    fused kernels (causal / cross attention, FFN block modes) and the per-op kernels read the device-resident dropout step, and
    the shifted target (train.py:130-131) is built inside the graph.  Replays with lr = 0 must reproduce eager gt_train_step."""
    import ctypes as C
    import groove_oracle as G
    from _util import build_model
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    cfg = G.GrooveCfg(32, 16, 64, 1, 2, 27, 27)
    model, _ = build_model(cfg, dropout=0.25, precision=precision)
    model.set_seed(21, 0, 0).train()
    n, steps = 12, 5
    xs, ys = zip(*[G.det_batch(cfg, n, tag=40 + i) for i in range(steps)])
    eager = []
    for i in range(steps):
        m, _ = model.train_step(xs[i].cuda(), ys[i].cuda(), 0.5)
        eager.append(m.cpu().numpy().copy())
    dev = torch.device("cuda")
    xbuf, ybuf = torch.zeros(n, 32, 27, device=dev), torch.zeros(n, 32, 27, device=dev)
    grads = torch.zeros_like(model.flat_parameters().detach())
    metrics, hvo = torch.zeros(6, device=dev), torch.empty(n, 32, 27, device=dev)
    ws = model._workspace(n, 1, dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    ring = torch.zeros(steps, 6, device=dev)
    handle = C.c_void_p()
    c = model._cfg()
    side = torch.cuda.Stream()                  # stream capture is not allowed on the legacy default stream
    sp = side.cuda_stream
    torch.cuda.synchronize()
    _lib.check(lib.gt_graph_train_create(C.byref(c), _lib.ptr(model.flat_parameters()), _lib.ptr(model._pe_flat()), _lib.ptr(xbuf), _lib.ptr(ybuf),
                                         n, 0.5, _lib.ptr(grads), _lib.ptr(metrics), _lib.ptr(hvo), _lib.ptr(ws), ws.numel(), 0, 0.0, None, None, 21,
                                         _lib.ptr(counters), None, None, None, _lib.ptr(ring), steps, sp, C.byref(handle)), "gt_graph_train_create")
    for i in range(steps):
        with torch.cuda.stream(side):
            xbuf.copy_(xs[i].cuda()); ybuf.copy_(ys[i].cuda())
            _lib.check(lib.gt_graph_launch(handle, 1, sp), "gt_graph_launch")
    torch.cuda.synchronize()
    _lib.check(lib.gt_graph_destroy(handle), "gt_graph_destroy")
    np.testing.assert_allclose(ring.cpu().numpy(), np.stack(eager), rtol=2e-5, atol=1e-7)
    assert int(counters[0]) == steps and len({round(float(v), 6) for v in ring[:, 0]}) == steps
