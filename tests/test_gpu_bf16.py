"""bf16 tensor-core path (fused tcgen05 encoder-layer kernels) against the oracle.
Tolerances per BASELINE.json north_star for the bf16 mode: per-step loss within 2e-3 relative,
thresholded hits agreeing on >= 99.9 % of cells."""
import numpy as np
import pytest
import torch

import groove_oracle as G
from _util import build_model, grads_by_name, rel_err
from transformergrooveinfilling_b200 import FusedSGD

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-3

SHAPES = {
    "c1": (G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 0.47, 0.18),
    "c2": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24),
    "c5enc": (G.GrooveCfg(32, 16, 512, 2, 0, 27, 27), 0.38, 0.24),
    "f48_h32": (G.GrooveCfg(32, 32, 48, 1, 0, 16, 27), 0.5, 0.1),
    "f96_h1": (G.GrooveCfg(32, 1, 96, 2, 0, 16, 27), 0.5, 0.1),
}


@pytest.mark.parametrize("name", sorted(SHAPES))
@pytest.mark.parametrize("n", [4, 7])
def test_eval_forward(name, n):
    cfg, pen, p = SHAPES[name]
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.eval()
    x, y = G.det_batch(cfg, n)
    with torch.no_grad():
        h, v, o = model(x.cuda())
    rh, rv, ro = G.forward_encoder_only(P, cfg, x)
    assert rel_err(h.cpu().numpy(), rh.numpy()) < 3e-2
    assert np.abs(v.cpu().numpy() - rv.numpy()).max() < 2e-2 and np.abs(o.cpu().numpy() - ro.numpy()).max() < 2e-2
    ph, pv, po = model.predict(x.cuda())
    oh, _, _ = G.predict_encoder_only(P, cfg, x)
    # random-init logits hover around 0 (|logit| ~ 0.3), so bf16 rounding flips ~1 % of the near-threshold cells; the 99.9 %
    # agreement of the north star is asserted on trained weights in tests/test_gpu_bf16_exact.py::test_hit_agreement_after_training
    assert (ph.cpu() == oh).float().mean() >= 0.97


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_train_forward_with_dropout(name):
    cfg, pen, p = SHAPES[name]
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(99, step=2, seq0=5).train()
    x, y = G.det_batch(cfg, 6)
    with torch.no_grad():
        h, v, o = model(x.cuda())
    drop = G.DropCtx(p=p, seed=99, step=2, seq0=5, train=True)
    rh, rv, ro = G.output_layer(P, G.encode(P, cfg, x, drop))
    assert rel_err(h.cpu().numpy(), rh.numpy()) < 3e-2
    assert np.abs(v.cpu().numpy() - rv.numpy()).max() < 2e-2


@pytest.mark.parametrize("name,n", [("c1", 5), ("c1", 64), ("c2", 4), ("c2", 64), ("c5enc", 9), ("f48_h32", 3), ("f96_h1", 6)])
def test_train_step_matches_oracle(name, n):
    cfg, pen, p = SHAPES[name]
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(7, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    loss6, grads, _ = G.train_step_oracle(P, cfg, x, y, pen, G.DropCtx(p, 7, 1, 0, True))
    got = metrics.cpu().numpy().astype(np.float64)
    assert abs(got[0] - loss6[0]) / abs(loss6[0]) < LOSS_RTOL, (got, loss6)
    gg = grads_by_name(model)
    worst = ("", 0.0)
    for k, w in grads.items():
        scale = float(w.abs().max())
        if scale < 1e-6:
            continue
        e = float((gg[k] - w).abs().max()) / scale
        if e > worst[1]:
            worst = (k, e)
    # distance to the FP32 restatement of the reference = accumulated bf16 operand rounding: ~3 % of each tensor's max at n=4,
    # falling as 1/sqrt(n) (tools/diag_bf16_grad.py).  The tight gradient check (5e-3) is against the bf16-operand oracle, which
    # rounds where the kernels round: tests/test_gpu_bf16_exact.py::test_train_step_matches_bf16_oracle
    assert worst[1] < (4e-2 if n >= 64 else 0.2), f"gradient mismatch {worst}"


@pytest.mark.parametrize("name,n", [("c1", 5), ("c5enc", 9)])
def test_autograd_path_equals_fused_step_bf16(name, n):
    """The reference's call sequence model(x) -> calculate_loss -> loss.backward() and the single-call fused step run the
    same layer kernels but different tail kernels (edge32.cu: the fused step folds calculate_loss into the tail forward
    and hands dL/dlogits to the tail backward; the autograd path applies the activation derivative from (d_hvo, hvo)):
    with the same dropout masks they must agree to fp32 re-ordering noise."""
    from transformergrooveinfilling_b200 import calculate_loss
    cfg, pen, p = SHAPES[name]
    x, y = [t.cuda() for t in G.det_batch(cfg, n)]
    m1, _ = build_model(cfg, dropout=p, precision="bf16")
    m2, _ = build_model(cfg, dropout=p, precision="bf16")
    m1.set_seed(5, step=3, seq0=0).train(); m2.set_seed(5, step=3, seq0=0).train()
    metrics, hvo = m1.train_step(x, y, pen)
    pred = m2(x)
    out = calculate_loss(pred, y, None, None, pen)
    out[0].backward()
    np.testing.assert_allclose(torch.cat(pred, 2).detach().cpu().numpy(), hvo.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.array([out[0].item(), *out[1:]]), metrics.cpu().numpy(), rtol=2e-6)
    g1, g2 = grads_by_name(m1), grads_by_name(m2)
    for k in g1:
        scale = float(g1[k].abs().max()) + 1e-12
        # the fused tail evaluates sigmoid / tanh / softplus with fast intrinsics, calculate_loss's own kernel with the accurate
        # functions: dL/dlogits differ in the last bits and the bf16 roundings downstream amplify that to ~2e-4 at the input layer
        assert float((g2[k] - g1[k]).abs().max()) / scale < 1e-3, k


def test_edge_gradients_against_oracle():
    """The ends of the path (input layer + positional encoding, final LayerNorm + head + loss and their gradients) are fp32
    kernels (edge32.cu); with one layer in between, the head / final-norm gradients only see the bf16 rounding of that
    layer's output and are compared with the oracle at a tighter tolerance than the layer gradients."""
    cfg = G.GrooveCfg(32, 4, 16, 1, 0, 27, 27)
    model, P = build_model(cfg, dropout=0.0, precision="bf16")
    model.train()
    x, y = G.det_batch(cfg, 33)
    metrics, _ = model.train_step(x.cuda(), y.cuda(), 0.6)
    loss6, grads, _ = G.train_step_oracle(P, cfg, x, y, 0.6, G.DropCtx(0.0, 0, 0, 0, True))
    assert abs(float(metrics[0]) - loss6[0]) / abs(loss6[0]) < LOSS_RTOL
    gg = grads_by_name(model)
    for k in ("OutputLayer.Linear.bias", "OutputLayer.Linear.weight", "Encoder.Encoder.norm.weight", "Encoder.Encoder.norm.bias"):
        scale = float(grads[k].abs().max())
        assert float((gg[k] - grads[k]).abs().max()) / scale < 5e-2, k
    for k in ("InputLayerEncoder.Linear.weight", "InputLayerEncoder.Linear.bias"):
        scale = float(grads[k].abs().max())
        assert float((gg[k] - grads[k]).abs().max()) / scale < 0.1, k


def test_loss_trajectory_bf16_vs_fp32():
    """20 SGD steps from identical weights / data / dropout masks: the bf16 path tracks the fp32 path
    (itself within 1e-4 of the reference) within 2e-3 relative at every step."""
    cfg, pen, p = SHAPES["c2"]
    x, y = [t.cuda() for t in G.det_batch(cfg, 64)]
    traj = {}
    for prec in ("fp32", "bf16"):
        model, _ = build_model(cfg, dropout=p, precision=prec)
        model.set_seed(3).train()
        opt = FusedSGD(model, 0.07)
        t = []
        for _ in range(20):
            m, _ = model.train_step(x, y, pen)
            opt.step()
            t.append(float(m[0]))
        traj[prec] = np.array(t)
    np.testing.assert_allclose(traj["bf16"], traj["fp32"], rtol=LOSS_RTOL)
    assert traj["fp32"][-1] < traj["fp32"][0]


def test_shapes_without_fused_kernels_run_on_gemm_tc():
    """precision='bf16' is available for every configuration: shapes the fused layer kernels are not instantiated for run the
    per-op path with their contractions on the generic tcgen05 GEMM (tests/test_gpu_gemm_tc.py), never an fp32 fallback."""
    import ctypes as C
    from transformergrooveinfilling_b200 import _lib
    cfg = G.GrooveCfg(64, 4, 64, 1, 0, 16, 27)
    model, P = build_model(cfg, dropout=0.0, precision="bf16")
    lib = _lib.load()
    assert lib.gt_path_kind(C.byref(model._cfg())) == _lib.PATH_GEMM_TC
    n0 = lib.gt_launch_count(22)
    x, y = G.det_batch(cfg, 4)
    metrics, _ = model.train_step(x.cuda(), y.cuda(), 1.0)
    assert lib.gt_launch_count(22) > n0
    loss6, _, _ = G.train_step_oracle(P, cfg, x, y, 1.0, G.DropCtx(0.0, 0, 0, 0, True))
    assert abs(float(metrics[0]) - loss6[0]) / abs(loss6[0]) < LOSS_RTOL
