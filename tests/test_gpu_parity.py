"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA library, called through the
drop-in Python API / C ABI, against the CPU oracle on the same deterministic inputs and against the
reference-generated golden fixtures.  fp32 precision mode; tolerances per BASELINE.json north_star:
per-step loss within 1e-4 relative, thresholded hits agreeing on >= 99.9 %."""
import os

import numpy as np
import pytest
import torch

import groove_oracle as G
from golden_cases import CASES, N_TRAJ, digest
from _util import build_model, grads_by_name, params_by_name, rel_err
from transformergrooveinfilling_b200 import FusedAdam, FusedSGD, calculate_loss

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOSS_RTOL = 1e-4          # north_star: fp32/TF32 mode
GRAD_TOL = 2e-4           # max-abs error relative to the tensor's max-abs, per parameter tensor


def _check_grads(got, want, tol=GRAD_TOL, skip_tiny=True):
    worst = ("", 0.0)
    for k, w in want.items():
        g = got[k]
        scale = float(w.abs().max())
        if skip_tiny and scale < 1e-7:       # mathematically-zero gradients (K bias): compare absolutely
            assert float((g - w).abs().max()) < 1e-6, k
            continue
        e = float((g - w).abs().max()) / scale
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < tol, f"gradient mismatch {worst}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_fused_step_matches_golden_and_oracle(name):
    cfg, n, pen, lr = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    model, P = build_model(cfg, dropout=0.0)
    x, y = G.det_batch(cfg, n)
    model.train()
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    m = metrics.cpu().numpy().astype(np.float64)
    hvo = hvo.cpu().numpy()
    nv = cfg.e_tgt // 3
    np.testing.assert_allclose(hvo[..., 0:nv], gold["h"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(hvo[..., nv:2 * nv], gold["v"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(hvo[..., 2 * nv:], gold["o"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(m, gold["loss6"], rtol=LOSS_RTOL)
    loss6, grads, _ = G.train_step_oracle(P, cfg, x, y, pen, G.DropCtx(0.0))
    _check_grads(grads_by_name(model), grads)
    names = [k for k, _ in G.param_shapes(cfg)]
    got = grads_by_name(model)
    dg = np.stack([digest(got[k], i) for i, k in enumerate(names)])
    scale = np.abs(gold["grad_digest"][:, 1:2]) + 1e-6
    np.testing.assert_allclose(dg / scale, gold["grad_digest"] / scale, atol=5e-4)


@pytest.mark.parametrize("name", ["c1_closedhh_testing", "c5_symbolic_encdec"])
def test_autograd_path_equals_fused_path(name):
    """model(x) -> calculate_loss -> loss.backward() (the reference's call sequence) gives the same
    loss / metrics / gradients as the single-call fused step."""
    cfg, n, pen, lr = CASES[name]
    x, y = [t.cuda() for t in G.det_batch(cfg, n)]
    m1, _ = build_model(cfg, dropout=0.0)
    m2, _ = build_model(cfg, dropout=0.0)
    m1.train(); m2.train()
    metrics, _ = m1.train_step(x, y, pen)
    pred = m2(x, torch.cat((torch.zeros_like(y[:, :1]), y[:, :-1]), 1)) if cfg.n_dec > 0 else m2(x)
    out = calculate_loss(pred, y, torch.nn.BCEWithLogitsLoss(reduction="none"), torch.nn.MSELoss(reduction="none"), pen)
    assert out[0].requires_grad and out[0].dim() == 0 and all(isinstance(v, float) for v in out[1:])
    out[0].backward()
    np.testing.assert_allclose(np.array([out[0].item(), *out[1:]]), metrics.cpu().numpy(), rtol=1e-6)
    g1, g2 = grads_by_name(m1), grads_by_name(m2)
    for k in g1:
        np.testing.assert_allclose(g2[k].numpy(), g1[k].numpy(), rtol=1e-4, atol=1e-7, err_msg=k)
    # nn.Parameter.grad views are bound to the flat gradient, so torch.optim works unmodified
    p0 = m2.OutputLayer.Linear.weight.detach().clone()
    torch.optim.SGD(m2.parameters(), lr=0.5).step()
    assert not torch.equal(p0, m2.OutputLayer.Linear.weight.detach())
    np.testing.assert_allclose((p0 - 0.5 * m2.OutputLayer.Linear.weight.grad).cpu().numpy(),
                               m2.OutputLayer.Linear.weight.detach().cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_dropout_masks_match_oracle_generator():
    import ctypes as C
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    n = 1 << 16
    for (seed, step, site, p, idx0) in [(7, 0, 17, 0.24, 0), (2 ** 40 + 5, 9, 1, 0.15, 12345), (1, 3, 600, 0.3, 2 ** 33 + 1)]:
        keep = torch.empty(n, dtype=torch.uint8, device="cuda")
        _lib.check(lib.gt_debug_dropout_mask(seed, step, site, p, idx0, n, keep.data_ptr(), 0))
        torch.cuda.synchronize()
        want = G.dropout_keep(seed, step, site, np.arange(n, dtype=np.uint64) + np.uint64(idx0), p)
        assert (keep.cpu().numpy().astype(bool) == want).all()


@pytest.mark.parametrize("name,p", [("c1_closedhh_testing", 0.18), ("c2_closedhh", 0.24), ("c5_symbolic_encdec", 0.24),
                                    ("odd_small_encdec", 0.3)])
def test_training_step_with_dropout_matches_oracle_with_same_masks(name, p):
    cfg, n, pen, lr = CASES[name]
    model, P = build_model(cfg, dropout=p)
    model.set_seed(1234, step=5, seq0=3)
    x, y = G.det_batch(cfg, n)
    model.train()
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    drop = G.DropCtx(p=p, seed=1234, step=5, seq0=3, train=True)
    loss6, grads, pred = G.train_step_oracle(P, cfg, x, y, pen, drop)
    np.testing.assert_allclose(metrics.cpu().numpy().astype(np.float64), np.array(loss6), rtol=LOSS_RTOL)
    np.testing.assert_allclose(hvo[..., 0:9].cpu().numpy(), pred[0].numpy(), rtol=1e-4, atol=5e-5)
    _check_grads(grads_by_name(model), grads, tol=5e-4)
    # eval mode ignores dropout
    model.eval()
    with torch.no_grad():
        out = model(x.cuda(), G.shift_right(y).cuda()) if cfg.n_dec > 0 else model(x.cuda())
    ref = (G.forward_encdec(P, cfg, x, G.shift_right(y)) if cfg.n_dec > 0 else G.forward_encoder_only(P, cfg, x))
    np.testing.assert_allclose(out[1].cpu().numpy(), ref[1].numpy(), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_loss_trajectory_matches_reference(name, opt):
    cfg, n, pen, lr = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    model, _ = build_model(cfg, dropout=0.0)
    o = FusedSGD(model, lr) if opt == "sgd" else FusedAdam(model, 1e-3)
    x, y = [t.cuda() for t in G.det_batch(cfg, n)]
    model.train()
    traj = []
    for _ in range(N_TRAJ):
        o.zero_grad()
        metrics, _ = model.train_step(x, y, pen)
        o.step()
        traj.append(float(metrics[0]))
    np.testing.assert_allclose(np.array(traj), gold[f"traj_{opt}"], rtol=LOSS_RTOL)
    names = [k for k, _ in G.param_shapes(cfg)]
    got = params_by_name(model)
    dg = np.stack([digest(got[k], i) for i, k in enumerate(names)])
    scale = np.abs(gold[f"param_digest_{opt}"][:, 1:2]) + 1e-6
    keep = np.array([not (opt == "adam" and k.endswith("in_proj_bias")) for k in names])   # see test_oracle_golden
    # Adam divides by sqrt(v): an element whose gradient is rounding noise moves by +- lr per step with the SIGN of that noise, and
    # the LayerNorm / attention gradient partials are combined with fp32 atomics (order varies from run to run).  Measured on the
    # same binary: c5_symbolic_encdec passes 5e-4 in 3 runs of 5 and shows ONE digest of 128 at 1.6e-3 in the other two — the loss
    # trajectory above holds 1e-4 in all of them.  SGD keeps 5e-4.
    np.testing.assert_allclose((dg / scale)[keep], (gold[f"param_digest_{opt}"] / scale)[keep], atol=5e-4 if opt == "sgd" else 4e-3)


@pytest.mark.parametrize("name", sorted(CASES))
def test_predict_matches_reference(name):
    cfg, n, pen, lr = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    model, _ = build_model(cfg, dropout=0.2)           # predict() must switch to eval mode itself
    model.train()
    x, _ = G.det_batch(cfg, n)
    h, v, o = model.predict(x.cuda(), use_thres=True, thres=0.5)
    assert not model.training
    assert h.dtype == (torch.float32 if cfg.n_dec > 0 else torch.int64) and v.dtype == torch.float32
    assert (h.cpu().numpy() == gold["pred_h"]).mean() >= 0.999
    np.testing.assert_allclose(v.cpu().numpy(), gold["pred_v"], rtol=1e-4, atol=3e-5)
    np.testing.assert_allclose(o.cpu().numpy(), gold["pred_o"], rtol=1e-4, atol=3e-5)
    with pytest.raises(NotImplementedError):
        model.predict(x.cuda(), use_pd=True)


@pytest.mark.parametrize("name", ["c5_symbolic_encdec", "odd_small_encdec"])
@pytest.mark.parametrize("n", [1, 37])
def test_kv_cached_decode_equals_literal_32_pass_loop(name, n):
    """gt_predict (incremental decode with per-layer key/value caches) against the reference's literal loop of 32 full
    decoder passes (BGT/models/transformer.py:62-72) run by the same library: same hits, v/o to fp32 reorder."""
    cfg, _, _, _ = CASES[name]
    model, _ = build_model(cfg, dropout=0.1)
    model.eval()
    x, _ = G.det_batch(cfg, n)
    with torch.no_grad():
        fast = model._predict_hvo(x.cuda(), 0.5).cpu().numpy()
        slow = model._predict_hvo(x.cuda(), 0.5, literal=True).cpu().numpy()
    assert (fast[..., :9] == slow[..., :9]).mean() >= 0.999
    agree = (fast[..., :9] == slow[..., :9]).all(axis=(1, 2))       # a flipped hit legitimately changes later steps
    np.testing.assert_allclose(fast[agree][..., 9:], slow[agree][..., 9:], rtol=1e-4, atol=3e-5)
    assert agree.mean() >= 0.97


@pytest.mark.parametrize("n", [1, 3, 33, 130])
def test_ragged_batch_sizes_and_batch_independence(n):
    """Sequences are independent: the outputs for a batch equal the outputs of its rows run alone,
    for any N (the DataLoader's last batch is ragged)."""
    cfg = G.GrooveCfg(32, 4, 16, 2, 0, 16, 27)
    model, P = build_model(cfg, dropout=0.0)
    model.eval()
    x, y = G.det_batch(cfg, n)
    with torch.no_grad():
        h, v, o = model(x.cuda())
        h1, v1, o1 = model(x[n - 1:].cuda())
    assert h.shape == (n, 32, 9)
    np.testing.assert_allclose(h[n - 1:].cpu().numpy(), h1.cpu().numpy(), rtol=1e-5, atol=1e-6)
    ref = G.forward_encoder_only(P, cfg, x)
    np.testing.assert_allclose(v.cpu().numpy(), ref[1].numpy(), rtol=1e-4, atol=2e-5)
    model.train()
    metrics, _ = model.train_step(x.cuda(), y.cuda(), 0.5)
    loss6, grads, _ = G.train_step_oracle(P, cfg, x, y, 0.5, G.DropCtx(0.0))
    np.testing.assert_allclose(metrics.cpu().numpy().astype(np.float64), np.array(loss6), rtol=LOSS_RTOL)
    _check_grads(grads_by_name(model), grads)


def test_large_batch_properties():
    """At a size the oracle cannot run in seconds: (1) mean-of-shard-means == global mean (the DP
    identity), (2) gradients of a batch made of two copies equal the gradients of one copy."""
    cfg = G.GrooveCfg(32, 16, 512, 6, 0, 16, 27)
    model, _ = build_model(cfg, dropout=0.0)
    model.train()
    n = 4096
    x, y = [t.cuda() for t in G.det_batch(cfg, n)]
    m_all, _ = model.train_step(x, y, 0.38)
    g_all = model.flat_grad().clone()
    m_a, _ = model.train_step(x[: n // 2], y[: n // 2], 0.38)
    g_a = model.flat_grad().clone()
    m_b, _ = model.train_step(x[n // 2:], y[n // 2:], 0.38)
    g_b = model.flat_grad().clone()
    np.testing.assert_allclose(((m_a + m_b) / 2)[[0, 1, 3, 4, 5]].cpu().numpy(), m_all[[0, 1, 3, 4, 5]].cpu().numpy(), rtol=2e-5)
    assert rel_err(((g_a + g_b) / 2).cpu().numpy(), g_all.cpu().numpy()) < 2e-4
    m_2, _ = model.train_step(torch.cat((x[:64], x[:64])), torch.cat((y[:64], y[:64])), 0.38)
    g_2 = model.flat_grad().clone()
    m_1, _ = model.train_step(x[:64], y[:64], 0.38)
    assert rel_err(g_2.cpu().numpy(), model.flat_grad().cpu().numpy()) < 1e-4
    np.testing.assert_allclose(m_2.cpu().numpy(), m_1.cpu().numpy(), rtol=1e-5)


def test_optimizer_kernels_match_torch():
    torch.manual_seed(0)
    cfg = G.GrooveCfg(32, 4, 16, 1, 0, 16, 27)
    for kind in ("sgd", "adam"):
        model, _ = build_model(cfg, dropout=0.0)
        ref_p = model.flat_parameters().detach().clone().requires_grad_(True)
        ropt = torch.optim.SGD([ref_p], lr=0.07) if kind == "sgd" else torch.optim.Adam([ref_p], lr=3e-3)
        fopt = FusedSGD(model, 0.07) if kind == "sgd" else FusedAdam(model, 3e-3)
        for _ in range(5):
            g = torch.randn_like(ref_p) * 0.1
            ref_p.grad = g.clone()
            model.flat_grad().copy_(g)
            ropt.step(); fopt.step()
        np.testing.assert_allclose(model.flat_parameters().detach().cpu().numpy(), ref_p.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
        sd = fopt.state_dict()
        assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["lr"] in (0.07, 3e-3)
        if kind == "adam":
            assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
            f2 = FusedAdam(model, 1.0); f2.load_state_dict(sd)
            assert f2._t == 5 and torch.equal(f2._m[:-1], fopt._m[:-1]) and f2.param_groups[0]["lr"] == 3e-3


def test_error_behaviour():
    cfg = G.GrooveCfg(32, 4, 16, 1, 0, 16, 27)
    model, _ = build_model(cfg, dropout=0.0)
    with pytest.raises(ValueError):
        model(torch.zeros(2, 16, 16, device="cuda"))          # T != 32 (the reference raises too)
    with pytest.raises(ValueError):
        model(torch.zeros(0, 32, 16, device="cuda"))
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 32, 16))                           # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        build_model(G.GrooveCfg(30, 4, 16, 1, 0, 16, 27))       # d_model % nhead != 0, like nn.MultiheadAttention
