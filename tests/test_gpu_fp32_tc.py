"""precision = 'fp32_tc' (GT_PREC_FP32_TC): the 1e-4 parity mode of BASELINE.json north_star ("fp32/TF32 mode") ON the tensor
cores.  Every Linear contraction runs on the generic tcgen05 GEMM with its fp32 operands split exactly into three bf16 terms
(x = x0 + x1 + x2) and the six products down to 2^-18 contracted with fp32 accumulation — the bf16 form of "3xTF32"; the x0.y0
products keep an accumulator to themselves because the tensor core truncates on every accumulation (gemm_tc.cu).  Attention,
LayerNorm, loss and optimizers are the fp32 kernels of precision = 'fp32'.

Unit level: gt_debug_gemm(tc=2) against a float64 product of the UNROUNDED operands in the three operand layouts the model
uses — and against tc=1 on the same inputs, to show the split form is what ran (bf16 operands are ~1000x further away).
Model level: the fp32-mode assertions of tests/test_gpu_parity.py (reference-generated goldens: outputs, six loss metrics,
gradients, 20-step SGD / Adam trajectories at 1e-4, predict) repeated in this mode."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import groove_oracle as G
from golden_cases import CASES, N_TRAJ
from _util import build_model, grads_by_name, rel_err
from test_gpu_gemm_tc import _gemm
from test_gpu_parity import GOLD, LOSS_RTOL, _check_grads
from transformergrooveinfilling_b200 import FusedAdam, FusedSGD, _lib

pytestmark = pytest.mark.gpu
SPLIT_TOL = 1e-5           # max-abs error relative to the result's max-abs; measured <= 1e-6 (bf16 operands: ~3e-3)
# Gradients against the fp32 oracle: 2e-4 of the tensor's max-abs like precision = 'fp32'.  Measured 0.5 - 2e-6 (FFMA kernels:
# 5e-7): tools/diag_fp32_tc.py -> profiles/r03/r03_diag_fp32_tc_grads.txt.
# 20-step trajectories: 1e-4 like precision = 'fp32', except the two FULL-DEPTH d_model = 256 cases (6 / 11 layers on a 3-sequence
# batch), which oracle/golden_cases.py already documents as chaotic (at the yaml learning rates float32 and float64 runs of the
# same arithmetic part by 2 - 10 %): 2e-6 instead of 5e-7 per step grows to 7e-4 over 20 Adam steps there.  Their first steps
# and their 2-layer versions hold 1e-4.
TRAJ_TOL = {"c3_kicksnares_full": 2e-3, "c4_random_large_full": 2e-3}

SHAPES = [(128, 32, 32), (4 * 32, 768, 256), (13 * 32, 256, 512), (1000, 100, 72), (33, 40, 33), (2048, 768, 256),
          (4099, 100, 72), (2048 + 33, 512, 512), (64 * 32, 32, 32), (2500, 1024, 192)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_linear_and_dgrad_forms(m, n, k):
    torch.manual_seed(m + n + k)
    x, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    ref = (x.double() @ w.double().T).cpu().numpy()
    out = torch.full((m, n), float("nan"), device="cuda")
    _gemm(2, x, k, 1, w, k, 1, out, n, m, n, k)
    e_split = rel_err(out.cpu().numpy(), ref)
    _gemm(1, x, k, 1, w, k, 1, out, n, m, n, k)
    e_bf16 = rel_err(out.cpu().numpy(), ref)
    assert e_split < SPLIT_TOL and e_split < e_bf16 / 200, (e_split, e_bf16)
    wt = w.T.contiguous()                                   # dX = dY W: B MN-major
    _gemm(2, x, k, 1, wt, 1, n, out, n, m, n, k)
    assert rel_err(out.cpu().numpy(), ref) < SPLIT_TOL


@pytest.mark.parametrize("tokens,n_out,n_in,chunk", [(128, 96, 32, 0), (4096, 768, 256, 2048), (4099 * 32, 64, 64, 2048),
                                                      (7 * 32, 512, 256, 64), (1000, 40, 72, 256)])
def test_wgrad_form_split_k(tokens, n_out, n_in, chunk):
    torch.manual_seed(tokens + n_out)
    dy, x = torch.randn(tokens, n_out, device="cuda"), torch.randn(tokens, n_in, device="cuda")
    dw = torch.ones(n_out, n_in, device="cuda")
    _gemm(2, dy, 1, n_out, x, 1, n_in, dw, n_in, n_out, n_in, tokens, flags=4, split=chunk)
    ref = 1.0 + (dy.double().T @ x.double())
    assert rel_err(dw.cpu().numpy(), ref.cpu().numpy()) < SPLIT_TOL


def test_epilogues():
    """bias + ReLU + dropout / ReLU mask / residual / accumulate: same elements kept, same values as the FFMA kernel."""
    torch.manual_seed(11)
    m, n, k = 8 * 32, 512, 256
    x, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    bias, res, mask = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda"), torch.randn(m, n, device="cuda")
    for kw in (dict(flags=1, bias=bias, drop_p=0.24, row0=96), dict(mask=mask, ld_mask=n, mask_scale=1.25),
               dict(residual=res, ld_res=n, bias=bias), dict(flags=2)):
        outs = []
        for tc in (0, 2):
            c = torch.full((m, n), 0.5, device="cuda")
            _gemm(tc, x, k, 1, w, k, 1, c, n, m, n, k, **kw)
            outs.append(c.cpu().numpy())
        assert rel_err(outs[1], outs[0]) < SPLIT_TOL, kw.keys()
        if kw.get("drop_p"):
            assert ((outs[0] == 0) == (outs[1] == 0)).mean() > 0.9999


def test_path_kind():
    lib = _lib.load()
    for name in ("c2_closedhh", "c4_random_large_l2", "c5_symbolic_encdec"):
        model, _ = build_model(CASES[name][0], dropout=0.0, precision="fp32_tc")
        assert lib.gt_path_kind(C.byref(model._cfg())) == _lib.PATH_GEMM_TC_SPLIT


@pytest.mark.parametrize("name", sorted(CASES))
def test_step_matches_golden_and_oracle(name):
    cfg, n, pen, lr = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    model, P = build_model(cfg, dropout=0.0, precision="fp32_tc")
    x, y = G.det_batch(cfg, n)
    model.train()
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    hvo = hvo.cpu().numpy()
    nv = cfg.e_tgt // 3
    np.testing.assert_allclose(hvo[..., 0:nv], gold["h"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(hvo[..., nv:2 * nv], gold["v"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(hvo[..., 2 * nv:], gold["o"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(metrics.cpu().numpy().astype(np.float64), gold["loss6"], rtol=LOSS_RTOL)
    _, grads, _ = G.train_step_oracle(P, cfg, x, y, pen, G.DropCtx(0.0))
    _check_grads(grads_by_name(model), grads)


@pytest.mark.parametrize("name,p", [("c2_closedhh", 0.24), ("c5_symbolic_encdec", 0.24)])
def test_step_with_dropout_matches_oracle_with_same_masks(name, p):
    cfg, n, pen, lr = CASES[name]
    model, P = build_model(cfg, dropout=p, precision="fp32_tc")
    model.set_seed(1234, step=5, seq0=3)
    x, y = G.det_batch(cfg, n)
    model.train()
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    loss6, grads, pred = G.train_step_oracle(P, cfg, x, y, pen, G.DropCtx(p=p, seed=1234, step=5, seq0=3, train=True))
    np.testing.assert_allclose(metrics.cpu().numpy().astype(np.float64), np.array(loss6), rtol=LOSS_RTOL)
    _check_grads(grads_by_name(model), grads, tol=5e-4)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_loss_trajectory_matches_reference(name, opt):
    cfg, n, pen, lr = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    model, _ = build_model(cfg, dropout=0.0, precision="fp32_tc")
    o = FusedSGD(model, lr) if opt == "sgd" else FusedAdam(model, 1e-3)
    x, y = [t.cuda() for t in G.det_batch(cfg, n)]
    model.train()
    traj = []
    for _ in range(N_TRAJ):
        o.zero_grad()
        metrics, _ = model.train_step(x, y, pen)
        o.step()
        traj.append(float(metrics[0]))
    np.testing.assert_allclose(np.array(traj)[:3], gold[f"traj_{opt}"][:3], rtol=LOSS_RTOL)
    np.testing.assert_allclose(np.array(traj), gold[f"traj_{opt}"], rtol=TRAJ_TOL.get(name, LOSS_RTOL))


@pytest.mark.parametrize("name", sorted(CASES))
def test_predict_matches_reference(name):
    cfg, n, pen, lr = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    model, _ = build_model(cfg, dropout=0.2, precision="fp32_tc")
    x, _ = G.det_batch(cfg, n)
    h, v, o = model.predict(x.cuda(), use_thres=True, thres=0.5)
    assert (h.cpu().numpy() == gold["pred_h"]).mean() >= 0.999
    np.testing.assert_allclose(v.cpu().numpy(), gold["pred_v"], rtol=1e-4, atol=3e-5)
    np.testing.assert_allclose(o.cpu().numpy(), gold["pred_o"], rtol=1e-4, atol=3e-5)


def test_large_batch_step_runs_on_the_split_gemm():
    """A batch large enough that every contraction is pre-imaged and split (the kernel class counters say which kernels ran)
    agrees with precision = 'fp32' on the same weights and inputs."""
    cfg = G.GrooveCfg(256, 4, 512, 2, 0, 16, 27)
    x, y = G.det_batch(cfg, 256)
    out = {}
    for prec in ("fp32", "fp32_tc"):
        model, _ = build_model(cfg, dropout=0.0, precision=prec)
        model.train()
        metrics, hvo = model.train_step(x.cuda(), y.cuda(), 0.7)
        out[prec] = (metrics.cpu().numpy().astype(np.float64), hvo.cpu().numpy(), grads_by_name(model))
    np.testing.assert_allclose(out["fp32_tc"][0], out["fp32"][0], rtol=LOSS_RTOL)
    np.testing.assert_allclose(out["fp32_tc"][1], out["fp32"][1], rtol=1e-4, atol=2e-5)
    _check_grads(out["fp32_tc"][2], out["fp32"][2])
