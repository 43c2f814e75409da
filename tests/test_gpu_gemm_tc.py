"""Generic tcgen05 GEMM (gemm_tc.cu) and the precision='bf16' path built on it for every shape the fused layer
kernels do not cover: d_model other than 32 / 256, head_dim 128 at d_model = 256 (InfillingKicksAndSnares_training.yaml)
and the encoder-decoder GrooveTransformer.

Unit level: gt_debug_gemm(tc=1) against torch fp32 on bf16-rounded operands (the kernel's arithmetic: bf16 operands,
fp32 accumulation) for the three operand layouts the model uses, ragged M / N / K, split-K and every epilogue option;
the epilogues are cross-checked element-wise against the fp32 SIMT kernel fed the same bf16-rounded operands.
Model level: bf16-mode tolerances of BASELINE.json north_star — per-step loss within 2e-3 relative of the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import groove_oracle as G
from _util import build_model, grads_by_name, rel_err
from transformergrooveinfilling_b200 import _lib

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-3


_SCRATCH = None


def _bf16r(t):
    return t.bfloat16().float()


def _gemm(tc, a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, flags=0, bias=None, residual=None, ld_res=0, mask=None, ld_mask=0,
          mask_scale=1.0, drop_p=0.0, row0=0, split=0):
    lib = _lib.load()
    p = lambda t: 0 if t is None else t.data_ptr()
    if tc:
        # the operand-image scratch is caller-owned (the model passes carve it out of their workspace); 256 MB covers every
        # unit-test problem, so the pre-imaged bulk-TMA variants of the kernel are the ones exercised here
        global _SCRATCH
        if _SCRATCH is None:
            _SCRATCH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        _lib.check(lib.gt_debug_gemm_scratch(_SCRATCH.data_ptr(), _SCRATCH.numel()), "gt_debug_gemm_scratch")
    _lib.check(lib.gt_debug_gemm(tc, p(a), sam, sak, p(b), sbn, sbk, p(c), ldc, m, n, k, flags, p(bias), p(residual), ld_res,
                                 p(mask), ld_mask, mask_scale, drop_p, 5, 2, 77, row0, split, 0), "gt_debug_gemm")
    torch.cuda.synchronize()


# (M, N, K): token-rows x out-features x in-features of the model's Linear layers, plus ragged cases
LINEAR_SHAPES = [(128, 32, 32), (256, 96, 32), (4 * 32, 768, 256), (7 * 32, 512, 256), (13 * 32, 256, 512), (5 * 32, 64, 64),
                 (9 * 32, 192, 64), (1000, 100, 72), (33, 40, 33), (640, 384, 128),
                 # 16 or more M tiles: the weight operand is pre-built as bf16 images once per GEMM and fetched by bulk TMA
                 (2048, 768, 256), (4099, 100, 72), (2048 + 33, 512, 512), (3000, 40, 33), (64 * 32, 32, 32), (2500, 1024, 192)]


@pytest.mark.parametrize("m,n,k", LINEAR_SHAPES)
def test_linear_form(m, n, k):
    """out[M,N] = X[M,K] W[N,K]^T: both operands K-major."""
    torch.manual_seed(m + n + k)
    x, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    out = torch.full((m, n), float("nan"), device="cuda")
    _gemm(1, x, k, 1, w, k, 1, out, n, m, n, k)
    ref = _bf16r(x) @ _bf16r(w).T
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("m,n,k", LINEAR_SHAPES)
def test_dgrad_form(m, n, k):
    """dX[M,K'] = dY[M,N'] W[N',K']: A K-major, B MN-major (its contraction index is the strided one)."""
    torch.manual_seed(m * 3 + n + k)
    dy, w = torch.randn(m, k, device="cuda"), torch.randn(k, n, device="cuda")      # contraction length k, output width n
    out = torch.full((m, n), float("nan"), device="cuda")
    _gemm(1, dy, k, 1, w, 1, n, out, n, m, n, k)
    ref = _bf16r(dy) @ _bf16r(w)
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("tokens,n_out,n_in,chunk", [(128, 96, 32, 0), (4096, 768, 256, 2048), (4099 * 32, 64, 64, 2048),
                                                      (7 * 32, 512, 256, 64), (1000, 40, 72, 256), (2048, 256, 512, 512)])
def test_wgrad_form_split_k(tokens, n_out, n_in, chunk):
    """dW[N,K] += dY[tokens,N]^T X[tokens,K]: both operands MN-major, contraction over the tokens, split-K with atomics."""
    torch.manual_seed(tokens + n_out)
    dy, x = torch.randn(tokens, n_out, device="cuda"), torch.randn(tokens, n_in, device="cuda")
    dw = torch.ones(n_out, n_in, device="cuda")
    _gemm(1, dy, 1, n_out, x, 1, n_in, dw, n_in, n_out, n_in, tokens, flags=4, split=chunk)
    ref = 1.0 + (_bf16r(dy).double().T @ _bf16r(x).double()).float()
    assert rel_err(dw.cpu().numpy(), ref.cpu().numpy()) < 1e-4


def test_strided_views():
    """operands / results that are column slices of wider buffers (the packed q|k|v and k|v projections)."""
    torch.manual_seed(3)
    m, d = 6 * 32, 64
    qkv = torch.randn(m, 3 * d, device="cuda")
    w = torch.randn(d, d, device="cuda")
    out = torch.zeros(m, 2 * d, device="cuda")
    k_view = qkv[:, d:2 * d]
    _gemm(1, k_view, 3 * d, 1, w, d, 1, out[:, d:], 2 * d, m, d, d)
    ref = _bf16r(k_view) @ _bf16r(w).T
    assert rel_err(out[:, d:].cpu().numpy(), ref.cpu().numpy()) < 2e-5
    assert float(out[:, :d].abs().max()) == 0.0


@pytest.mark.parametrize("m,n,k", [(8 * 32, 512, 256), (5 * 32, 96, 32), (300, 100, 72)])
def test_epilogues_match_simt_kernel(m, n, k):
    """bias + ReLU + dropout, ReLU-mask, residual and accumulate epilogues: the tcgen05 kernel and the fp32 SIMT kernel run
    on the same bf16-rounded operands must agree to fp32 summation-order noise, and keep / drop the same elements."""
    torch.manual_seed(11)
    x, w = _bf16r(torch.randn(m, k, device="cuda")), _bf16r(torch.randn(n, k, device="cuda"))
    bias, res, mask = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda"), torch.randn(m, n, device="cuda")
    cases = [dict(flags=1, bias=bias, drop_p=0.24, row0=96),
             dict(mask=mask, ld_mask=n, mask_scale=1.25),
             dict(residual=res, ld_res=n, bias=bias),
             dict(flags=2)]
    for kw in cases:
        outs = []
        for tc in (0, 1):
            c = torch.full((m, n), 0.5, device="cuda")
            _gemm(tc, x, k, 1, w, k, 1, c, n, m, n, k, **kw)
            outs.append(c.cpu().numpy())
        assert rel_err(outs[1], outs[0]) < 2e-5, kw.keys()
        if kw.get("drop_p"):
            assert ((outs[0] == 0) == (outs[1] == 0)).mean() > 0.9999


SHAPES = {
    # InfillingKicksAndSnares_training.yaml (C3): d_model 256, 2 heads of 128, FFN 512 — 2 of its 6 layers.  Round 2: the fused
    # d_model = 256 kernels cover head dim 128, so C3 left the per-op path; the same shape with 4 heads of 64 still runs per-op
    "c3_l2": (G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 0.73, 0.30),
    "d256_h4_l2": (G.GrooveCfg(256, 4, 512, 2, 0, 16, 27), 0.73, 0.30),
    "d64_h4": (G.GrooveCfg(64, 4, 128, 2, 0, 16, 27), 0.5, 0.1),
    "d128_h8_sym": (G.GrooveCfg(128, 8, 96, 1, 0, 27, 27), 0.7, 0.2),
    # InfillingClosedHH_Symbolic_training.yaml with encoder_only = 0 (C5 encoder-decoder), 2 + 2 of its 6 + 6 layers
    "c5_encdec_l2": (G.GrooveCfg(32, 16, 512, 2, 2, 27, 27), 0.38, 0.24),
    "d64_encdec": (G.GrooveCfg(64, 2, 64, 1, 1, 16, 27), 1.0, 0.1),
    # fused encoder-decoder blocks at the other mma.sync head dims: head dim 4 with a single 16-wide FFN chunk, head dim 8
    "d32_h8_f16_encdec": (G.GrooveCfg(32, 8, 16, 1, 2, 16, 27), 0.6, 0.2),
    "d32_h4_f96_encdec": (G.GrooveCfg(32, 4, 96, 2, 1, 27, 27), 0.9, 0.15),
    # head dim 16: the encoder stack and the decoder FFN blocks are fused, the decoder attention blocks run per-op
    "d32_h2_encdec": (G.GrooveCfg(32, 2, 64, 1, 1, 16, 27), 0.5, 0.1),
}
FUSED_ENCDEC = ("c5_encdec_l2", "d32_h8_f16_encdec", "d32_h4_f96_encdec")


def test_path_kind():
    lib = _lib.load()
    kinds = {}
    for name, (cfg, _, _) in SHAPES.items():
        model, _ = build_model(cfg, precision="bf16")
        kinds[name] = lib.gt_path_kind(C.byref(model._cfg()))
    # the C5 encoder-decoder (d_model 32, head dim 2) runs every block of every layer in the fused kernels (hybrid path of
    # runner.cu: fused encoder stack + three fused blocks per decoder layer); the other shapes run per-op on gemm_tc
    for k in FUSED_ENCDEC:
        assert kinds.pop(k) == _lib.PATH_FUSED_D32
    assert kinds.pop("c3_l2") == _lib.PATH_FUSED_D256
    assert set(kinds.values()) == {_lib.PATH_GEMM_TC}, kinds
    m32, _ = build_model(G.GrooveCfg(32, 4, 16, 1, 0, 16, 27), precision="bf16")
    assert lib.gt_path_kind(C.byref(m32._cfg())) == _lib.PATH_FUSED_D32
    m256, _ = build_model(G.GrooveCfg(256, 16, 64, 1, 0, 16, 27), precision="bf16")
    assert lib.gt_path_kind(C.byref(m256._cfg())) == _lib.PATH_FUSED_D256
    m256.set_precision("fp32")
    assert lib.gt_path_kind(C.byref(m256._cfg())) == _lib.PATH_FP32_SIMT


def _oracle_forward(P, cfg, x, y):
    if cfg.n_dec > 0:
        return G.forward_encdec(P, cfg, x, G.shift_right(y))
    return G.forward_encoder_only(P, cfg, x)


@pytest.mark.parametrize("name", sorted(SHAPES))
@pytest.mark.parametrize("n", [4, 13])
def test_eval_forward(name, n):
    cfg, pen, p = SHAPES[name]
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.eval()
    x, y = G.det_batch(cfg, n)
    with torch.no_grad():
        if cfg.n_dec > 0:
            h, v, o = model(x.cuda(), G.shift_right(y).cuda())
        else:
            h, v, o = model(x.cuda())
    rh, rv, ro = _oracle_forward(P, cfg, x, y)
    assert rel_err(h.cpu().numpy(), rh.numpy()) < 3e-2
    assert np.abs(v.cpu().numpy() - rv.numpy()).max() < 2e-2 and np.abs(o.cpu().numpy() - ro.numpy()).max() < 2e-2


def _worst_grad_err(model, grads):
    gg = grads_by_name(model)
    worst = ("", 0.0)
    for k, w in grads.items():
        scale = float(w.abs().max())
        if scale < 1e-6:
            continue
        e = float((gg[k] - w).abs().max()) / scale
        if e > worst[1]:
            worst = (k, e)
    return worst


@pytest.mark.parametrize("name,n", [("c3_l2", 4), ("c3_l2", 64), ("d256_h4_l2", 4), ("d256_h4_l2", 64), ("d64_h4", 5), ("d64_h4", 67), ("d128_h8_sym", 64),
                                    ("c5_encdec_l2", 6), ("c5_encdec_l2", 64), ("d64_encdec", 64), ("d32_h8_f16_encdec", 64),
                                    ("d32_h8_f16_encdec", 7), ("d32_h4_f96_encdec", 64), ("d32_h2_encdec", 64),
                                    ("c5_encdec_l2", 4 * 148 + 3)])
def test_train_step_matches_oracle(name, n):
    cfg, pen, p = SHAPES[name]
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(7, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    loss6, grads, _ = G.train_step_oracle(P, cfg, x, y, pen, G.DropCtx(p, 7, 1, 0, True))
    got = metrics.cpu().numpy().astype(np.float64)
    assert abs(got[0] - loss6[0]) / abs(loss6[0]) < LOSS_RTOL, (got, loss6)
    worst = _worst_grad_err(model, grads)
    # bf16 operand rounding is independent per sample: a few % of each tensor's max at small n, falling as 1/sqrt(n)
    assert worst[1] < (4e-2 if n >= 64 else 0.2), f"gradient mismatch {worst}"


def test_bf16_vs_fp32_same_masks_large_batch():
    """C3 hyper-parameters (2 layers), 512 sequences, dropout on: the bf16 (fused d_model = 256 kernels, head dim 128) and fp32 (SIMT) paths draw identical masks."""
    cfg, pen, p = SHAPES["c3_l2"]
    x, y = [t.cuda() for t in G.det_batch(cfg, 512)]
    out = {}
    for prec in ("fp32", "bf16"):
        model, _ = build_model(cfg, dropout=p, precision=prec)
        model.set_seed(21, step=3).train()
        m, _ = model.train_step(x, y, pen)
        out[prec] = (m.cpu().numpy().astype(np.float64), model.flat_grad().detach().cpu().numpy().copy())
    assert abs(out["bf16"][0][0] - out["fp32"][0][0]) / abs(out["fp32"][0][0]) < LOSS_RTOL
    gb, gf = out["bf16"][1], out["fp32"][1]
    assert np.abs(gb - gf).max() / np.abs(gf).max() < 2e-2


def test_loss_trajectory_bf16_vs_fp32_encdec():
    from transformergrooveinfilling_b200 import FusedSGD
    cfg, pen, p = SHAPES["c5_encdec_l2"]
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
    np.testing.assert_allclose(traj["bf16"], traj["fp32"], rtol=2 * LOSS_RTOL)
    assert traj["fp32"][-1] < traj["fp32"][0]


def test_predict_encdec_bf16_agrees_with_fp32():
    """KV-cached autoregressive predict on the gemm_tc path against the fp32 path.  Step 0 has no feedback: its hits
    must agree on >= 99.5 % of cells.  Later steps feed the thresholded hits back, so a near-threshold flip (random-init
    logits hover around 0, where bf16 rounding flips ~1 % of cells) changes that sequence's later inputs: >= 97 % overall."""
    cfg, pen, p = SHAPES["d64_encdec"]
    x, _ = G.det_batch(cfg, 64)
    res = {}
    for prec in ("fp32", "bf16"):
        model, _ = build_model(cfg, dropout=p, precision=prec)
        h, v, o = model.predict(x.cuda())
        res[prec] = (h.cpu().numpy(), v.cpu().numpy())
    assert (res["bf16"][0][:, 0] == res["fp32"][0][:, 0]).mean() >= 0.995
    assert (res["bf16"][0] == res["fp32"][0]).mean() > 0.97
    assert np.abs(res["bf16"][1][:, 0] - res["fp32"][1][:, 0]).max() < 2e-2


@pytest.mark.parametrize("name,n", [("c5_encdec_l2", 9), ("d32_h4_f96_encdec", 6)])
def test_autograd_path_equals_fused_step_encdec(name, n):
    """model(x, y_shifted) -> calculate_loss -> loss.backward() against the single-call fused step on the fused encoder-decoder
    path: same layer / block kernels, different tail kernels (the fused step folds calculate_loss into the decoder tail)."""
    from transformergrooveinfilling_b200 import calculate_loss
    cfg, pen, p = SHAPES[name]
    x, y = [t.cuda() for t in G.det_batch(cfg, n)]
    m1, _ = build_model(cfg, dropout=p, precision="bf16")
    m2, _ = build_model(cfg, dropout=p, precision="bf16")
    m1.set_seed(5, step=3, seq0=0).train(); m2.set_seed(5, step=3, seq0=0).train()
    metrics, hvo = m1.train_step(x, y, pen)
    pred = m2(x, G.shift_right(y.cpu()).cuda())
    out = calculate_loss(pred, y, None, None, pen)
    out[0].backward()
    np.testing.assert_allclose(torch.cat(pred, 2).detach().cpu().numpy(), hvo.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.array([out[0].item(), *out[1:]]), metrics.cpu().numpy(), rtol=2e-6)
    from _util import grads_by_name as _g
    g1, g2 = _g(m1), _g(m2)
    for k in g1:
        scale = float(g1[k].abs().max()) + 1e-12
        assert float((g2[k] - g1[k]).abs().max()) / scale < 1e-3, k     # see tests/test_gpu_bf16.py: last-bit dL/dlogits differences
