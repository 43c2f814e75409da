"""Parity at bench.py's FULL sizes through size-independent properties (the oracle finishes only small batches in seconds):

* linearity of the batch mean: the loss / gradient of a full batch equals the mean of the losses / gradients of its two halves
  (the second half run with seq0 = N / 2, so both draw the dropout masks of the full run) — the data-parallel property at size;
* row independence: rows of the full-batch eval forward equal the same rows run alone (tile-aligned subsets: bit-identical
  bf16 roundings) and match the ORACLE run on those rows;
* internal consistency of the six metrics; determinism of predict()."""
import numpy as np
import pytest
import torch

import groove_oracle as G
from _util import build_model

pytestmark = pytest.mark.gpu

# (config, hit_loss_penalty, dropout, full per-GPU batch of bench.py)
FULL = {
    "c2": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24, 32768),      # InfillingClosedHH_training.yaml (headline)
    "c3": (G.GrooveCfg(256, 2, 512, 6, 0, 16, 27), 0.73, 0.3, 8192),        # InfillingKicksAndSnares_training.yaml (head dim 128)
    "c4": (G.GrooveCfg(256, 16, 64, 11, 0, 16, 27), 1.0, 0.15, 8192),       # InfillingRandom_test_large.yaml
    "c5_encdec": (G.GrooveCfg(32, 16, 512, 6, 6, 27, 27), 0.38, 0.24, 16384),  # InfillingClosedHH_Symbolic (encoder_only = 0)
}


def _synth(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    hits = (torch.rand(n, 32, 9, generator=g) < 0.15).float()
    y = torch.cat((hits, torch.rand(n, 32, 9, generator=g) * hits, (torch.rand(n, 32, 9, generator=g) - 0.5) * hits), 2)
    if cfg.e_src == 27:
        h2 = (torch.rand(n, 32, 9, generator=g) < 0.15).float()
        x = torch.cat((h2, torch.rand(n, 32, 9, generator=g) * h2, (torch.rand(n, 32, 9, generator=g) - 0.5) * h2), 2)
    else:
        m = (torch.rand(n, 32, 8, generator=g) < 0.5).float()
        x = torch.cat((torch.rand(n, 32, 8, generator=g) * m, (torch.rand(n, 32, 8, generator=g) - 0.5) * m), 2)
    return x.contiguous(), y.contiguous()


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_batch_is_the_mean_of_its_halves(name):
    cfg, pen, p, n = FULL[name]
    x, y = [t.cuda() for t in _synth(cfg, n, 7)]
    model, _ = build_model(cfg, dropout=p, precision="bf16")
    model.train()
    model.set_seed(11, step=4, seq0=0)
    m_full, _ = model.train_step(x, y, pen)
    g_full = model.flat_grad().detach().clone()
    m_full = m_full.cpu().numpy().astype(np.float64)
    halves = []
    for k in range(2):
        model.set_seed(11, step=4, seq0=k * n // 2)
        m, _ = model.train_step(x[k * n // 2:(k + 1) * n // 2], y[k * n // 2:(k + 1) * n // 2], pen)
        halves.append((m.cpu().numpy().astype(np.float64), model.flat_grad().detach().clone()))
    m_mean = 0.5 * (halves[0][0] + halves[1][0])
    # loss, accuracy, bce, mse_v, mse_o are batch means (perplexity = exp(bce) is not linear)
    for i in (0, 1, 3, 4, 5):
        assert abs(m_full[i] - m_mean[i]) <= 2e-6 * max(1.0, abs(m_full[i])), (i, m_full, m_mean)
    g_mean = 0.5 * (halves[0][1] + halves[1][1])
    err = float((g_full - g_mean).abs().max()) / float(g_full.abs().max())
    assert err < 2e-4, err                                   # fp32 re-association of the token sums only
    # internal consistency of the six metrics (BGT/models/train.py:9-40)
    assert abs(m_full[0] - (m_full[3] + m_full[4] + m_full[5])) < 1e-5 * m_full[0]
    assert abs(m_full[2] - np.exp(m_full[3])) < 1e-4 * m_full[2] and 0.0 <= m_full[1] <= 1.0


@pytest.mark.parametrize("name", sorted(FULL))
def test_rows_of_the_full_batch_match_solo_runs_and_the_oracle(name):
    cfg, pen, p, n = FULL[name]
    x, y = _synth(cfg, n, 8)
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.eval()
    xc, yc = x.cuda(), y.cuda()
    tgt = G.shift_right(y).cuda() if cfg.n_dec > 0 else None
    with torch.no_grad():
        full = torch.cat(model(xc, tgt) if cfg.n_dec > 0 else model(xc), 2)
    picks = [0, 4 * 37, n // 2 + 8, n - 8]                    # tile-aligned (4 sequences per tile) windows of 8 sequences
    for k0 in picks:
        sl = slice(k0, k0 + 8)
        with torch.no_grad():
            solo = torch.cat(model(xc[sl], tgt[sl]) if cfg.n_dec > 0 else model(xc[sl]), 2)
        assert torch.equal(solo, full[sl]), k0                # same tiles, same kernels: bit-identical
        if cfg.n_dec > 0:
            ref = G.forward_encdec(P, cfg, x[sl], G.shift_right(y[sl]))
        else:
            ref = G.forward_encoder_only(P, cfg, x[sl])
        got = full[sl].cpu().numpy()
        rh = ref[0].numpy()
        # full depth (6 / 11 / 6 + 6 layers): the bf16 operand roundings of every layer add up along the residual stream; the
        # 2-layer parity tests hold 3e-2 of the largest logit, the 11-layer C4 stack measures 3.2e-2
        assert np.abs(got[..., :9] - rh).max() / (np.abs(rh).max() + 1e-12) < 5e-2
        assert np.abs(got[..., 9:18] - ref[1].numpy()).max() < 3e-2 and np.abs(got[..., 18:] - ref[2].numpy()).max() < 3e-2
    # one full-depth training step on a window of the batch against the oracle with the same dropout masks (seq0 = window start):
    # the north star's bf16 bound on the per-step loss
    k0 = picks[2]
    sl = slice(k0, k0 + 8)
    model.train()
    model.set_seed(3, step=1, seq0=k0)
    m, _ = model.train_step(xc[sl], yc[sl], pen)
    loss6, _, _ = G.train_step_oracle(P, cfg, x[sl], y[sl], pen, G.DropCtx(p, 3, 1, k0, True))
    assert abs(float(m[0]) - loss6[0]) / abs(loss6[0]) < 2e-3, (m.cpu().numpy(), loss6)


@pytest.mark.parametrize("name", ["c2", "c5_encdec"])
def test_predict_is_deterministic_at_full_size(name):
    cfg, pen, p, n = FULL[name]
    n = n if cfg.n_dec == 0 else n // 4
    x, _ = _synth(cfg, n, 9)
    model, _ = build_model(cfg, dropout=p, precision="bf16")
    a = model.predict(x.cuda())
    b = model.predict(x.cuda())
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    assert a[0].shape == (n, 32, 9) and float(a[1].min()) >= 0.0 and float(a[1].max()) <= 1.0 and float(a[2].abs().max()) <= 0.5
