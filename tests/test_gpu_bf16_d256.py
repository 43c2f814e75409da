"""d_model = 256 tensor-core path (tc256.cu: weight-streaming fused layer kernels) against the oracle.
bf16-mode tolerances of BASELINE.json north_star: per-step loss within 2e-3 relative."""
import numpy as np
import pytest
import torch

import groove_oracle as G
from _util import build_model, grads_by_name, rel_err

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-3

SHAPES = {
    "c4_l2": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15),       # InfillingRandom_test_large.yaml, 2 of its 11 layers
    "h8_f128": (G.GrooveCfg(256, 8, 128, 1, 0, 16, 27), 0.5, 0.1),      # head_dim 32, two FFN chunks
    "h16_f192_sym": (G.GrooveCfg(256, 16, 192, 1, 0, 27, 27), 0.7, 0.3),
}


@pytest.mark.parametrize("name", sorted(SHAPES))
@pytest.mark.parametrize("n", [4, 7, 13])
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


def test_many_tiles_persistent_loop():
    """more tiles than SMs: every CTA walks several tiles, exercising every mbarrier phase flip"""
    cfg, pen, p = SHAPES["c4_l2"]
    model, P = build_model(cfg, dropout=0.0, precision="bf16")
    model.eval()
    x, y = G.det_batch(cfg, 4 * 148 * 3 + 2)
    with torch.no_grad():
        h, v, o = model(x.cuda())
    rh, rv, ro = G.forward_encoder_only(P, cfg, x)
    assert rel_err(h.cpu().numpy(), rh.numpy()) < 3e-2
