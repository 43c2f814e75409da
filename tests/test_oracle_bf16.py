"""CPU: the bf16-operand oracle (oracle/groove_oracle_bf16.py) is pinned to the fp32 oracle — which is itself pinned to the
reference-generated golden vectors (tests/test_oracle_golden.py).  With its rounding function replaced by the identity the
hand-written backward passes (rounded linear, mma-style attention core, bf16-statistics LayerNorm, edge256 stem / tail) must
reproduce autograd over the fp32 restatement to fp32 noise; with rounding on, it must stay within bf16 distance of it."""
import pytest
import torch

import groove_oracle as G
import groove_oracle_bf16 as B

CASES = {
    "c2_l2_fused_d32": (G.GrooveCfg(32, 16, 512, 2, 0, 16, 27), 5, 0.24),
    "c1_l2_fused_d32": (G.GrooveCfg(32, 4, 16, 2, 0, 16, 27), 5, 0.18),
    "h1_simt_attention": (G.GrooveCfg(32, 1, 96, 2, 0, 16, 27), 5, 0.1),
    "c5_encdec_l1": (G.GrooveCfg(32, 16, 64, 1, 1, 27, 27), 5, 0.1),
    "c4_l2_fused_d256": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 4, 0.15),
    "c3_l1_fused_d256_h128": (G.GrooveCfg(256, 2, 128, 1, 0, 16, 27), 4, 0.3),   # head dim 128: attention variant 'h'
    "d256_h4_per_op": (G.GrooveCfg(256, 4, 128, 1, 0, 16, 27), 4, 0.3),          # head dim 64 has no fused kernel
    "d64_per_op": (G.GrooveCfg(64, 4, 64, 1, 0, 16, 27), 4, 0.1),
}


def _worst(g1, g0):
    return max(float((g1[k] - g0[k]).abs().max() / (g0[k].abs().max() + 1e-12)) for k in g0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_identity_rounding_reproduces_fp32_oracle(name, monkeypatch):
    cfg, n, p = CASES[name]
    P = G.det_params(cfg)
    x, y = G.det_batch(cfg, n)
    drop = G.DropCtx(p, 7, 1, 0, True)
    l0, g0, _ = G.train_step_oracle(P, cfg, x, y, 0.5, drop)
    monkeypatch.setattr(B, "bf16", lambda t: t)
    l1, g1, _ = B.train_step_oracle_b(P, cfg, x, y, 0.5, drop)
    assert abs(l1[0] - l0[0]) <= 1e-6 * abs(l0[0])
    assert _worst(g1, g0) < 5e-6


@pytest.mark.parametrize("name", sorted(CASES))
def test_bf16_rounding_stays_within_bf16_distance(name):
    cfg, n, p = CASES[name]
    P = G.det_params(cfg)
    x, y = G.det_batch(cfg, n)
    drop = G.DropCtx(p, 7, 1, 0, True)
    l0, g0, _ = G.train_step_oracle(P, cfg, x, y, 0.5, drop)
    l1, g1, _ = B.train_step_oracle_b(P, cfg, x, y, 0.5, drop)
    assert 0 < abs(l1[0] - l0[0]) <= 2e-3 * abs(l0[0])          # north star's bf16 loss tolerance; and rounding is really on
    assert 1e-3 < _worst(g1, g0) < 0.2


def test_path_selection_mirrors_gt_path_kind():
    import ctypes as C
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    want = {_lib.PATH_FUSED_D32: B.PATH_FUSED_D32, _lib.PATH_FUSED_D256: B.PATH_FUSED_D256, _lib.PATH_GEMM_TC: B.PATH_PER_OP}
    for cfg, _, _ in CASES.values():
        c = _lib.GtConfig(cfg.d_model, cfg.nhead, cfg.dim_ff, cfg.n_enc, cfg.n_dec, cfg.e_src, cfg.e_tgt, _lib.PREC_BF16, 0.0, 0)
        assert want[lib.gt_path_kind(C.byref(c))] == B.path_for(cfg), cfg
