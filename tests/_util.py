"""Shared helpers for the parity tests (oracle = checker, CUDA library = thing under test)."""
import numpy as np
import torch

import groove_oracle as G
from transformergrooveinfilling_b200 import GrooveTransformer, GrooveTransformerEncoder


def build_model(cfg: G.GrooveCfg, device="cuda", dropout=None, precision="fp32", P=None):
    p = cfg.dropout if dropout is None else dropout
    if cfg.n_dec > 0:
        m = GrooveTransformer(cfg.d_model, cfg.e_src, cfg.e_tgt, cfg.nhead, cfg.dim_ff, p, cfg.n_enc, cfg.n_dec, 32, device)
    else:
        m = GrooveTransformerEncoder(cfg.d_model, cfg.e_src, cfg.e_tgt, cfg.nhead, cfg.dim_ff, p, cfg.n_enc, 32, device)
    P = G.det_params(cfg) if P is None else P
    sd = m.state_dict()
    for k, v in P.items():
        sd[k] = v.clone()
    m.load_state_dict(sd, strict=True)
    m.set_precision(precision)
    return m, P


def grads_by_name(model):
    g = model.flat_grad()
    return {name: g[o:o + s].view(shape).detach().cpu() for (name, shape), (o, s) in zip(model._spec_list, model._offsets)}


def params_by_name(model):
    return {name: p.detach().cpu().clone() for (name, _), p in zip(model._spec_list, model._views)}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))
