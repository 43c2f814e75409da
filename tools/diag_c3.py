"""GPU diagnostic for the head_dim 128 units of the fused d_model = 256 kernels: per-tensor gradient error against the bf16 oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import groove_oracle as G
import groove_oracle_bf16 as B
from _util import build_model, grads_by_name

CASES = [
    ("c3_l1", G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.3, 16, 7),
    ("c3_l1", G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.3, 16, 8),
    ("c3_l1_p0", G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.0, 16, 7),
    ("c3_l1_n64", G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.3, 64, 7),
    ("h16_f512_l1", G.GrooveCfg(256, 16, 512, 1, 0, 16, 27), 0.73, 0.3, 16, 7),
    ("h2_f64", G.GrooveCfg(256, 2, 64, 1, 0, 16, 27), 0.73, 0.3, 16, 7),
]
for name, cfg, pen, p, n, seed in CASES:
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(seed, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    gg = grads_by_name(model)
    drop = G.DropCtx(p, seed, 1, 0, True)
    l1, g1, pr = B.train_step_oracle_b(P, cfg, x, y, pen, drop)
    print(f"== {name} n={n} seed={seed} loss rel {abs(float(metrics[0]) - l1[0]) / l1[0]:.2e} hvo max {float((hvo.cpu() - torch.cat(pr, 2)).abs().max()):.2e}")
    for k, v in g1.items():
        s = float(v.abs().max())
        if s < 1e-6:
            continue
        d = gg[k] - v
        print(f"   {k.replace('Encoder.Encoder.layers.', 'L'):40s} max {float(d.abs().max()) / s:.2e}  l2 {float(d.norm()) / float(v.norm()):.2e}  (|g|max {s:.2e})")
