"""GPU diagnostic: gradient / loss error of the bf16 kernels against the fp32 oracle and against the bf16-operand oracle
(oracle/groove_oracle_bf16.py), per configuration — the numbers the tolerances in tests/test_gpu_bf16_exact.py come from."""
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
    ("c1", G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 0.47, 0.18, 5),
    ("c1", G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 0.47, 0.18, 64),
    ("c2", G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24, 4),
    ("c2", G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24, 64),
    ("c2_p0", G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.0, 16),
    ("f96_h1", G.GrooveCfg(32, 1, 96, 2, 0, 16, 27), 0.5, 0.1, 6),
    ("c5_encdec", G.GrooveCfg(32, 16, 512, 2, 2, 27, 27), 0.38, 0.24, 9),
    ("c4_l2", G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15, 4),
    ("c4_l2", G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15, 64),
    ("h8_f128", G.GrooveCfg(256, 8, 128, 1, 0, 16, 27), 0.5, 0.1, 5),
    ("c3_l2", G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 0.73, 0.3, 8),
    ("d64", G.GrooveCfg(64, 4, 64, 1, 0, 16, 27), 1.0, 0.1, 8),
    ("c2", G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24, 256),
    ("c4_l2", G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15, 256),
    ("c3_l2", G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 0.73, 0.3, 64),
    ("c5_encdec", G.GrooveCfg(32, 16, 512, 2, 2, 27, 27), 0.38, 0.24, 64),
    ("c2_l1", G.GrooveCfg(32, 16, 512, 1, 0, 16, 27), 0.38, 0.24, 16),
    ("c1_l1", G.GrooveCfg(32, 4, 16, 1, 0, 16, 27), 0.47, 0.18, 16),
    ("c4_l1", G.GrooveCfg(256, 16, 64, 1, 0, 16, 27), 1.0, 0.15, 16),
    ("h8_l1", G.GrooveCfg(256, 8, 128, 1, 0, 16, 27), 0.5, 0.1, 16),
    ("c3_l1", G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.3, 16),
    ("c5_encdec_l1", G.GrooveCfg(32, 16, 512, 1, 1, 27, 27), 0.38, 0.24, 16),
    ("c5_dec_h4_l1", G.GrooveCfg(32, 4, 64, 1, 1, 27, 27), 0.38, 0.1, 16),
]


def worst(gg, grads):
    w = ("", 0.0)
    for k, v in grads.items():
        s = float(v.abs().max())
        if s < 1e-6:
            continue
        e = float((gg[k] - v).abs().max()) / s
        if e > w[1]:
            w = (k, e)
    return w


def l2(gg, grads):
    w = ("", 0.0)
    for k, v in grads.items():
        s = float(v.norm())
        if s < 1e-6:
            continue
        e = float((gg[k] - v).norm()) / s
        if e > w[1]:
            w = (k, e)
    return w


for name, cfg, pen, p, n in CASES:
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(7, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    got = metrics.cpu().numpy().astype(np.float64)
    gg = grads_by_name(model)
    drop = G.DropCtx(p, 7, 1, 0, True)
    l0, g0, _ = G.train_step_oracle(P, cfg, x, y, pen, drop)
    l1, g1, pr = B.train_step_oracle_b(P, cfg, x, y, pen, drop)
    hv = float((hvo.cpu() - torch.cat(pr, 2)).abs().max())
    print(f"{name:10s} n={n:3d} {B.path_for(cfg):10s} loss rel: fp32-oracle {abs(got[0]-l0[0])/l0[0]:.2e}  bf16-oracle {abs(got[0]-l1[0])/l1[0]:.2e}"
          f" | hvo maxabs vs bf16-oracle {hv:.2e} | grad worst: fp32-oracle {worst(gg, g0)[1]:.3e}  bf16-oracle {worst(gg, g1)[1]:.3e} ({worst(gg, g1)[0]}) | grad L2rel worst: fp32-oracle {l2(gg, g0)[1]:.3e} bf16-oracle {l2(gg, g1)[1]:.3e} ({l2(gg, g1)[0]})", flush=True)
