"""GPU diagnostic: where does the bf16 kernel path leave the bf16-operand oracle?  Grid over depth / heads / FFN width at p = 0."""
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


GRID = [(32, 4, 16, 1), (32, 4, 16, 2), (32, 4, 16, 6), (32, 16, 512, 1), (32, 16, 512, 2), (32, 1, 96, 6), (32, 4, 96, 2), (32, 16, 96, 2),
        (32, 1, 512, 2), (32, 1, 16, 2), (32, 8, 128, 2), (32, 16, 16, 2),
        (256, 16, 64, 1), (256, 8, 64, 1), (256, 8, 128, 1), (256, 16, 128, 1), (256, 16, 512, 1), (256, 2, 64, 1), (256, 2, 512, 1), (256, 4, 64, 1)]
P_DROP = float(os.environ.get("DIAG_P", "0"))
for d, H, F, L in GRID:
    cfg = G.GrooveCfg(d, H, F, L, 0, 16, 27)
    n = 8
    model, P = build_model(cfg, dropout=P_DROP, precision="bf16")
    model.set_seed(7, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), 0.5)
    gg = grads_by_name(model)
    drop = G.DropCtx(P_DROP, 7, 1, 0, True)
    l0, g0, p0 = G.train_step_oracle(P, cfg, x, y, 0.5, drop)
    l1, g1, p1 = B.train_step_oracle_b(P, cfg, x, y, 0.5, drop)
    hv0 = float((hvo.cpu() - torch.cat(p0, 2)).abs().max())
    hv1 = float((hvo.cpu() - torch.cat(p1, 2)).abs().max())
    hm1 = float((hvo.cpu() - torch.cat(p1, 2)).abs().mean())
    w0, w1 = worst(gg, g0), worst(gg, g1)
    print(f"d{d} H{H:2d} F{F:3d} L{L} {B.path_for(cfg):10s} hvo max: fp32 {hv0:.2e} bf16o {hv1:.2e} (mean {hm1:.1e}) | grad worst: fp32 {w0[1]:.2e} bf16o {w1[1]:.2e} ratio {w1[1]/w0[1]:.2f} {w1[0].replace('Encoder.Encoder.','')}", flush=True)
