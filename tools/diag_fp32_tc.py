"""precision = fp32_tc against the fp32 oracle: per-parameter gradient errors (max norm and L2) of every golden case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import groove_oracle as G
from golden_cases import CASES
from _util import build_model, grads_by_name

for name in sorted(CASES):
    cfg, n, pen, lr = CASES[name]
    x, y = G.det_batch(cfg, n)
    _, want, _ = G.train_step_oracle(G.det_params(cfg), cfg, x, y, pen, G.DropCtx(0.0))
    for prec in ("fp32", "fp32_tc"):
        model, P = build_model(cfg, dropout=0.0, precision=prec)
        model.train()
        model.train_step(x.cuda(), y.cuda(), pen)
        got = grads_by_name(model)
        rows = []
        for k, w in want.items():
            s = float(w.abs().max())
            if s < 1e-7:
                continue
            d = (got[k] - w)
            rows.append((float(d.abs().max()) / s, float(d.norm() / (w.norm() + 1e-30)), k, int((d.abs() > 1e-4 * s).sum()), d.numel()))
        rows.sort(reverse=True)
        print(name, prec, "worst:", " | ".join(f"{k} max {a:.2e} l2 {b:.2e} n>{c}/{t}" for a, b, k, c, t in rows[:3]))
