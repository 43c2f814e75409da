"""GPU diagnostic: WHERE the linear1 gradient of the fused d_model = 256 path differs from the bf16 oracle (F = 512, dropout on)."""
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

for name, cfg, pen, p, n, seed in [("h16_f512_l1", G.GrooveCfg(256, 16, 512, 1, 0, 16, 27), 0.73, 0.3, 16, 7),
                                   ("h16_f512_l1_p0", G.GrooveCfg(256, 16, 512, 1, 0, 16, 27), 0.73, 0.0, 16, 7),
                                   ("h16_f64_l1", G.GrooveCfg(256, 16, 64, 1, 0, 16, 27), 0.73, 0.3, 16, 7),
                                   ("c3_l1", G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.3, 16, 7)]:
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(seed, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    gg = grads_by_name(model)
    drop = G.DropCtx(p, seed, 1, 0, True)
    l1, g1, pr = B.train_step_oracle_b(P, cfg, x, y, pen, drop)
    kb, kw = "Encoder.Encoder.layers.0.linear1.bias", "Encoder.Encoder.layers.0.linear1.weight"
    db = (gg[kb] - g1[kb]).abs()
    dw = (gg[kw] - g1[kw]).abs()
    print(f"== {name}: |gb|max {float(g1[kb].abs().max()):.3e} |gw|max {float(g1[kw].abs().max()):.3e}")
    top = torch.topk(db, 8)
    print("   bias err top units:", [(int(i), f"{float(v):.2e}", f"g={float(g1[kb][i]):.2e}") for v, i in zip(top.values, top.indices)])
    rowmax = dw.max(1).values
    top = torch.topk(rowmax, 8)
    print("   weight err top rows:", [(int(i), f"{float(v):.2e}") for v, i in zip(top.values, top.indices)])
    colmax = dw.max(0).values
    top = torch.topk(colmax, 8)
    print("   weight err top cols:", [(int(i), f"{float(v):.2e}") for v, i in zip(top.values, top.indices)])
    print(f"   bias err: median {float(db.median()):.2e} mean {float(db.mean()):.2e} max {float(db.max()):.2e}; n units with err > 10 x median: {int((db > 10 * db.median()).sum())}")
    # per 64-unit chunk
    print("   bias err per chunk (max):", [f"{float(db[c * 64:(c + 1) * 64].max()):.1e}" for c in range(cfg.dim_ff // 64)])
