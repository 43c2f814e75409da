"""GPU diagnostic for test_hit_agreement_after_training[c2-learnable]: distribution of the logit differences."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import groove_oracle as G
import groove_oracle_bf16 as B
from _util import build_model, params_by_name
from test_gpu_bf16_exact import CASES, _learnable_batch
from transformergrooveinfilling_b200 import FusedAdam

for name in ("c2", "c1"):
    cfg, pen, p, _ = CASES[name]
    model, _ = build_model(cfg, dropout=0.0, precision="fp32")
    model.train()
    opt = FusedAdam(model, 3e-3)
    for step in range(4000):
        xb, yb = _learnable_batch(cfg, 256, 100 + step % 8)
        m, _ = model.train_step(xb.cuda(), yb.cuda(), pen)
        opt.step()
        if step % 250 == 0:
            print("   step", step, "loss", float(m[0]), "acc", float(m[1]))
        if step % 50 == 0 and float(m[1]) > 0.998:
            break
    print(name, "final train loss", float(m[0]), "acc", float(m[1]))
    P = params_by_name(model)
    x, y = _learnable_batch(cfg, 512, 999)
    oh, ov, oo = G.forward_encoder_only(P, cfg, x)
    bh, _, _ = B.forward_b(P, cfg, x)
    model.eval()
    with torch.no_grad():
        f32h = model(x.cuda())[0].cpu()
        model.set_precision("bf16")
        kh = model(x.cuda())[0].cpu()
    for nm, a, b in (("kernel-fp32 vs oracle", f32h, oh), ("kernel-bf16 vs oracle", kh, oh), ("kernel-bf16 vs bf16-oracle", kh, bh.detach()), ("bf16-oracle vs oracle", bh.detach(), oh)):
        d = (a - b).abs()
        print(f"  {nm:28s} max|dlogit| {float(d.max()):.3e} mean {float(d.mean()):.3e}  sign agree {float(((a > 0) == (b > 0)).float().mean()):.5f}")
    print("  |logit|<0.1 frac", float((oh.abs() < 0.1).float().mean()), " |logit|<0.5", float((oh.abs() < 0.5).float().mean()), "logit absmax", float(oh.abs().max()))
