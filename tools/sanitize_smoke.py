"""Smoke-sized steps for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): one fused train step + optimizer +
predict on C1 (d_model 32, 2 layers), C4-l2 (d_model 256 fused), C3-l1 (per-op gemm_tc + attn_mma) and C5 encoder-decoder
(1 + 1 layers), in both precisions.  Usage (on the GPU box):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py [case ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import groove_oracle as G
from _util import build_model
from transformergrooveinfilling_b200 import FusedAdam

CASES = {
    "c1": (G.GrooveCfg(32, 4, 16, 2, 0, 16, 27), 0.18, 6),
    "c2": (G.GrooveCfg(32, 16, 512, 1, 0, 16, 27), 0.24, 5),
    "c4": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 0.15, 5),
    "c3": (G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.3, 5),
    "c5": (G.GrooveCfg(32, 16, 512, 1, 1, 27, 27), 0.24, 5),
}
want = sys.argv[1:] or list(CASES)
precs = os.environ.get("SAN_PREC", "fp32,bf16").split(",")
for name in want:
    cfg, p, n = CASES[name]
    for prec in precs:
        model, P = build_model(cfg, dropout=p, precision=prec)
        model.set_seed(11, 0, 0).train()
        opt = FusedAdam(model, 1e-3)
        x, y = G.det_batch(cfg, n)
        for _ in range(2):
            metrics, _ = model.train_step(x.cuda(), y.cuda(), 0.5)
            opt.step()
        out = model.predict(x.cuda())
        torch.cuda.synchronize()
        print(f"sanitize_smoke {name} {prec}: loss {float(metrics[0]):.5f} ok", flush=True)
