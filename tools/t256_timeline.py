"""GPU: clock64 phase time line of CTA 0 of the d_model = 256 layer kernels (developer build: GROOVE_B200_DEV_FLAGS=-DGT_T256_TIMELINE,
GT_T256_DBG=<launches to trace>).  usage: t256_timeline.py c3|c4 [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import groove_oracle as G
from _util import build_model

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg, pen, p = {"c3": (G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 0.73, 0.3), "c4": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15)}[which]
model, P = build_model(cfg, dropout=p, precision="bf16")
model.set_seed(7, step=1, seq0=0).train()
x = torch.rand(n, 32, 16, device="cuda")
y = (torch.rand(n, 32, 27, device="cuda") < 0.2).float()
for _ in range(3):
    model.train_step(x, y, pen)
torch.cuda.synchronize()
