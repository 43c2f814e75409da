"""Golden-fixture case table + digest helper shared by oracle/make_golden.py (generator, needs the
reference) and the tests (consumers, never touch the reference).  Test infrastructure only."""
import numpy as np
import torch

import groove_oracle as G

# name -> (cfg, batch, hit_loss_penalty, lr).  Hyper-parameters follow the yamls named in
# BASELINE.json (C1..C5); layer counts are reduced for the d=256 cases to keep the
# fixtures small — the per-layer arithmetic is identical.
CASES = {
    "c1_closedhh_testing": (G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 5, 0.47, 0.094),
    "c2_closedhh":         (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 4, 0.38, 0.07),
    "c3_kicksnares_l2":    (G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 3, 0.73, 0.004),
    "c4_random_large_l2":  (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 3, 1.0, 0.004),
    "c5_symbolic_encdec":  (G.GrooveCfg(32, 16, 512, 2, 2, 27, 27), 4, 0.38, 0.07),
    "odd_small_encdec":    (G.GrooveCfg(24, 3, 40, 1, 1, 16, 27), 2, 0.5, 0.05),
    # full depth of the two d_model = 256 configurations (outputs, loss, gradient / parameter digests, 20-step trajectories)
    "c3_kicksnares_full":  (G.GrooveCfg(256, 2, 512, 6, 0, 16, 27), 3, 0.73, 0.004),
    "c4_random_large_full": (G.GrooveCfg(256, 16, 64, 11, 0, 16, 27), 3, 1.0, 0.004),
    # embedding_size_tgt other than 27: the reference splits the head's output and y into thirds whatever the voice count
    # (BGT/models/io_layers.py:34-40, train.py:12-13) — 4 voices encoder-only, 5 voices encoder-decoder
    "voices4_enc":         (G.GrooveCfg(32, 4, 64, 2, 0, 16, 12), 4, 0.6, 0.05),
    "voices5_encdec":      (G.GrooveCfg(32, 8, 48, 1, 2, 16, 15), 3, 0.4, 0.05),
}
# The SGD learning rate of the d_model = 256 trajectories is 0.004, not the yamls' 0.089 / 0.04: on a 3-sequence batch the yaml
# rates make the loss jump 7 -> 15 -> 8 -> 13 ..., and float32 and float64 runs of the SAME arithmetic differ by 2 - 10 % after
# eight steps — no two implementations (or BLAS builds) can agree over 20 such steps.  At 0.004 float32 and float64 agree to 1e-6.

# per-step loss compared over >= 20 optimisation steps
N_TRAJ = 20



def digest(t: torch.Tensor, tag: int):
    """(dot with a fixed pseudo-random vector, L2 norm) — compact and discriminating."""
    r = torch.from_numpy(G.det_uniform(900 + tag, t.numel(), -1, 1)).double()
    td = t.detach().double().reshape(-1)
    return np.array([float(td @ r), float(td.norm())])


