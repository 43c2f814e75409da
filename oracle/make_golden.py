"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python -B oracle/make_golden.py

For each case the reference modules are constructed with the case's hyper-parameters, loaded
(strict) with the deterministic weights of ``groove_oracle.det_params`` and run on the deterministic
batch of ``groove_oracle.det_batch``; only OUTPUTS are stored (forward h/v/o, loss + 5 metrics,
per-parameter gradient digests or full gradients, parameters after an SGD / Adam step, a short loss
trajectory, predict outputs).  Dropout is 0 (a legal constructor argument): torch's Philox stream
cannot be reproduced by any other implementation, so p=0 is where exact parity is defined.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")          # the UNMODIFIED reference; must shadow the repo's alias package
warnings.filterwarnings("ignore")

import groove_oracle as G  # noqa: E402
from golden_cases import CASES, N_TRAJ, digest  # noqa: E402
from BaseGrooveTransformers.models.transformer import GrooveTransformerEncoder, GrooveTransformer  # noqa: E402
from BaseGrooveTransformers.models.train import calculate_loss  # noqa: E402
import BaseGrooveTransformers as _ref  # noqa: E402
assert _ref.__file__.startswith("/root/reference/"), "make_golden must run against the real reference"

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

def build_ref(cfg):
    if cfg.n_dec > 0:
        m = GrooveTransformer(cfg.d_model, cfg.e_src, cfg.e_tgt, cfg.nhead, cfg.dim_ff, 0.0,
                              cfg.n_enc, cfg.n_dec, 32, "cpu")
    else:
        m = GrooveTransformerEncoder(cfg.d_model, cfg.e_src, cfg.e_tgt, cfg.nhead, cfg.dim_ff, 0.0,
                                     cfg.n_enc, 32, "cpu")
    sd = m.state_dict()
    P = G.det_params(cfg)
    for k, v in P.items():
        assert sd[k].shape == v.shape, k
        sd[k] = v.clone()
    m.load_state_dict(sd, strict=True)
    return m, P


def run_case(name, cfg, n, penalty, lr):
    torch.manual_seed(0)
    model, P = build_ref(cfg)
    x, y = G.det_batch(cfg, n)
    bce = torch.nn.BCEWithLogitsLoss(reduction="none")
    mse = torch.nn.MSELoss(reduction="none")
    out = {}

    def fwd(m):
        return m(x, G.shift_right(y)) if cfg.n_dec > 0 else m(x)

    model.train()
    pred = fwd(model)
    res = calculate_loss(pred, y, bce, mse, penalty)
    res[0].backward()
    out["h"], out["v"], out["o"] = [t.detach().numpy() for t in pred]
    out["loss6"] = np.array([float(res[0])] + [float(r) for r in res[1:]], dtype=np.float64)
    names = [k for k, _ in G.param_shapes(cfg)]
    params = dict(model.named_parameters())
    out["grad_digest"] = np.stack([digest(params[k].grad, i) for i, k in enumerate(names)])
    small = sum(p.numel() for p in params.values()) < 60000
    if small:
        for k in names:
            out["grad/" + k] = params[k].grad.numpy().copy()

    # loss trajectories: N_TRAJ (20) SGD steps and N_TRAJ Adam steps from the same start
    for opt_name in ("sgd", "adam"):
        m2, _ = build_ref(cfg)
        m2.train()
        opt = torch.optim.SGD(m2.parameters(), lr=lr) if opt_name == "sgd" else torch.optim.Adam(m2.parameters(), lr=1e-3)
        traj = []
        for _ in range(N_TRAJ):
            opt.zero_grad()
            r = calculate_loss(fwd(m2), y, bce, mse, penalty)
            r[0].backward()
            opt.step()
            traj.append(float(r[0]))
        out[f"traj_{opt_name}"] = np.array(traj)
        p2 = dict(m2.named_parameters())
        out[f"param_digest_{opt_name}"] = np.stack([digest(p2[k], i) for i, k in enumerate(names)])

    # inference
    model.eval()
    with torch.no_grad():
        ph, pv, po = model.predict(x, use_thres=True, thres=0.5)
    out["pred_h"], out["pred_v"], out["pred_o"] = ph.numpy(), pv.numpy(), po.numpy()
    out["pred_h_dtype"] = np.array(str(ph.dtype))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss6", out["loss6"], "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def demo_checkpoint_keys():
    """State-dict key/shape list of the reference's demo checkpoint (strict-load fixture)."""
    ck = torch.load("/root/reference/demo/transformer_run_171tyqit_Epoch_1.Model", map_location="cpu", weights_only=False)
    sd = ck["model_state_dict"]
    with open(os.path.join(OUT, "demo_checkpoint_keys.txt"), "w") as f:
        for k, v in sd.items():
            f.write(f"{k} {' '.join(map(str, v.shape))}\n")
        f.write(f"#optimizer_state_dict {sorted(ck['optimizer_state_dict'].keys())} "
                f"param_groups0_keys {sorted(ck['optimizer_state_dict']['param_groups'][0].keys())}\n")
    print("demo keys:", len(sd))


if __name__ == "__main__":
    only = sys.argv[1:]                            # python -B oracle/make_golden.py [case ...]: regenerate only the named cases
    for nm, (cfg, n, pen, lr) in CASES.items():
        if not only or nm in only:
            run_case(nm, cfg, n, pen, lr)
    if not only:
        demo_checkpoint_keys()
