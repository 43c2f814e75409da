"""bf16 mode against the bf16-OPERAND oracle (oracle/groove_oracle_bf16.py: the reference arithmetic with operands rounded
to bf16 exactly where the kernels round).  Against the fp32 oracle the bf16 kernels can only be held to bf16 distance
(tests/test_gpu_bf16*.py: loss 2e-3, gradients a few % of a tensor's max at small batches); against this one the loss, every
gradient tensor and a 20-step trajectory are held to fp32 re-ordering noise plus the odd rounding flip, so a dropped bias
term or a wrong 1/(1-p) at one site cannot hide.  Also: the north star's hit-agreement tolerance (>= 99.9 % of
(sequence, step, voice) cells, BGT/models/utils.py:59-63 thresholding) on TRAINED weights, over >= 1e5 cells, against the
fp32 oracle — i.e. against the reference's own arithmetic."""
import numpy as np
import pytest
import torch

import groove_oracle as G
import groove_oracle_bf16 as B
from _util import build_model, grads_by_name, params_by_name
from transformergrooveinfilling_b200 import FusedAdam, FusedSGD

pytestmark = pytest.mark.gpu

GRAD_TOL = 5e-3        # of each tensor's max |gradient| AND of its L2 norm   (was 0.2 / 4e-2 against the fp32 oracle)
LOSS_TOL = 2e-4        # relative, single step                               (north star for bf16 vs the reference: 2e-3)

# What the tolerance can and cannot be (measured: tools/diag_bf16_exact.py, tools/diag_bf16_grid.py; numbers in DESIGN.md §2).
# With ONE layer the kernels reproduce the oracle to 4e-5 .. 2.5e-3 of a tensor's max on every path: every rounding point is
# modelled.  With more layers a second effect appears that no oracle can remove: the kernels and the oracle add the same fp32
# numbers in different orders, so a value that lies within ~1e-7 of a bf16 rounding boundary is occasionally rounded the other
# way ("flip": 0.4 % of that one element).  A flip in layer l perturbs that token's row by ~1e-4, which makes further flips in
# layer l + 1 a thousand times more likely for the whole sequence: single sequences drift apart at bf16-noise level while all
# others stay exact (forward outputs: mean |diff| 5e-5, max 6e-3).  The drift is per sequence, so its share of a gradient
# falls as 1/n: the full-depth configurations are held to 5e-3 at batch >= 128 and to a documented looser bound at batch 4..9.
CASES = {
    # name: (cfg, hit_loss_penalty, dropout, batch, tolerance)
    # ---- one layer: every rounding point of every path, held to GRAD_TOL at any batch
    "c1_l1": (G.GrooveCfg(32, 4, 16, 1, 0, 16, 27), 0.47, 0.18, 16, GRAD_TOL),           # fused d32, head dim 8 (mma.sync attention)
    "c2_l1": (G.GrooveCfg(32, 16, 512, 1, 0, 16, 27), 0.38, 0.24, 16, GRAD_TOL),         # fused d32, head dim 2, 4 FFN chunks
    "c2_l1_n3": (G.GrooveCfg(32, 16, 512, 1, 0, 16, 27), 0.38, 0.24, 3, GRAD_TOL),       # ragged last tile
    "c5enc_l1": (G.GrooveCfg(32, 16, 512, 1, 0, 27, 27), 0.38, 0.24, 9, GRAD_TOL),       # symbolic input
    "f96_h1": (G.GrooveCfg(32, 1, 96, 2, 0, 16, 27), 0.5, 0.1, 6, GRAD_TOL),             # fused d32 with fp32 (SIMT) attention
    "c5_encdec_l1": (G.GrooveCfg(32, 16, 512, 1, 1, 27, 27), 0.38, 0.24, 64, GRAD_TOL),  # decoder blocks: causal / cross / FFN (4 blocks deep)
    "c5_dec_h4_l1": (G.GrooveCfg(32, 4, 64, 1, 1, 27, 27), 0.38, 0.1, 16, GRAD_TOL),
    "c4_l1": (G.GrooveCfg(256, 16, 64, 1, 0, 16, 27), 1.0, 0.15, 16, GRAD_TOL),          # fused d256, head dim 16
    "h8_l1": (G.GrooveCfg(256, 8, 128, 1, 0, 16, 27), 0.5, 0.1, 16, GRAD_TOL),           # fused d256, head dim 32, two FFN chunks
    # C3 shape (fused d256, head dim 128: a head spans two 64-column groups).  512 hidden units x 2048 tokens = 1 M ReLU inputs: a handful
    # lie within fp32 summation-order noise of 0 and get the other side's mask (tools/diag_c3b.py: the whole error sits in 3 - 5 hidden
    # units, with or without dropout, at head dim 16 as well); one such flip is 1 / n of that unit's gradient: 5.5e-3 at n = 16, 1.3e-3 here
    "c3_l1": (G.GrooveCfg(256, 2, 512, 1, 0, 16, 27), 0.73, 0.3, 64, GRAD_TOL),
    "d64_per_op": (G.GrooveCfg(64, 4, 64, 1, 0, 16, 27), 1.0, 0.1, 8, GRAD_TOL),         # per-op path: gemm_tc + attn_mma
    # ---- full depth at batch >= 128: GRAD_TOL
    "c1_n256": (G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 0.47, 0.18, 256, GRAD_TOL),
    "c2_n256": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24, 256, GRAD_TOL),
    "c5_encdec_n128": (G.GrooveCfg(32, 16, 512, 2, 2, 27, 27), 0.38, 0.24, 128, GRAD_TOL),
    "c4_l2_n128": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15, 128, GRAD_TOL),
    "c3_l2_n128": (G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 0.73, 0.3, 128, GRAD_TOL),
    # ---- full depth at the small batches of the other parity tests: per-sequence drift not averaged out (see above)
    "c1_n5": (G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 0.47, 0.18, 5, 4e-2),
    "c2_n4": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24, 4, 8e-2),
    "c2_p0_n16": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.0, 16, 2e-2),
    "c4_l2_n4": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15, 4, 1e-2),
}


def _worst(gg, grads, l2=False):
    w = ("", 0.0)
    for k, v in grads.items():
        s = float(v.norm()) if l2 else float(v.abs().max())
        if s < 1e-6:
            continue
        d = gg[k] - v
        e = (float(d.norm()) if l2 else float(d.abs().max())) / s
        if e > w[1]:
            w = (k, e)
    return w


@pytest.mark.parametrize("name", sorted(CASES))
def test_train_step_matches_bf16_oracle(name):
    cfg, pen, p, n, tol = CASES[name]
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(7, step=1, seq0=0).train()
    x, y = G.det_batch(cfg, n)
    metrics, hvo = model.train_step(x.cuda(), y.cuda(), pen)
    drop = G.DropCtx(p, 7, 1, 0, True)
    loss6, grads, pred = B.train_step_oracle_b(P, cfg, x, y, pen, drop)
    got = metrics.cpu().numpy().astype(np.float64)
    assert abs(got[0] - loss6[0]) / abs(loss6[0]) < LOSS_TOL, (got, loss6)
    np.testing.assert_allclose(got[1:], np.array(loss6[1:]), rtol=2e-3, atol=1e-5)
    dh = (hvo.cpu() - torch.cat(pred, 2)).abs()
    assert float(dh.mean()) < (1e-3 if tol == GRAD_TOL else 5e-3) and float(dh.max()) < 5e-2      # drifted sequences are sparse
    gg = grads_by_name(model)
    worst, worst_l2 = _worst(gg, grads), _worst(gg, grads, l2=True)
    assert worst[1] < tol, f"gradient mismatch (max norm) {worst}"
    assert worst_l2[1] < tol, f"gradient mismatch (L2) {worst_l2}"
    # the bf16 oracle is a STRICTLY better predictor of the kernels than the fp32 oracle (rounding is modelled, not just bounded)
    _, g32, _ = G.train_step_oracle(P, cfg, x, y, pen, drop)
    assert worst_l2[1] < 0.6 * _worst(gg, g32, l2=True)[1]


TRAJ = {"c2": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38, 0.24), "c4_l2": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0, 0.15)}


@pytest.mark.parametrize("name,batch,opt_name,lr", [("c2", 32, "sgd", 0.07), ("c4_l2", 16, "adam", 1e-3)])
def test_loss_trajectory_20_steps_vs_bf16_oracle(name, batch, opt_name, lr):
    """per-step loss over >= 20 steps from identical weights / inputs.  The oracle runs the SAME 20 steps on the CPU
    (its own gradients, its own optimizer restatement) with the kernels' dropout masks; the two trajectories must agree within
    the bf16 tolerance at every step without ever being re-synchronised."""
    cfg, pen, p = TRAJ[name]
    x, y = G.det_batch(cfg, batch)
    model, P = build_model(cfg, dropout=p, precision="bf16")
    model.set_seed(3).train()
    opt = FusedSGD(model, lr) if opt_name == "sgd" else FusedAdam(model, lr)
    xg, yg = x.cuda(), y.cuda()
    got = []
    for _ in range(20):
        m, _ = model.train_step(xg, yg, pen)
        opt.step()
        got.append(float(m[0]))
    Pm = {k: v.clone() for k, v in P.items()}
    mom = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    want = []
    for t in range(20):
        loss6, grads, _ = B.train_step_oracle_b(Pm, cfg, x, y, pen, G.DropCtx(p, 3, t, 0, True))
        want.append(loss6[0])
        for k in Pm:
            if opt_name == "sgd":
                Pm[k] = G.sgd_step(Pm[k], grads[k], lr)
            else:
                Pm[k], m1, v1 = G.adam_step(Pm[k], grads[k], mom[k][0], mom[k][1], t + 1, lr)
                mom[k] = (m1, v1)
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=1e-3)
    assert want[-1] < want[0]
    # ... and both track the fp32 restatement of the reference within the north star's bf16 tolerance at step 0 and the END
    l0, _, _ = G.train_step_oracle(P, cfg, x, y, pen, G.DropCtx(p, 3, 0, 0, True))
    assert abs(got[0] - l0[0]) / l0[0] < 2e-3


def _learnable_batch(cfg, n, seed):
    """An infilling-like synthetic set the model can actually learn (the SURVEY §8d set draws hits independently of the
    input, so a trained model only learns the prior): input strengths are 0 or in [0.5, 1] (a clear gap, like real MSO
    onsets), and voice k hits where input band k mod 8 is active; velocities / offsets follow the input's."""
    g = torch.Generator().manual_seed(seed)
    half = cfg.e_src // 2 if cfg.e_src != 27 else 9
    m = (torch.rand(n, 32, half, generator=g) < 0.3).float()
    strength = (0.5 + 0.5 * torch.rand(n, 32, half, generator=g)) * m
    timing = (torch.rand(n, 32, half, generator=g) - 0.5) * m
    if cfg.e_src == 27:
        x = torch.cat((m, strength, timing), 2)
    else:
        x = torch.cat((strength, timing), 2)
    idx = torch.arange(9) % half
    y = torch.cat((m[..., idx], strength[..., idx], timing[..., idx]), 2)
    return x.contiguous(), y.contiguous()


HITS = {"c1": (G.GrooveCfg(32, 4, 16, 6, 0, 16, 27), 0.47), "c2": (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), 0.38),
        "c4_l2": (G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), 1.0), "c3_l2": (G.GrooveCfg(256, 2, 512, 2, 0, 16, 27), 0.73)}


@pytest.mark.parametrize("name,data", [("c1", "learnable"), ("c2", "survey"), ("c2", "learnable"), ("c4_l2", "survey"),
                                       ("c4_l2", "learnable"), ("c3_l2", "learnable")])
def test_hit_agreement_after_training(name, data):
    """north star: thresholded hit predictions agree with the reference on >= 99.9 % of cells.  Random-init logits sit at the
    threshold (|logit| ~ 0.3), where ANY change of precision flips cells; the tolerance is about a model that has learnt
    something, so: train in fp32 on the GPU until the logits are decisive, then compare predict() of the bf16 mode on those
    weights with the fp32 oracle's predict (the reference's arithmetic) over 512 sequences = 147 456 cells."""
    cfg, pen = HITS[name]
    model, _ = build_model(cfg, dropout=0.0, precision="fp32")
    model.train()
    opt = FusedAdam(model, 1e-3)
    mk = (lambda n, s: _learnable_batch(cfg, n, s)) if data == "learnable" else (lambda n, s: G.det_batch(cfg, n, tag=s))
    for step in range(3000 if data == "learnable" else 300):
        xb, yb = mk(256, 100 + step % 8)
        m, _ = model.train_step(xb.cuda(), yb.cuda(), pen)
        opt.step()
        if data == "learnable" and step % 50 == 49 and float(m[1]) > 0.9985:
            break
    P = params_by_name(model)
    x, y = mk(512, 999)
    oh, ov, oo = G.predict_encoder_only(P, cfg, x)
    if data == "learnable":
        truth = y[..., :9].to(torch.int64)
        assert float((oh == truth).float().mean()) > 0.99, "the fp32 training run did not learn the mapping (test set-up, not parity)"
        assert 0.05 < float(oh.float().mean()) < 0.6               # both classes are predicted
    model.set_precision("bf16")
    h, v, o = model.predict(x.cuda())
    agree = float((h.cpu() == oh).float().mean())
    assert h.numel() >= 100_000
    assert agree >= 0.999, f"bf16 hits agree with the fp32 oracle on {agree:.5f} of {h.numel()} cells"
    assert float((v.cpu() - ov).abs().max()) < 0.15 and float((o.cpu() - oo).abs().max()) < 0.15
    # and the bf16 oracle's hits agree with the kernel's on (essentially) every cell
    bh, _, _ = B.predict_encoder_only_b(P, cfg, x)
    assert float((h.cpu() == bh).float().mean()) >= 0.9995


def test_eval_mode_autograd_matches_oracle():
    """model.eval(); loss.backward() — the reference supports it (nn.Dropout is the identity, autograd still runs).  The library
    saves activations only in its training plan, so this case runs that plan with p = 0 (it used to run the
    inference plan and hand garbage to backward)."""
    from transformergrooveinfilling_b200 import calculate_loss
    for prec, tol in (("fp32", 2e-4), ("bf16", GRAD_TOL)):
        cfg = G.GrooveCfg(32, 4, 16, 2 if prec == "fp32" else 1, 0, 16, 27)
        model, P = build_model(cfg, dropout=0.3, precision=prec)
        model.eval()
        x, y = G.det_batch(cfg, 6)
        pred = model(x.cuda())
        out = calculate_loss(pred, y.cuda(), None, None, 0.6)
        out[0].backward()
        oracle = B.train_step_oracle_b if prec == "bf16" else G.train_step_oracle
        loss6, grads, _ = oracle(P, cfg, x, y, 0.6, G.DropCtx(0.0, 0, 0, 0, False))
        assert abs(out[0].item() - loss6[0]) / loss6[0] < (1e-4 if prec == "fp32" else LOSS_TOL)
        worst = _worst(grads_by_name(model), grads)
        assert worst[1] < tol, (prec, worst)
        assert not model.training


def test_d512_single_head_trains():
    """d_model = 512 with ONE head of 512 (inside the reference's sweep ranges, configs/*_sweep.yaml): the SIMT attention
    backward tiles the head's columns, so this shape trains in both precisions (it needed 271 KB of shared memory)."""
    cfg = G.GrooveCfg(512, 1, 32, 1, 0, 16, 27)
    x, y = G.det_batch(cfg, 3)
    for prec, ltol, gtol in (("fp32", 1e-4, 5e-4), ("bf16", 2e-3, 0.2)):
        model, P = build_model(cfg, dropout=0.1, precision=prec)
        model.set_seed(5, step=2, seq0=0).train()
        metrics, _ = model.train_step(x.cuda(), y.cuda(), 0.8)
        loss6, grads, _ = G.train_step_oracle(P, cfg, x, y, 0.8, G.DropCtx(0.1, 5, 2, 0, True))
        assert abs(float(metrics[0]) - loss6[0]) / loss6[0] < ltol
        assert _worst(grads_by_name(model), grads)[1] < gtol
