"""bf16-OPERAND restatement of the train / inference step — the checker of the bf16 (tcgen05) mode.

TEST INFRASTRUCTURE ONLY (same rules as groove_oracle.py: imported by tests/, smoke() and bench.py's
CPU legs, never by the product).

groove_oracle.py restates the reference in fp32; against it the bf16 kernels can only be held to the
north star's loose bf16 tolerances (loss 2e-3), and their GRADIENTS differ by the accumulated operand
rounding (a few % of a tensor's max at small batches) — a tolerance wide enough to hide a dropped term.
This module closes that gap: it is the same arithmetic (it calls groove_oracle for everything that stays
fp32) with every operand rounded to bf16 (round-to-nearest-even, like ``__floats2bfloat162_rn``) at
EXACTLY the points where the CUDA kernels round, forward and backward, so kernel-vs-oracle gradient
comparisons tighten to fp32 re-ordering noise.  The rounding points were read off the kernels:

  every tensor-core contraction          both operands bf16, fp32 accumulate; the incoming gradient is rounded
  (csrc/tc_layers.cu, tc256*.cu,         ONCE and that image feeds the data gradient AND the weight gradient;
   gemm_tc.cu)                           bias gradients are column sums of the UNROUNDED fp32 gradient
  attention on mma.sync                  q * log2(e)/sqrt(dh), k, v -> bf16; P (after dropout and 1/(1-p)) -> bf16;
  (tc_attn32.cuh, tc256.cu/_bwd.cu,      backward: dO -> bf16, dS/sqrt(dh) -> bf16 (dq), dS*ln2 -> bf16 against the
   attn_mma.cu)                          scaled q (fused kernels) or dS/sqrt(dh) against bf16(q) (per-op kernel)
  d_model = 32 fused path                residual stream, LayerNorm, input layer and head in fp32 (edge32.cu);
  (tc_layers.cu, tc_attn32.cuh)          head dims outside {2, 4, 8}: attention in fp32 (SIMT).  mma attention: the P V
                                         contraction takes E = bf16(exp2(s - max)) with dropped keys zeroed, the row factor
                                         1 / (rowsum (1 - p)) lands on the context; backward: c * dropped-P and c * dS as bf16
                                         (c = 1/sqrt(dh)), dk from (c dS)^T against the scaled q.  FFN: b1 rides in the
                                         contraction as bf16; H = bf16(relu(.)) with dropped units zeroed and NO 1 / (1 - p),
                                         which multiplies the FFN2 accumulator instead (backward: the da2 image); dH is
                                         rounded, then masked by H != 0; the linear1 / in-projection bias gradients are
                                         column sums of the ROUNDED dH / dq | dk | dv images (taken on the tensor cores)
  d_model = 256 fused path               residual stream between layers and the saved pre-LayerNorm sums are bf16
                                         images: LayerNorm backward takes its statistics / x-hat from bf16(u);
                                         stem: dW_in | db_in = bf16(g)^T [bf16(src) | 1]; tail: dW_out, dgamma from
                                         T = bf16(x_L)^T bf16(dlogits * rstd) (edge256.cu header formulas)
  per-op path (gemm_tc + attn_mma)       fp32 activations between the kernels; input layer (K < 32) and the
                                         27-wide head run on the fp32 SIMT GEMM

Reference lines restated are the ones groove_oracle.py cites (BGT/models/*.py, torch/nn/modules/transformer.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

import groove_oracle as G

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453

PATH_FUSED_D32 = "fused_d32"
PATH_FUSED_D256 = "fused_d256"
PATH_PER_OP = "per_op"


def path_for(cfg: G.GrooveCfg) -> str:
    """Which bf16 implementation the library picks (include/groove_b200.h GT_PATH_*, gt_path_kind)."""
    dh = cfg.dh
    if cfg.d_model == 32 and cfg.dim_ff % 16 == 0 and cfg.dim_ff <= 512 and (dh == 1 or dh % 2 == 0):
        if cfg.n_dec == 0 or dh in (2, 4, 8):
            return PATH_FUSED_D32
    if cfg.d_model == 256 and cfg.n_dec == 0 and dh in (16, 32, 128) and cfg.dim_ff % 64 == 0 and 64 <= cfg.dim_ff <= 512:
        return PATH_FUSED_D256
    return PATH_PER_OP


def bf16(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 (nearest even) and return in the input dtype."""
    return x.to(torch.bfloat16).to(x.dtype)


class _RoundSTE(torch.autograd.Function):
    """Forward: bf16 rounding of a stored activation image; backward: identity (the gradient buffers are fp32)."""

    @staticmethod
    def forward(ctx, x):
        return bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


rb = _RoundSTE.apply


class _LinB(torch.autograd.Function):
    """y = bf16(x) bf16(W)^T + b.  Backward: gb = bf16(g) feeds dx = gb bf16(W) and dW = gb^T bf16(x); db = sum(g) in fp32."""

    @staticmethod
    def forward(ctx, x, w, b):
        xb, wb = bf16(x), bf16(w)
        ctx.save_for_backward(xb, wb)
        return xb @ wb.T + b

    @staticmethod
    def backward(ctx, g):
        xb, wb = ctx.saved_tensors
        gb = bf16(g)
        dx = gb @ wb
        dw = gb.reshape(-1, gb.shape[-1]).T @ xb.reshape(-1, xb.shape[-1])
        db = g.reshape(-1, g.shape[-1]).sum(0)
        return dx, dw, db


def lin_b(x, w, b):
    return _LinB.apply(x, w, b)


class _LinBr(torch.autograd.Function):
    """_LinB whose bias gradient is the column sum of the ROUNDED gradient image (fused d_model = 32 kernels: in-projection)."""

    @staticmethod
    def forward(ctx, x, w, b):
        xb, wb = bf16(x), bf16(w)
        ctx.save_for_backward(xb, wb)
        return xb @ wb.T + b

    @staticmethod
    def backward(ctx, g):
        xb, wb = ctx.saved_tensors
        gb = bf16(g)
        g2 = gb.reshape(-1, gb.shape[-1])
        return gb @ wb, g2.T @ xb.reshape(-1, xb.shape[-1]), g2.sum(0)


class _FFN32(torch.autograd.Function):
    """Feed-forward block of the fused d_model = 32 kernels (tc_layers.cu): see the module header."""

    @staticmethod
    def forward(ctx, x1, w1, b1, w2, b2, keep, fs):
        x1b, w1b, w2b = bf16(x1), bf16(w1), bf16(w2)
        h = bf16(torch.relu(x1b @ w1b.T + bf16(b1)))
        if keep is not None:
            h = h * keep
        ctx.save_for_backward(x1b, w1b, w2b, h)
        ctx.fs = fs
        return (h @ w2b.T) * fs + b2

    @staticmethod
    def backward(ctx, g):
        x1b, w1b, w2b, h = ctx.saved_tensors
        d, f = w1b.shape[1], w1b.shape[0]
        ga = bf16(g * ctx.fs)
        dh = bf16(ga @ w2b) * (h != 0).to(g.dtype)
        dh2, ga2 = dh.reshape(-1, f), ga.reshape(-1, d)
        return (dh @ w1b, dh2.T @ x1b.reshape(-1, d), dh2.sum(0), ga2.T @ h.reshape(-1, f), g.reshape(-1, d).sum(0), None, None)


class _AttnCore32(torch.autograd.Function):
    """mma.sync attention of the fused d_model = 32 kernels (tc_attn32.cuh), heads laid out [N, H, T, dh]."""

    @staticmethod
    def forward(ctx, q, k, v, keep, scale, causal):
        dh = q.shape[-1]
        c = np.float32(1.0 / math.sqrt(dh)) * np.float32(LOG2E)
        qs, kb, vb = bf16(q * float(c)), bf16(k), bf16(v)
        s2 = qs @ kb.transpose(-1, -2)
        if causal:
            t = s2.shape[-1]
            s2 = s2.masked_fill(torch.triu(torch.ones(t, t, dtype=torch.bool), diagonal=1), -1e30)
        e = torch.exp2(s2 - s2.max(-1, keepdim=True).values)
        inv = 1.0 / e.sum(-1, keepdim=True)
        eb = bf16(e) if keep is None else bf16(e) * keep.to(e.dtype)
        ctx.save_for_backward(qs, kb, vb, e, inv, keep if keep is not None else torch.empty(0))
        ctx.scale, ctx.has_keep = scale, keep is not None
        return bf16((eb @ vb) * (inv * scale))

    @staticmethod
    def backward(ctx, dctx):
        qs, kb, vb, e, inv, keep = ctx.saved_tensors
        dh = qs.shape[-1]
        c1, rc1 = float(np.float32(1.0 / math.sqrt(dh))), float(np.float32(math.sqrt(dh)))
        dob = bf16(dctx)
        dp = dob @ vb.transpose(-1, -2)
        i = inv * c1
        pd = e * (i * ctx.scale)
        if ctx.has_keep:
            pd = pd * keep.to(pd.dtype)
        t = pd * dp
        delta = t.sum(-1, keepdim=True)
        dsq = bf16(t - e * (delta * rc1 * i))
        pdb = bf16(pd)
        dq = dsq @ kb
        dk = (dsq.transpose(-1, -2) @ qs) * (LN2 * rc1)
        dv = (pdb.transpose(-1, -2) @ dob) * rc1
        return dq, dk, dv, None, None, None


class _AttnCore(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) (dropout) v for heads laid out [N, H, T, dh], with the mma.sync kernels' rounding points.
    variant 'f': fused kernels (tc_attn32.cuh, tc256.cu / tc256_bwd.cu); 'h': the head_dim 128 units of tc256*.cu (forward as
    'f'); 'p': per-op kernel (attn_mma.cu)."""

    @staticmethod
    def forward(ctx, q, k, v, keep, scale, causal, variant):
        dh = q.shape[-1]
        c = np.float32(1.0 / math.sqrt(dh)) * np.float32(LOG2E)
        qs, kb, vb = bf16(q * float(c)), bf16(k), bf16(v)
        s2 = qs @ kb.transpose(-1, -2)
        if causal:
            t = s2.shape[-1]
            s2 = s2.masked_fill(torch.triu(torch.ones(t, t, dtype=torch.bool), diagonal=1), -1e30)
        e = torch.exp2(s2 - s2.max(-1, keepdim=True).values)
        p = e / e.sum(-1, keepdim=True)
        pd = p * scale if keep is None else p * keep.to(p.dtype) * scale
        pdb = bf16(pd)
        ctx.save_for_backward(q, qs, kb, vb, p, pdb, keep if keep is not None else torch.empty(0))
        ctx.scale, ctx.variant, ctx.has_keep = scale, variant, keep is not None
        return bf16(pdb @ vb)

    @staticmethod
    def backward(ctx, dctx):
        q, qs, kb, vb, p, pdb, keep = ctx.saved_tensors
        dh = q.shape[-1]
        inv_sqrt = float(np.float32(1.0 / math.sqrt(dh)))
        dob = bf16(dctx)
        dp = dob @ vb.transpose(-1, -2)
        dp = dp * ctx.scale if not ctx.has_keep else dp * keep.to(dp.dtype) * ctx.scale
        delta = (dp * p).sum(-1, keepdim=True)
        ds = p * (dp - delta)
        dsq = bf16(ds * inv_sqrt)
        dq = dsq @ kb
        if ctx.variant == "h":
            # head_dim 128 of the fused d_model = 256 kernels (tc256_bwd.cu:t256_attn_bwd128): the backward arithmetic of
            # tc_attn32.cuh — c * dropped-P and c * dS as bf16 (c = 1/sqrt(dh)), ONE dS fragment feeds dq and dk
            rc1 = float(np.float32(math.sqrt(dh)))
            dsq = bf16(ds * inv_sqrt)
            pdc = p * (ctx.scale * inv_sqrt) if not ctx.has_keep else p * keep.to(p.dtype) * (ctx.scale * inv_sqrt)
            dq = dsq @ kb
            dk = (dsq.transpose(-1, -2) @ qs) * (LN2 * rc1)
            dv = (bf16(pdc).transpose(-1, -2) @ dob) * rc1
            return dq, dk, dv, None, None, None, None
        if ctx.variant == "f":
            dk = bf16(ds * LN2).transpose(-1, -2) @ qs
        else:
            dk = dsq.transpose(-1, -2) @ bf16(q)
        dv = pdb.transpose(-1, -2) @ dob
        return dq, dk, dv, None, None, None, None


class _LNq(torch.autograd.Function):
    """LayerNorm whose forward sees the fp32 sum u and whose backward sees the SAVED bf16 image of u (d_model = 256 path:
    tc256.cu writes u1 / u2 as bf16 images, tc256_bwd.cu:t256_ln_bwd recomputes mean / rstd / x-hat from them)."""

    @staticmethod
    def forward(ctx, u, g, b):
        ctx.save_for_backward(bf16(u), g)
        mu = u.mean(-1, keepdim=True)
        var = ((u - mu) ** 2).mean(-1, keepdim=True)
        return (u - mu) / torch.sqrt(var + 1e-5) * g + b

    @staticmethod
    def backward(ctx, dy):
        ub, g = ctx.saved_tensors
        mu = ub.mean(-1, keepdim=True)
        var = ((ub * ub).mean(-1, keepdim=True) - mu * mu).clamp_min(0)
        rs = 1.0 / torch.sqrt(var + 1e-5)
        xh = (ub - mu) * rs
        gd = dy * g
        m1, m2 = gd.mean(-1, keepdim=True), (gd * xh).mean(-1, keepdim=True)
        du = rs * (gd - m1 - xh * m2)
        d = dy.shape[-1]
        return du, (dy * xh).reshape(-1, d).sum(0), dy.reshape(-1, d).sum(0)


class _Stem256(torch.autograd.Function):
    """edge256.cu stem: x0 image = bf16(dropout(relu(src W^T + b) + pe)); backward: g = dx0 * mask * (r > 0) as a bf16 image,
    dW_in | db_in = bf16(g)^T [bf16(src) | 1]."""

    @staticmethod
    def forward(ctx, src, w, b, pe, mask):
        r = torch.relu(src @ w.T + b)
        ctx.save_for_backward(src, r, mask)
        return bf16((r + pe) * mask)

    @staticmethod
    def backward(ctx, dx0):
        src, r, mask = ctx.saved_tensors
        gb = bf16(dx0 * mask * (r > 0).to(dx0.dtype))
        g2 = gb.reshape(-1, gb.shape[-1])
        return None, g2.T @ bf16(src).reshape(-1, src.shape[-1]), g2.sum(0), None, None


class _Tail256(torch.autograd.Function):
    """edge256.cu tail: logits = LN(x_L image) W_out^T + b in fp32; backward: dx through the LayerNorm in fp32, parameter
    gradients from T[c][j] = sum_r x[r][c] bf16(dlogits[r][j] rstd[r]) and the two fp32 27-vectors of the file header."""

    @staticmethod
    def forward(ctx, x, gamma, beta, w, b):
        mu = x.mean(-1, keepdim=True)
        var = ((x * x).mean(-1, keepdim=True) - mu * mu).clamp_min(0)
        rs = 1.0 / torch.sqrt(var + 1e-5)
        z = (x - mu) * rs * gamma + beta
        ctx.save_for_backward(x, mu, rs, gamma, beta, w)
        return z @ w.T + b

    @staticmethod
    def backward(ctx, dl):
        x, mu, rs, gamma, beta, w = ctx.saved_tensors
        d = x.shape[-1]
        dz = dl @ w
        xh = (x - mu) * rs
        gd = dz * gamma
        dx = rs * (gd - gd.mean(-1, keepdim=True) - xh * (gd * xh).mean(-1, keepdim=True))
        dl2, x2 = dl.reshape(-1, dl.shape[-1]), x.reshape(-1, d)
        dlp = dl2 * rs.reshape(-1, 1)
        tm = x2.T @ bf16(dlp)                                   # [d, 27]; x is the bf16 image already
        s = (dlp * mu.reshape(-1, 1)).sum(0)                   # fp32, unrounded
        sb = dl2.sum(0)
        dw = (gamma[:, None] * (tm - s[None, :]) + beta[:, None] * sb[None, :]).T
        dgamma = (w.T * (tm - s[None, :])).sum(1)
        dbeta = (w.T * sb[None, :]).sum(1)
        return dx, dgamma, dbeta, dw, sb


# ----------------------------------------------------------------------------------------------
def _keep_rows(drop: G.DropCtx, shape, site):
    """fp32 multiplier (0 or 1/(1-p)) of a [N, 32, W] row site — DropCtx.rows without the multiply."""
    if not drop.active():
        return None
    n, t, w = shape
    idx = np.arange(n * t * w, dtype=np.uint64) + np.uint64(drop.seq0 * t * w)
    keep = G.dropout_keep(drop.seed, drop.step, site, idx, drop.p).reshape(n, t, w)
    return torch.from_numpy(keep).to(torch.float32) * G.dropout_scale(drop.p)


def _keep_probs(drop: G.DropCtx, n, h, site):
    if not drop.active():
        return None
    rows = (np.arange(n * h * 32, dtype=np.uint64) + np.uint64(drop.seq0 * h * 32)) * np.uint64(32)
    idx = rows[:, None] + G.key_perm(np.arange(32)).astype(np.uint64)[None, :]
    return torch.from_numpy(G.dropout_keep(drop.seed, drop.step, site, idx, drop.p).reshape(n, h, 32, 32))


def _attn_variant(path, dh, causal_or_cross_block=False):
    if path == PATH_FUSED_D32:
        return "g" if dh in (2, 4, 8) else "fp32"
    if path == PATH_FUSED_D256:
        return "h" if dh == 128 else "f"
    return "p" if dh in (16, 32, 64, 128) else "fp32"


def mha_b(P, pre, xq, xkv, nhead, drop, site, path, causal=False):
    n, t, d = xq.shape
    dh = d // nhead
    w, b = P[pre + ".in_proj_weight"], P[pre + ".in_proj_bias"]
    variant = _attn_variant(path, dh)
    lin_in = _LinBr.apply if variant == "g" else lin_b
    if xq is xkv:
        q, k, v = lin_in(xq, w, b).split(d, dim=-1)
    else:
        q = lin_in(xq, w[:d], b[:d])
        k, v = lin_in(xkv, w[d:], b[d:]).split(d, dim=-1)
    split = lambda z: z.reshape(n, -1, nhead, dh).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    if variant == "fp32":
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        if causal:
            s = s + torch.triu(torch.full((t, t), float("-inf"), dtype=s.dtype), diagonal=1)
        ctx = drop.probs(torch.softmax(s, dim=-1), site) @ v
    elif variant == "g":
        keep = _keep_probs(drop, n, nhead, site)
        ctx = _AttnCore32.apply(q, k, v, keep, G.dropout_scale(drop.p) if drop.active() else 1.0, causal)
    else:
        keep = _keep_probs(drop, n, nhead, site)
        ctx = _AttnCore.apply(q, k, v, keep, G.dropout_scale(drop.p) if drop.active() else 1.0, causal, variant)
    ctx = ctx.permute(0, 2, 1, 3).reshape(n, t, d)
    return lin_b(ctx, P[pre + ".out_proj.weight"], P[pre + ".out_proj.bias"])


def _ln(path, u, g, b):
    return _LNq.apply(u, g, b) if path == PATH_FUSED_D256 else G.layer_norm(u, g, b)


def _ffn(P, pre, x1, drop, site, path):
    w1, b1, w2, b2 = P[pre + ".linear1.weight"], P[pre + ".linear1.bias"], P[pre + ".linear2.weight"], P[pre + ".linear2.bias"]
    if path == PATH_FUSED_D32:
        keep = None
        if drop.active():
            n, t, _ = x1.shape
            idx = np.arange(n * t * w1.shape[0], dtype=np.uint64) + np.uint64(drop.seq0 * t * w1.shape[0])
            keep = torch.from_numpy(G.dropout_keep(drop.seed, drop.step, site, idx, drop.p).reshape(n, t, w1.shape[0])).to(x1.dtype)
        return _FFN32.apply(x1, w1, b1, w2, b2, keep, G.dropout_scale(drop.p) if drop.active() else 1.0)
    h = drop.rows(torch.relu(lin_b(x1, w1, b1)), site)
    return lin_b(h, w2, b2)


def encoder_layer_b(P, pre, x, nhead, drop, li, path):
    a = mha_b(P, pre + ".self_attn", x, x, nhead, drop, G.site_id(0, li, 0), path)
    x1 = _ln(path, x + drop.rows(a, G.site_id(0, li, 1)), P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    f = _ffn(P, pre, x1, drop, G.site_id(0, li, 2), path)
    out = _ln(path, x1 + drop.rows(f, G.site_id(0, li, 3)), P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    return rb(out) if path == PATH_FUSED_D256 else out


def decoder_layer_b(P, pre, y, mem, nhead, drop, li, path):
    a = mha_b(P, pre + ".self_attn", y, y, nhead, drop, G.site_id(1, li, 0), path, causal=True)
    y = G.layer_norm(y + drop.rows(a, G.site_id(1, li, 1)), P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    c = mha_b(P, pre + ".multihead_attn", y, mem, nhead, drop, G.site_id(1, li, 4), path)
    y = G.layer_norm(y + drop.rows(c, G.site_id(1, li, 5)), P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    f = _ffn(P, pre, y, drop, G.site_id(1, li, 2), path)
    return G.layer_norm(y + drop.rows(f, G.site_id(1, li, 3)), P[pre + ".norm3.weight"], P[pre + ".norm3.bias"])


def encode_b(P, cfg, src, drop, path, final_norm=True):
    pe = G.positional_table(cfg.d_model)
    if path == PATH_FUSED_D256:
        m = _keep_rows(drop, (src.shape[0], 32, cfg.d_model), G.SITE_IN_ENC)
        m = torch.ones(1) if m is None else m
        x = _Stem256.apply(src, P["InputLayerEncoder.Linear.weight"], P["InputLayerEncoder.Linear.bias"], pe, m)
    else:
        x = G.input_layer(P, "InputLayerEncoder", src, pe, drop, G.SITE_IN_ENC)
    for li in range(cfg.n_enc):
        x = encoder_layer_b(P, f"Encoder.Encoder.layers.{li}", x, cfg.nhead, drop, li, path)
    if not final_norm:
        return x
    return G.layer_norm(x, P["Encoder.Encoder.norm.weight"], P["Encoder.Encoder.norm.bias"])


def forward_b(P, cfg, src, tgt_in=None, drop=None, path=None):
    """(h, v, o) of the bf16 mode; tgt_in (already shifted) selects the encoder-decoder model."""
    drop = drop or G.DropCtx(train=False)
    path = path or path_for(cfg)
    if cfg.n_dec == 0 and path == PATH_FUSED_D256:
        x = encode_b(P, cfg, src, drop, path, final_norm=False)
        y = _Tail256.apply(x, P["Encoder.Encoder.norm.weight"], P["Encoder.Encoder.norm.bias"], P["OutputLayer.Linear.weight"],
                           P["OutputLayer.Linear.bias"])
        return y[..., :9], torch.sigmoid(y[..., 9:18]), 0.5 * torch.tanh(y[..., 18:])
    mem = encode_b(P, cfg, src, drop, path)
    if cfg.n_dec == 0:
        return G.output_layer(P, mem)
    pe = G.positional_table(cfg.d_model)
    y = G.input_layer(P, "InputLayerDecoder", tgt_in, pe, drop, G.SITE_IN_DEC)
    for li in range(cfg.n_dec):
        y = decoder_layer_b(P, f"Decoder.Decoder.layers.{li}", y, mem, cfg.nhead, drop, li, path)
    return G.output_layer(P, G.layer_norm(y, P["Decoder.Decoder.norm.weight"], P["Decoder.Decoder.norm.bias"]))


def predict_encoder_only_b(P, cfg, src, thres=0.5, path=None):
    with torch.no_grad():
        h, v, o = forward_b(P, cfg, src, path=path)
    return (torch.sigmoid(h) > thres).to(torch.int64), v, o


def train_step_oracle_b(P, cfg, x, y, penalty, drop: G.DropCtx | None = None, path=None):
    """groove_oracle.train_step_oracle in the bf16 mode: (loss6 floats, grads dict, (h, v, o))."""
    drop = drop or G.DropCtx(p=cfg.dropout, train=True)
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    pred = forward_b(Pg, cfg, x, G.shift_right(y) if cfg.n_dec > 0 else None, drop, path)
    out = G.groove_loss(pred, y, penalty)
    out[0].backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return tuple(float(t.detach()) for t in out), grads, tuple(t.detach() for t in pred)
