"""CPU oracle for the Transformer Groove Infilling train / inference step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``transformergrooveinfilling_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and there only as the checker or the reported baseline.

This is a *functional restatement* (plain tensor arithmetic on CPU, fp32 or fp64) of what the
reference computes through ``torch.nn.Transformer*`` modules.  Every function cites the reference
lines it follows (paths relative to /root/reference, ``BGT`` = ``BaseGrooveTransformers``;
``torch/`` = the PyTorch source the reference delegates its arithmetic to).

Parity pinning: the reference ships no golden vectors or known-answer tests for this path
(SURVEY.md §4, §8c).  The restatement is therefore pinned against *outputs of the reference
itself*: ``oracle/make_golden.py`` imports the unmodified reference from /root/reference, runs it
on deterministic inputs/weights (``det_uniform`` below) and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those vectors.

Dropout: the reference draws masks from torch's global Philox stream, which no fused kernel can
reproduce.  Exact parity is defined at p=0; for p>0 the product uses the counter-based generator
restated in ``dropout_keep`` (bit-exact integer arithmetic), and this oracle consumes the same
masks, so p>0 results are also compared exactly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

T_STEPS = 32          # BGT max_len; train.py:128 hard-wires 32
N_VOICES = 9

# When the port is TIMED as the CPU baseline (bench.py) it uses the same fused ATen CPU ops the
# reference's nn modules dispatch to (F.linear, F.layer_norm, F.scaled_dot_product_attention,
# F.dropout), so the baseline is not handicapped by the spelled-out arithmetic used for parity.
FAST_BASELINE = False

# ----------------------------------------------------------------------------------------------
# deterministic integer hashing (shared definition with csrc/rng.cuh)
# ----------------------------------------------------------------------------------------------
_M32 = np.uint64(0xFFFFFFFF)


def _u32(x):
    return np.asarray(x, dtype=np.uint64) & _M32


def mix32(x):
    """lowbias32 integer finaliser on uint32 values held in uint64 arrays."""
    x = _u32(x)
    x ^= x >> np.uint64(16)
    x = _u32(x * np.uint64(0x7FEB352D))
    x ^= x >> np.uint64(15)
    x = _u32(x * np.uint64(0x846CA68B))
    x ^= x >> np.uint64(16)
    return x


def site_key(seed: int, step: int, site: int) -> int:
    """32-bit key for one dropout site of one optimisation step."""
    k = int(mix32(np.uint64((seed & 0xFFFFFFFF) ^ 0x9E3779B9)))
    k = int(mix32(np.uint64(k ^ ((seed >> 32) & 0xFFFFFFFF))))
    k = int(mix32(np.uint64((k + (step & 0xFFFFFFFF) * 0x85EBCA6B) & 0xFFFFFFFF)))
    k = int(mix32(np.uint64((k ^ ((site & 0xFFFFFFFF) * 0xC2B2AE35)) & 0xFFFFFFFF)))
    return k


DROP_FIELD_ONE = 1 << 14      # csrc/common.cuh: a dropout decision is a 14-bit field compared with a 14-bit threshold


def dropout_threshold(p: float) -> int:
    """14-bit drop threshold: an element is KEPT iff its 14-bit random field >= threshold."""
    return int(min(DROP_FIELD_ONE - 1, max(0, round(p * float(DROP_FIELD_ONE)))))


def hash_quad(v, key):
    """Four 14-bit fields (the low 14 bits of each 16-bit half of lo, hi; uint32 values held in uint64 arrays) per quad
    value ``v`` — csrc/common.cuh:hash_quad."""
    x = _u32(_u32(v) * np.uint64(0x9E3779B1)) ^ np.uint64(key)
    x ^= x >> np.uint64(16)
    p = x * np.uint64(0x7FEB352D)
    plo, phi = _u32(p), p >> np.uint64(32)
    y = plo ^ phi
    q = y * np.uint64(0x846CA68B)
    lo = (_u32(q) ^ phi) & np.uint64(0x3FFF3FFF)
    hi = ((q >> np.uint64(32)) ^ _u32((y << np.uint64(16)) | (y >> np.uint64(16)))) & np.uint64(0x3FFF3FFF)
    return lo, hi


def dropout_keep(seed: int, step: int, site: int, idx: np.ndarray, p: float) -> np.ndarray:
    """Keep-mask for element indices ``idx`` (uint64).  One 64-bit hash serves the quad of elements
    4w..4w+3: element k of the quad reads 16-bit field k of (lo, hi) — csrc/common.cuh:drop_keep."""
    key = site_key(seed, step, site)
    idx = np.asarray(idx, dtype=np.uint64)
    w = idx >> np.uint64(2)
    v = _u32(w) ^ _u32((w >> np.uint64(32)) * np.uint64(0x85EBCA6B))
    lo, hi = hash_quad(v, key)
    k = idx & np.uint64(3)
    word = np.where(k >= 2, hi, lo)
    half = np.where((k & np.uint64(1)) == 1, word >> np.uint64(16), word & np.uint64(0xFFFF))
    return half >= np.uint64(dropout_threshold(p))


def key_perm(k):
    """Position of key k inside a query row at the attention-probability sites (csrc/common.cuh:key_perm)."""
    k = np.asarray(k)
    return (k & 16) | (((k >> 1) & 3) << 2) | (((k >> 3) & 1) << 1) | (k & 1)


def dropout_scale(p: float) -> float:
    """1/(1-p_eff) with p_eff the 14-bit quantised probability actually applied."""
    thr = dropout_threshold(p)
    return 1.0 if thr == 0 else float(np.float32(float(DROP_FIELD_ONE) / (float(DROP_FIELD_ONE) - thr)))


def det_uniform(tag: int, n: int, lo: float, hi: float) -> np.ndarray:
    """Deterministic pseudo-random float32 vector in [lo,hi): used for weights / inputs of the
    golden fixtures so that fixtures only need to store OUTPUTS."""
    i = np.arange(n, dtype=np.uint64)
    h = mix32(_u32(i * np.uint64(0x9E3779B1)) ^ np.uint64(mix32(np.uint64(tag & 0xFFFFFFFF))))
    u = (h >> np.uint64(8)).astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# configuration + site numbering (shared with csrc/groove_config.h)
# ----------------------------------------------------------------------------------------------
@dataclass
class GrooveCfg:
    d_model: int
    nhead: int
    dim_ff: int
    n_enc: int
    n_dec: int = 0
    e_src: int = 16
    e_tgt: int = 27
    dropout: float = 0.0

    @property
    def dh(self):
        return self.d_model // self.nhead


SITE_IN_ENC = 1
SITE_IN_DEC = 2


def site_id(stack: int, layer: int, k: int) -> int:
    """stack 0 = encoder, 1 = decoder.  k: 0 self-attn probs, 1 dropout after self-attn out-proj,
    2 FFN hidden dropout, 3 dropout after FFN, 4 cross-attn probs, 5 dropout after cross out-proj."""
    return 16 + (stack * 64 + layer) * 8 + k


class DropCtx:
    """Carries (p, seed, step, first global sequence index) and applies the shared masks."""

    def __init__(self, p=0.0, seed=0, step=0, seq0=0, train=True, native=False):
        self.p, self.seed, self.step, self.seq0, self.train = p, seed, step, seq0, train
        # native=True: torch's own dropout (what the reference's nn.Dropout does) — used only when the
        # oracle is TIMED as the CPU baseline, never for parity
        self.native = native

    def active(self):
        return self.train and dropout_threshold(self.p) > 0

    def rows(self, x: torch.Tensor, site: int) -> torch.Tensor:
        """x: [N, 32, W]; element index = ((seq0+n)*32 + t)*W + c."""
        if not self.active():
            return x
        if self.native:
            return torch.nn.functional.dropout(x, self.p, True)
        n, t, w = x.shape
        idx = (np.arange(n * t * w, dtype=np.uint64) + np.uint64(self.seq0 * t * w))
        keep = dropout_keep(self.seed, self.step, site, idx, self.p).reshape(n, t, w)
        return x * torch.from_numpy(keep).to(x.dtype) * dropout_scale(self.p)

    def probs(self, pr: torch.Tensor, site: int) -> torch.Tensor:
        """pr: [N, H, 32, 32]; element index = (((seq0+n)*H + h)*32 + i)*32 + key_perm(j)."""
        if not self.active():
            return pr
        if self.native:
            return torch.nn.functional.dropout(pr, self.p, True)
        n, h, a, b = pr.shape
        rows = (np.arange(n * h * a, dtype=np.uint64) + np.uint64(self.seq0 * h * a)) * np.uint64(b)
        idx = rows[:, None] + key_perm(np.arange(b)).astype(np.uint64)[None, :]
        keep = dropout_keep(self.seed, self.step, site, idx, self.p).reshape(n, h, a, b)
        return pr * torch.from_numpy(keep).to(pr.dtype) * dropout_scale(self.p)


# ----------------------------------------------------------------------------------------------
# model arithmetic
# ----------------------------------------------------------------------------------------------
def positional_table(d_model: int, max_len: int = T_STEPS) -> torch.Tensor:
    """BGT/models/utils.py:26-37 — pe[t,2i]=sin(t*w_i), pe[t,2i+1]=cos(t*w_i), w_i=exp(-2i ln(1e4)/d);
    computed in float32 like the reference.  Shape (1, max_len, d)."""
    t = torch.arange(max_len, dtype=torch.float32)[:, None]
    w = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    tab = torch.zeros(max_len, d_model, dtype=torch.float32)
    tab[:, 0::2] = torch.sin(t * w)
    tab[:, 1::2] = torch.cos(t * w)[:, : d_model // 2]
    return tab[None]


def lin(x, w, b):
    """y = x W^T + b (torch.nn.Linear)."""
    if FAST_BASELINE:
        return torch.nn.functional.linear(x, w, b)
    return x @ w.T + b


def layer_norm(x, g, b, eps=1e-5):
    """torch.nn.LayerNorm over the last dim, biased variance (torch/nn/functional.py layer_norm)."""
    if FAST_BASELINE:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def input_layer(P, pre, x, pe, drop: DropCtx, site):
    """BGT/models/io_layers.py:17-22 — dropout(relu(x W^T + b) + pe)."""
    r = torch.relu(lin(x, P[pre + ".Linear.weight"], P[pre + ".Linear.bias"]))
    return drop.rows(r + pe.to(r.dtype), site)


def mha(P, pre, xq, xkv, nhead, drop: DropCtx, site, causal=False):
    """torch/nn/functional.py:6244 multi_head_attention_forward with packed in-proj (:6478), heads =
    contiguous dh-wide column slices (:6554), softmax(q k^T / sqrt(dh)) with dropout on the
    probabilities (:6682) and out-proj (:6690).  Batch-first restatement: x is [N, 32, d]."""
    n, t, d = xq.shape
    dh = d // nhead
    w, b = P[pre + ".in_proj_weight"], P[pre + ".in_proj_bias"]
    if FAST_BASELINE and xq is xkv:
        q, k, v = torch.nn.functional.linear(xq, w, b).chunk(3, dim=-1)
    else:
        q = xq @ w[:d].T + b[:d]
        k = xkv @ w[d:2 * d].T + b[d:2 * d]
        v = xkv @ w[2 * d:].T + b[2 * d:]
    split = lambda z: z.reshape(n, -1, nhead, dh).permute(0, 2, 1, 3)      # [N,H,T,dh]
    q, k, v = split(q), split(k), split(v)
    if FAST_BASELINE:
        ctx = torch.nn.functional.scaled_dot_product_attention(q, k, v, dropout_p=drop.p if drop.active() else 0.0,
                                                               is_causal=causal)
        ctx = ctx.permute(0, 2, 1, 3).reshape(n, t, d)
        return torch.nn.functional.linear(ctx, P[pre + ".out_proj.weight"], P[pre + ".out_proj.bias"])
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if causal:   # BGT/models/utils.py:53-56 — 0 on/below the diagonal, -inf above
        s = s + torch.triu(torch.full((t, t), float("-inf"), dtype=s.dtype), diagonal=1)
    pr = drop.probs(torch.softmax(s, dim=-1), site)
    ctx = (pr @ v).permute(0, 2, 1, 3).reshape(n, t, d)
    return lin(ctx, P[pre + ".out_proj.weight"], P[pre + ".out_proj.bias"])


def encoder_layer(P, pre, x, nhead, drop: DropCtx, li):
    """torch/nn/modules/transformer.py:951-956 (post-norm), _sa_block :961-977, _ff_block :980-982."""
    a = mha(P, pre + ".self_attn", x, x, nhead, drop, site_id(0, li, 0))
    x = layer_norm(x + drop.rows(a, site_id(0, li, 1)), P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    h = drop.rows(torch.relu(lin(x, P[pre + ".linear1.weight"], P[pre + ".linear1.bias"])), site_id(0, li, 2))
    f = lin(h, P[pre + ".linear2.weight"], P[pre + ".linear2.bias"])
    return layer_norm(x + drop.rows(f, site_id(0, li, 3)), P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])


def decoder_layer(P, pre, y, mem, nhead, drop: DropCtx, li):
    """torch/nn/modules/transformer.py:1143-1153 (post-norm): causal self-attn, cross-attn, FFN."""
    a = mha(P, pre + ".self_attn", y, y, nhead, drop, site_id(1, li, 0), causal=True)
    y = layer_norm(y + drop.rows(a, site_id(1, li, 1)), P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    c = mha(P, pre + ".multihead_attn", y, mem, nhead, drop, site_id(1, li, 4))
    y = layer_norm(y + drop.rows(c, site_id(1, li, 5)), P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    h = drop.rows(torch.relu(lin(y, P[pre + ".linear1.weight"], P[pre + ".linear1.bias"])), site_id(1, li, 2))
    f = lin(h, P[pre + ".linear2.weight"], P[pre + ".linear2.bias"])
    return layer_norm(y + drop.rows(f, site_id(1, li, 3)), P[pre + ".norm3.weight"], P[pre + ".norm3.bias"])


def output_layer(P, z):
    """BGT/models/io_layers.py:36-48 — channels 0-8 raw hit logits, 9-17 sigmoid, 18-26 0.5*tanh."""
    y = lin(z, P["OutputLayer.Linear.weight"], P["OutputLayer.Linear.bias"])
    v9 = y.shape[-1] // 3
    return y[..., :v9], torch.sigmoid(y[..., v9:2 * v9]), 0.5 * torch.tanh(y[..., 2 * v9:])


def encode(P, cfg: GrooveCfg, src, drop: DropCtx):
    """BGT/models/encoder.py:12-16 — L encoder layers + final LayerNorm."""
    pe = positional_table(cfg.d_model)
    x = input_layer(P, "InputLayerEncoder", src, pe, drop, SITE_IN_ENC)
    for li in range(cfg.n_enc):
        x = encoder_layer(P, f"Encoder.Encoder.layers.{li}", x, cfg.nhead, drop, li)
    return layer_norm(x, P["Encoder.Encoder.norm.weight"], P["Encoder.Encoder.norm.bias"])


def forward_encoder_only(P, cfg, src, drop=None):
    """BGT/models/transformer.py:108-115 GrooveTransformerEncoder.forward."""
    drop = drop or DropCtx(train=False)
    return output_layer(P, encode(P, cfg, src, drop))


def decode(P, cfg, tgt, mem, drop: DropCtx):
    """BGT/models/decoder.py:12-23."""
    pe = positional_table(cfg.d_model)
    y = input_layer(P, "InputLayerDecoder", tgt, pe, drop, SITE_IN_DEC)
    for li in range(cfg.n_dec):
        y = decoder_layer(P, f"Decoder.Decoder.layers.{li}", y, mem, cfg.nhead, drop, li)
    return layer_norm(y, P["Decoder.Decoder.norm.weight"], P["Decoder.Decoder.norm.bias"])


def forward_encdec(P, cfg, src, tgt, drop=None):
    """BGT/models/transformer.py:35-46 GrooveTransformer.forward."""
    drop = drop or DropCtx(train=False)
    mem = encode(P, cfg, src, drop)
    return output_layer(P, decode(P, cfg, tgt, mem, drop))


def shift_right(y):
    """BGT/models/train.py:130-131 — prepend one all-zero step, drop the last."""
    return torch.cat([torch.zeros_like(y[:, :1]), y[:, :-1]], dim=1)


def predict_encoder_only(P, cfg, src, thres=0.5):
    """BGT/models/transformer.py:117-125 + utils.py:59-69 — h is int64, v/o float."""
    h, v, o = forward_encoder_only(P, cfg, src)
    return (torch.sigmoid(h) > thres).to(torch.int64), v, o


def predict_encdec(P, cfg, src, thres=0.5):
    """BGT/models/transformer.py:48-83 — 32 full decoder passes, feeding back thresholded hits and raw
    v,o of step i into position i+1 of the shifted target.  Returns float tensors."""
    drop = DropCtx(train=False)
    mem = encode(P, cfg, src, drop)
    n = src.shape[0]
    tgt = torch.zeros(n, T_STEPS + 1, cfg.e_tgt, dtype=src.dtype)
    v9 = cfg.e_tgt // 3
    for i in range(T_STEPS):
        h, v, o = output_layer(P, decode(P, cfg, tgt[:, :-1], mem, drop))
        tgt[:, i + 1, :v9] = (torch.sigmoid(h[:, i]) > thres).to(src.dtype)
        tgt[:, i + 1, v9:2 * v9] = v[:, i]
        tgt[:, i + 1, 2 * v9:] = o[:, i]
    out = tgt[:, 1:]
    return out[..., :v9], out[..., v9:2 * v9], out[..., 2 * v9:]


def groove_loss(pred, y, penalty):
    """BGT/models/train.py:9-40 calculate_loss.  Returns (total, acc, perplexity, bce, mse_v, mse_o)
    as 0-dim tensors (the reference returns .item() for the last five)."""
    h, v, o = pred
    v9 = y.shape[2] // 3
    yh, yv, yo = y[..., :v9], y[..., v9:2 * v9], y[..., 2 * v9:]
    w = torch.where(yh == 1, torch.ones_like(yh), torch.full_like(yh, float(penalty)))
    # BCEWithLogits(reduction='none'): softplus(h) - h*y, computed stably
    bce_el = torch.clamp(h, min=0) - h * yh + torch.log1p(torch.exp(-h.abs()))
    bce = (bce_el * w).sum(2).mean()
    msev = (((v - yv) ** 2) * w).sum(2).mean()
    mseo = (((o - yo) ** 2) * w).sum(2).mean()
    hit = (torch.sigmoid(h) > 0.5).to(yh.dtype)
    acc = (hit == yh).to(h.dtype).reshape(h.shape[0], -1).mean(-1).mean()
    return bce + msev + mseo, acc, torch.exp(bce), bce, msev, mseo


def sgd_step(p, g, lr):
    """torch.optim.SGD(lr) with momentum 0, wd 0 (BGT/models/train.py:65-66): p -= lr*g."""
    return p - lr * g


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (BGT/models/train.py:64).  ``step`` is 1-based."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return p - lr * mhat / (vhat.sqrt() + eps), m, v


# ----------------------------------------------------------------------------------------------
# helpers shared by tests / bench / golden generation
# ----------------------------------------------------------------------------------------------
def param_shapes(cfg: GrooveCfg):
    """state_dict names -> shapes in reference order (enumerated from the live reference modules,
    SURVEY.md §8b); excludes the ``pe`` buffers."""
    d, f = cfg.d_model, cfg.dim_ff
    out = [("InputLayerEncoder.Linear.weight", (d, cfg.e_src)), ("InputLayerEncoder.Linear.bias", (d,))]

    def attn(pre):
        return [(pre + ".in_proj_weight", (3 * d, d)), (pre + ".in_proj_bias", (3 * d,)),
                (pre + ".out_proj.weight", (d, d)), (pre + ".out_proj.bias", (d,))]

    def ffn(pre):
        return [(pre + ".linear1.weight", (f, d)), (pre + ".linear1.bias", (f,)),
                (pre + ".linear2.weight", (d, f)), (pre + ".linear2.bias", (d,))]

    def norms(pre, k):
        r = []
        for i in range(1, k + 1):
            r += [(f"{pre}.norm{i}.weight", (d,)), (f"{pre}.norm{i}.bias", (d,))]
        return r

    for li in range(cfg.n_enc):
        pre = f"Encoder.Encoder.layers.{li}"
        out += attn(pre + ".self_attn") + ffn(pre) + norms(pre, 2)
    out += [("Encoder.Encoder.norm.weight", (d,)), ("Encoder.Encoder.norm.bias", (d,))]
    if cfg.n_dec > 0:
        out += [("InputLayerDecoder.Linear.weight", (d, cfg.e_tgt)), ("InputLayerDecoder.Linear.bias", (d,))]
        for li in range(cfg.n_dec):
            pre = f"Decoder.Decoder.layers.{li}"
            out += attn(pre + ".self_attn") + attn(pre + ".multihead_attn") + ffn(pre) + norms(pre, 3)
        out += [("Decoder.Decoder.norm.weight", (d,)), ("Decoder.Decoder.norm.bias", (d,))]
    out += [("OutputLayer.Linear.weight", (cfg.e_tgt, d)), ("OutputLayer.Linear.bias", (cfg.e_tgt,))]
    return out


def det_params(cfg: GrooveCfg, tag: int = 7, dtype=torch.float32):
    """Deterministic weights: matrices U(-a,a) with a = sqrt(3/fan_in)-ish, LN gains near 1,
    biases small but NON-zero (so bias gradients/paths are exercised)."""
    P = {}
    for i, (name, shp) in enumerate(param_shapes(cfg)):
        n = int(np.prod(shp))
        if len(shp) == 2:
            a = math.sqrt(3.0 / shp[1])
            vals = det_uniform(tag * 1000 + i, n, -a, a)
        elif "norm" in name and name.endswith("weight"):
            vals = det_uniform(tag * 1000 + i, n, 0.8, 1.2)
        else:
            vals = det_uniform(tag * 1000 + i, n, -0.1, 0.1)
        P[name] = torch.from_numpy(vals.reshape(shp).copy()).to(dtype)
    return P


def det_batch(cfg: GrooveCfg, n: int, tag: int = 11, dtype=torch.float32):
    """Synthetic batch per SURVEY.md §8d: y = cat(hits~B(.15), vel*hits, off*hits); MSO x (E=16) =
    cat(strength*m, timing*m), m~B(.5); symbolic x (E=27) = an independent HVO draw."""
    def hvo(tg, nv=9):
        hits = det_uniform(tg, n * 32 * nv, 0, 1).reshape(n, 32, nv) < 0.15
        vel = det_uniform(tg + 1, n * 32 * nv, 0, 1).reshape(n, 32, nv) * hits
        off = (det_uniform(tg + 2, n * 32 * nv, 0, 1).reshape(n, 32, nv) - 0.5) * hits
        return np.concatenate([hits.astype(np.float32), vel, off], axis=2).astype(np.float32)

    y = hvo(tag * 100, cfg.e_tgt // 3)          # embedding_size_tgt = 3 x voices (9 in every set of the reference)
    if cfg.e_src == 27:
        x = hvo(tag * 100 + 10)
    else:
        half = cfg.e_src // 2
        m = det_uniform(tag * 100 + 20, n * 32 * half, 0, 1).reshape(n, 32, half) < 0.5
        s = det_uniform(tag * 100 + 21, n * 32 * half, 0, 1).reshape(n, 32, half) * m
        t = (det_uniform(tag * 100 + 22, n * 32 * half, 0, 1).reshape(n, 32, half) - 0.5) * m
        x = np.concatenate([s, t], axis=2).astype(np.float32)
        if x.shape[2] < cfg.e_src:
            x = np.concatenate([x, np.zeros((n, 32, cfg.e_src - x.shape[2]), np.float32)], axis=2)
    return torch.from_numpy(x).to(dtype), torch.from_numpy(y).to(dtype)


def train_step_oracle(P, cfg, x, y, penalty, drop: DropCtx | None = None):
    """One forward + loss + backward on CPU through autograd over the restatement.
    Returns (loss tuple of floats, grads dict, (h, v, o) detached)."""
    drop = drop or DropCtx(p=cfg.dropout, train=True)
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    if cfg.n_dec > 0:
        pred = forward_encdec(Pg, cfg, x, shift_right(y), drop)
    else:
        pred = output_layer(Pg, encode(Pg, cfg, x, drop))
    out = groove_loss(pred, y, penalty)
    out[0].backward()
    grads = {k: v.grad.detach() for k, v in Pg.items()}
    return tuple(float(t.detach()) for t in out), grads, tuple(t.detach() for t in pred)
