"""Host-side mirror of ``BaseGrooveTransformers/models/transformer.py``.

``GrooveTransformerEncoder`` / ``GrooveTransformer`` keep the reference's constructor arguments
(positional order as passed by ``initialize_model``, BGT/models/train.py:49-61), attribute names,
``state_dict`` keys and shapes, ``forward`` -> ``(h, v, o)`` and ``predict``
semantics — but hold every parameter as a view into ONE flat fp32 vector and run all arithmetic in
the sm_100a CUDA library (``_lib``).  There is no CPU / eager fallback: calling ``forward`` on
tensors that are not on a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Tuple

import torch
from torch import nn

from . import _lib

T_STEPS = _lib.T_STEPS


# ----------------------------------------------------------------------------------------------
# parameter containers (names only — they are never called)
# ----------------------------------------------------------------------------------------------
class _Bag(nn.Module):
    """A module whose only job is to own parameters / sub-bags under reference-compatible names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the fused CUDA path; not callable")


def _positional_table(d_model: int, max_len: int) -> torch.Tensor:
    """Sinusoidal table of BGT/models/utils.py:26-37: column 2i holds sin(t*w_i), 2i+1 cos(t*w_i),
    w_i = 10000^(-2i/d).  float32, shape (1, max_len, d_model) — a persistent buffer named ``pe``."""
    pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    ang = pos * freq
    table = torch.zeros(max_len, d_model, dtype=torch.float32)
    table[:, 0::2] = torch.sin(ang)
    table[:, 1::2] = torch.cos(ang)[:, : d_model // 2]
    return table.unsqueeze(0)


def _spec(d, f, e_src, e_tgt, n_enc, n_dec) -> List[Tuple[str, Tuple[int, ...]]]:
    """(state_dict name, shape) in the order of the flat vector == the C library's gt_param_layout."""
    out = [("InputLayerEncoder.Linear.weight", (d, e_src)), ("InputLayerEncoder.Linear.bias", (d,))]

    def attn(p):
        return [(p + ".in_proj_weight", (3 * d, d)), (p + ".in_proj_bias", (3 * d,)),
                (p + ".out_proj.weight", (d, d)), (p + ".out_proj.bias", (d,))]

    def ffn(p):
        return [(p + ".linear1.weight", (f, d)), (p + ".linear1.bias", (f,)),
                (p + ".linear2.weight", (d, f)), (p + ".linear2.bias", (d,))]

    def norms(p, k):
        return [(f"{p}.norm{i}.{w}", (d,)) for i in range(1, k + 1) for w in ("weight", "bias")]

    for l in range(n_enc):
        p = f"Encoder.Encoder.layers.{l}"
        out += attn(p + ".self_attn") + ffn(p) + norms(p, 2)
    out += [("Encoder.Encoder.norm.weight", (d,)), ("Encoder.Encoder.norm.bias", (d,))]
    if n_dec > 0:
        out += [("InputLayerDecoder.Linear.weight", (d, e_tgt)), ("InputLayerDecoder.Linear.bias", (d,))]
        for l in range(n_dec):
            p = f"Decoder.Decoder.layers.{l}"
            out += attn(p + ".self_attn") + attn(p + ".multihead_attn") + ffn(p) + norms(p, 3)
        out += [("Decoder.Decoder.norm.weight", (d,)), ("Decoder.Decoder.norm.bias", (d,))]
    out += [("OutputLayer.Linear.weight", (e_tgt, d)), ("OutputLayer.Linear.bias", (e_tgt,))]
    return out


class _GrooveFn(torch.autograd.Function):
    """(src, tgt_in, flat_params) -> hvo[N,32,27]; backward fills the flat gradient."""

    @staticmethod
    def forward(ctx, model, src, tgt_in, flat):
        # eval() with grad enabled (fine-tuning a frozen-dropout model, saliency maps): the reference's nn.Dropout is the
        # identity and autograd still works.  The library saves activations only in its training plan, so that case runs the
        # training plan with p = 0 — the same arithmetic as the inference plan — and backward is told the same p.
        ctx.train = bool(model.training)
        hvo, ws, step = model._run_forward(src, tgt_in, train=ctx.train, save=True)
        ctx.model, ctx.ws, ctx.step = model, ws, step
        ctx.save_for_backward(src, tgt_in if tgt_in is not None else src.new_empty(0), hvo)
        ctx.has_tgt = tgt_in is not None
        return hvo

    @staticmethod
    def backward(ctx, d_hvo):
        model = ctx.model
        src, tgt_in, hvo = ctx.saved_tensors
        g = torch.zeros_like(model._flat)
        lib = _lib.load()
        d_hvo = d_hvo.contiguous()
        cfg = model._cfg(dropout=None if ctx.train else 0.0)
        _lib.check(lib.gt_backward(C.byref(cfg), _lib.ptr(model._flat), _lib.ptr(model._pe_flat()), _lib.ptr(src),
                                   _lib.ptr(tgt_in) if ctx.has_tgt else 0, src.shape[0], _lib.ptr(hvo), _lib.ptr(d_hvo),
                                   _lib.ptr(g), _lib.ptr(ctx.ws), ctx.ws.numel(), model._seed, ctx.step, model._seq0,
                                   _lib.stream_ptr(src.device)), "gt_backward")
        ctx.ws = None
        if all(p.grad is None for p in model._views):      # optimizer.zero_grad(set_to_none=True) happened
            model._flat.grad = None
        return None, None, None, g


_PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp32_tc": _lib.PREC_FP32_TC}


class _GrooveBase(nn.Module):
    def _build(self, d_model, e_src, e_tgt, nhead, dim_ff, dropout, n_enc, n_dec, max_len, device):
        if max_len != T_STEPS:
            raise ValueError(f"max_len must be {T_STEPS} (the reference requires T == max_len, BGT/models/utils.py:49)")
        if d_model % nhead != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        if e_tgt < 3 or e_tgt % 3 != 0:
            # BGT/models/io_layers.py:34-40 splits the head's output into thirds (hits | velocities | offsets); every set of the
            # reference has 9 voices (27), which is what the fused stem / tail kernels cover — other widths run the generic kernels
            raise ValueError("embedding_size_tgt must be a positive multiple of 3 (n_voices x hit / velocity / offset)")
        self._spec_list = _spec(d_model, dim_ff, e_src, e_tgt, n_enc, n_dec)
        self.precision = "fp32"
        # dropout stream: tied to torch's seed like the reference's nn.Dropout (torch.manual_seed before construction gives a
        # reproducible run; independent processes draw independent masks); set_seed() overrides it
        self._seed, self._step, self._seq0 = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF, 0, 0
        self._train_ws = None

        lib = _lib.load()
        cfg = self._cfg()
        n = len(self._spec_list)
        offs, sizes = (C.c_int64 * n)(), (C.c_int64 * n)()
        got = lib.gt_param_layout(C.byref(cfg), offs, sizes, n)
        if got != n:
            raise RuntimeError(f"parameter layout mismatch: library has {got} tensors, host {n}: "
                               f"{lib.gt_last_error().decode()}")
        total = lib.gt_param_count(C.byref(cfg))
        flat = torch.zeros(total, dtype=torch.float32)
        self._offsets = []
        for i, (name, shape) in enumerate(self._spec_list):
            assert sizes[i] == math.prod(shape), name
            self._offsets.append((int(offs[i]), int(sizes[i])))

        # containers with reference-compatible attribute paths
        self._views: List[nn.Parameter] = []
        for (name, shape), (o, s) in zip(self._spec_list, self._offsets):
            parts = name.split(".")
            mod = self
            for part in parts[:-1]:
                if part.isdigit():
                    while len(mod) <= int(part):
                        mod.append(_Bag())
                    mod = mod[int(part)]
                else:
                    if not hasattr(mod, part):
                        setattr(mod, part, nn.ModuleList() if part == "layers" else _Bag())
                    mod = getattr(mod, part)
            p = nn.Parameter(flat[o:o + s].view(shape))
            mod.register_parameter(parts[-1], p)
            self._views.append(p)
        pe = _positional_table(d_model, max_len)
        self.InputLayerEncoder.PositionalEncoding = _Bag()
        self.InputLayerEncoder.PositionalEncoding.register_buffer("pe", pe.clone())
        if n_dec > 0:
            self.InputLayerDecoder.PositionalEncoding = _Bag()
            self.InputLayerDecoder.PositionalEncoding.register_buffer("pe", pe.clone())
        object.__setattr__(self, "_flat", flat.requires_grad_(True))
        self._flat.register_post_accumulate_grad_hook(self._bind_grads)
        self._train_ws = None
        self.reset_parameters()
        if device is not None and str(device) != "cpu":
            self.to(device)

    # ---- initialisation: same distributions as the reference modules --------------------------
    def reset_parameters(self):
        """torch defaults for nn.Linear / nn.MultiheadAttention / nn.LayerNorm as used by the reference
        (torch/nn/modules/activation.py:1232-1246, linear.py reset_parameters) + ``init_weights`` of the
        encoder input layer and the output layer (BGT/models/io_layers.py:13-15, 32-34; called at
        transformer.py:32-33,105-106).  All layers of a stack start IDENTICAL, because
        nn.TransformerEncoder deep-copies one layer (torch/nn/modules/transformer.py:1202-1204)."""
        sd = {n: p for (n, _), p in zip(self._spec_list, self._views)}
        with torch.no_grad():
            for name, p in sd.items():
                leaf = name.split(".")[-1]
                if ".layers." in name and ".layers.0." not in name:
                    continue                                      # filled from layer 0 below
                if name.startswith(("InputLayerEncoder.Linear", "OutputLayer.Linear")):
                    p.uniform_(-0.1, 0.1) if leaf == "weight" else p.zero_()
                elif "norm" in name:
                    p.fill_(1.0) if leaf == "weight" else p.zero_()
                elif leaf == "in_proj_weight":
                    nn.init.xavier_uniform_(p)
                elif leaf == "in_proj_bias" or name.endswith("out_proj.bias"):
                    p.zero_()
                elif leaf == "weight":
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                else:                                             # bias of a default nn.Linear
                    w = sd[name[: -len("bias")] + "weight"]
                    bound = 1.0 / math.sqrt(w.shape[1])
                    p.uniform_(-bound, bound)
            for name, p in sd.items():
                if ".layers." in name and ".layers.0." not in name:
                    head, rest = name.split(".layers.")
                    p.copy_(sd[head + ".layers.0." + rest.split(".", 1)[1]])

    # ---- flat-vector plumbing --------------------------------------------------------------------
    def _bind_grads(self, flat):
        g = flat.grad
        for p, (o, s) in zip(self._views, self._offsets):
            p.grad = g[o:o + s].view(p.shape)

    def _rebind_views(self, flat_data: torch.Tensor, grad: torch.Tensor | None = None):
        """Make ``flat_data`` THE flat vector: every nn.Parameter becomes a view of it again (after .to(), deepcopy, unpickling)."""
        flat = flat_data.detach().requires_grad_(True)
        object.__setattr__(self, "_flat", flat)
        flat.register_post_accumulate_grad_hook(self._bind_grads)
        for p, (o, s) in zip(self._views, self._offsets):
            p.data = flat.detach()[o:o + s].view(p.shape)
            p.grad = None
        if grad is not None:
            flat.grad = grad
            self._bind_grads(flat)
        self._train_ws = None

    def _apply(self, fn, recurse=True):
        new = fn(self._flat.detach())
        if new.dtype != torch.float32:
            raise TypeError("groove_b200 keeps fp32 master parameters; precision is selected with set_precision()")
        had_grad = self._flat.grad
        self._rebind_views(new, fn(had_grad) if had_grad is not None else None)
        for m in self.modules():
            for k, b in m._buffers.items():
                if b is not None:
                    m._buffers[k] = fn(b)
        return self

    # copy.deepcopy (best-model snapshots, EMA copies) and pickling: the default implementations would clone the flat vector and
    # every Parameter separately, leaving the copy's parameters detached from the vector the kernels read.
    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in ("_flat", "_train_ws"):
                new.__dict__[k] = copy.deepcopy(v, memo)
        g = self._flat.grad
        new._rebind_views(self._flat.detach().clone(), g.detach().clone() if g is not None else None)
        return new

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items() if k not in ("_flat", "_train_ws")}
        state["_flat_data"] = self._flat.detach()
        return state

    def __setstate__(self, state):
        flat = state.pop("_flat_data")
        self.__dict__.update(state)
        self._rebind_views(flat)

    def flat_parameters(self) -> torch.Tensor:  # noqa: D401
        """The single fp32 vector every parameter is a view of (C layout: gt_param_layout)."""
        return self._flat

    def flat_grad(self) -> torch.Tensor:
        if self._flat.grad is None:
            self._flat.grad = torch.zeros_like(self._flat)
            self._bind_grads(self._flat)
        return self._flat.grad

    def _pe_flat(self):
        return self.InputLayerEncoder.PositionalEncoding.pe

    def set_precision(self, mode: str):
        """'fp32' (exact SIMT kernels), 'fp32_tc' (the same 1e-4 parity mode with every Linear contraction on the tcgen05
        tensor cores: fp32 operands split exactly into three bf16 terms, GT_PREC_FP32_TC) or 'bf16' (fused tcgen05 kernels)."""
        if mode not in _PRECISIONS:
            raise ValueError("precision must be 'fp32', 'fp32_tc' or 'bf16'")
        self.precision = mode
        self._train_ws = None
        return self

    def set_seed(self, seed: int, step: int = 0, seq0: int = 0):
        """Dropout stream: masks are a pure function of (seed, step, site, global element index)."""
        self._seed, self._step, self._seq0 = int(seed), int(step), int(seq0)
        return self

    def _cfg(self, dropout=None):
        n_dec = getattr(self, "num_decoder_layers", 0)
        return _lib.GtConfig(self.d_model, self.nhead, self.dim_feedforward, self.num_encoder_layers, n_dec,
                             self.embedding_size_src, self.embedding_size_tgt,
                             _PRECISIONS[getattr(self, "precision", "fp32")],
                             float(self.dropout if dropout is None else dropout), 0)

    def _check_input(self, x, e, what):
        if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[1] != T_STEPS or x.shape[2] != e:
            raise ValueError(f"{what} must have shape [N, {T_STEPS}, {e}], got {tuple(getattr(x, 'shape', ()))}")
        if x.shape[0] == 0:
            raise ValueError(f"{what}: empty batch")
        if not x.is_cuda or not self._flat.is_cuda:
            raise RuntimeError("groove_b200 runs on CUDA (sm_100a) only — there is no CPU fallback; "
                               "move the model and its inputs to a cuda device")
        if x.device != self._flat.device:
            raise RuntimeError("input and model are on different devices")
        return x.contiguous().float()

    def _workspace(self, n_seq, mode, device):
        lib = _lib.load()
        cfg = self._cfg()
        nbytes = lib.gt_workspace_bytes(C.byref(cfg), n_seq, mode)
        if nbytes < 0:
            raise RuntimeError(lib.gt_last_error().decode())
        return torch.empty(nbytes, dtype=torch.uint8, device=device)

    def _run_forward(self, src, tgt_in, train, save):
        lib = _lib.load()
        n = src.shape[0]
        ws = self._workspace(n, 1 if (save or train) else 0, src.device)     # gt_forward(train=1) saves activations
        hvo = torch.empty(n, T_STEPS, self.embedding_size_tgt, dtype=torch.float32, device=src.device)
        # save without train (eval-mode autograd): the training plan (activations saved) with dropout switched off
        cfg = self._cfg(dropout=0.0 if (save and not train) else None)
        step = self._step
        if train:
            self._step += 1
        _lib.check(lib.gt_forward(C.byref(cfg), _lib.ptr(self._flat), _lib.ptr(self._pe_flat()), _lib.ptr(src),
                                  _lib.ptr(tgt_in), n, _lib.ptr(hvo), _lib.ptr(ws), ws.numel(), 1 if (train or save) else 0,
                                  self._seed, step, self._seq0, _lib.stream_ptr(src.device)), "gt_forward")
        return hvo, ws, step

    def _forward_hvo(self, src, tgt_in):
        needs_grad = torch.is_grad_enabled() and self._flat.requires_grad
        if needs_grad:
            return _GrooveFn.apply(self, src, tgt_in, self._flat)
        hvo, _, _ = self._run_forward(src, tgt_in, train=self.training, save=False)
        return hvo

    @staticmethod
    def _split(hvo):
        nv = hvo.shape[-1] // 3
        h, v, o = hvo[..., 0:nv], hvo[..., nv:2 * nv], hvo[..., 2 * nv:3 * nv]
        for t in (h, v, o):
            t._groove_hvo = hvo       # lets calculate_loss find the packed [N,32,27] tensor without a copy
        return h, v, o

    # ---- fused training step (what train_loop / bench.py use) ------------------------------------
    def train_step(self, x, y, hit_loss_penalty: float, grads: torch.Tensor | None = None):
        """forward(train) + calculate_loss + backward in ONE library call (gt_train_step).
        Returns (metrics6 device tensor [loss, acc, ppl, bce, mse_v, mse_o], hvo).  The flat gradient
        (``flat_grad()``) is overwritten, not accumulated."""
        lib = _lib.load()
        x = self._check_input(x, self.embedding_size_src, "x")
        y = self._check_input(y, self.embedding_size_tgt, "y")
        if x.shape[0] != y.shape[0]:
            raise ValueError("x and y batch sizes differ")
        n = x.shape[0]
        g = self.flat_grad() if grads is None else grads
        cfg = self._cfg()
        key = (n, self.precision)
        if self._train_ws is None or self._train_ws[0] != key:
            self._train_ws = (key, self._workspace(n, 1, x.device),
                              torch.empty(n, T_STEPS, self.embedding_size_tgt, dtype=torch.float32, device=x.device))
        _, ws, hvo = self._train_ws
        metrics = torch.empty(6, dtype=torch.float32, device=x.device)
        step = self._step
        self._step += 1
        _lib.check(lib.gt_train_step(C.byref(cfg), _lib.ptr(self._flat), _lib.ptr(self._pe_flat()), _lib.ptr(x), _lib.ptr(y),
                                     n, float(hit_loss_penalty), _lib.ptr(g), _lib.ptr(metrics), _lib.ptr(hvo), _lib.ptr(ws),
                                     ws.numel(), self._seed, step, self._seq0, _lib.stream_ptr(x.device)), "gt_train_step")
        return metrics, hvo

    def _predict_hvo(self, src, thres, literal: bool = False, out=None):
        """gt_predict; ``literal=True`` runs the reference's 32 full decoder passes instead of the KV-cached decode
        (encoder-decoder models; used by the tests to cross-check the two).  ``out``: a contiguous [n, 32, 27] fp32 device
        tensor to write into (pipeline.HostPredictor's slots)."""
        lib = _lib.load()
        src = self._check_input(src, self.embedding_size_src, "src")
        n = src.shape[0]
        ws = self._workspace(n, 0 if literal else 2, src.device)
        if out is None:
            out = torch.empty(n, T_STEPS, self.embedding_size_tgt, dtype=torch.float32, device=src.device)
        elif tuple(out.shape) != (n, T_STEPS, self.embedding_size_tgt) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != src.device:
            raise ValueError("out must be a contiguous float32 [n, 32, embedding_size_tgt] tensor on the source's device")
        cfg = self._cfg()
        _lib.check(lib.gt_predict_variant(C.byref(cfg), _lib.ptr(self._flat), _lib.ptr(self._pe_flat()), _lib.ptr(src), n,
                                          float(thres), _lib.ptr(out), _lib.ptr(ws), ws.numel(), 1 if literal else 0,
                                          _lib.stream_ptr(src.device)), "gt_predict")
        return out

    @staticmethod
    def _check_predict_flags(use_thres, use_pd):
        # BGT/models/utils.py:59-69: the use_pd path raises a broadcast error in the reference and
        # use_thres=False leaves `h` undefined — only the threshold path exists.
        if use_pd:
            raise NotImplementedError("use_pd=True is broken in the reference (shape mismatch) and is not provided")
        if not use_thres:
            raise NotImplementedError("use_thres=False leaves the hits undefined in the reference")


class GrooveTransformerEncoder(_GrooveBase):
    """Drop-in for BGT/models/transformer.py:86-125."""

    def __init__(self, d_model, embedding_size_src, embedding_size_tgt, nhead, dim_feedforward, dropout,
                 num_encoder_layers, max_len, device):
        super().__init__()
        self.d_model = d_model
        self.embedding_size_src = embedding_size_src
        self.embedding_size_tgt = embedding_size_tgt
        self.nhead = nhead
        self.dim_feedforward = dim_feedforward
        self.dropout = dropout
        self.max_len = max_len
        self.num_encoder_layers = num_encoder_layers
        self.device = device
        self._build(d_model, embedding_size_src, embedding_size_tgt, nhead, dim_feedforward, dropout,
                    num_encoder_layers, 0, max_len, device)

    def forward(self, src):
        src = self._check_input(src, self.embedding_size_src, "src")
        return self._split(self._forward_hvo(src, None))

    def predict(self, src, use_thres=True, thres=0.5, use_pd=False):
        self._check_predict_flags(use_thres, use_pd)
        self.eval()
        with torch.no_grad():
            out = self._predict_hvo(src, thres)
        nv = self.embedding_size_tgt // 3
        return out[..., 0:nv].to(torch.int64), out[..., nv:2 * nv], out[..., 2 * nv:]


class GrooveTransformer(_GrooveBase):
    """Drop-in for BGT/models/transformer.py:9-83."""

    def __init__(self, d_model, embedding_size_src, embedding_size_tgt, nhead, dim_feedforward, dropout,
                 num_encoder_layers, num_decoder_layers, max_len, device):
        super().__init__()
        self.d_model = d_model
        self.embedding_size_src = embedding_size_src
        self.embedding_size_tgt = embedding_size_tgt
        self.nhead = nhead
        self.dim_feedforward = dim_feedforward
        self.dropout = dropout
        self.max_len = max_len
        self.num_encoder_layers = num_encoder_layers
        self.num_decoder_layers = num_decoder_layers
        self.device = device
        if num_decoder_layers < 1:
            raise ValueError("GrooveTransformer needs at least one decoder layer; use GrooveTransformerEncoder")
        self._build(d_model, embedding_size_src, embedding_size_tgt, nhead, dim_feedforward, dropout,
                    num_encoder_layers, num_decoder_layers, max_len, device)

    def forward(self, src, tgt):
        src = self._check_input(src, self.embedding_size_src, "src")
        tgt = self._check_input(tgt, self.embedding_size_tgt, "tgt")
        if src.shape[0] != tgt.shape[0]:
            raise ValueError("src and tgt batch sizes differ")
        return self._split(self._forward_hvo(src, tgt))

    def predict(self, src, use_thres=True, thres=0.5, use_pd=False):
        self._check_predict_flags(use_thres, use_pd)
        self.eval()
        with torch.no_grad():
            out = self._predict_hvo(src, thres)
        nv = self.embedding_size_tgt // 3
        return out[..., 0:nv], out[..., nv:2 * nv], out[..., 2 * nv:]      # all float32, like the reference
