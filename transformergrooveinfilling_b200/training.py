"""Host-side mirror of ``BaseGrooveTransformers/models/train.py``: ``calculate_loss``,
``initialize_model`` and ``train_loop`` with the reference's signatures, return values, wandb keys,
checkpoint format and resume rules — arithmetic in the CUDA library."""
from __future__ import annotations

import ctypes as C
import math
import os
import re

import numpy as np
import torch

from . import _lib
from .modules import GrooveTransformer, GrooveTransformerEncoder, _GrooveBase

try:  # logging only; never on the compute path
    import wandb  # type: ignore
except Exception:  # pragma: no cover
    wandb = None


# ----------------------------------------------------------------------------------------------
# calculate_loss  (BGT/models/train.py:9-40)
# ----------------------------------------------------------------------------------------------
class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hvo, y, penalty):
        lib = _lib.load()
        n = hvo.shape[0]
        metrics = torch.empty(6, dtype=torch.float32, device=hvo.device)
        d_hvo = torch.empty_like(hvo) if ctx.needs_input_grad[0] else None
        partials = torch.empty(lib.gt_loss_scratch_floats(n), dtype=torch.float32, device=hvo.device)
        _lib.check(lib.gt_loss_voices(_lib.ptr(hvo), _lib.ptr(y), n, hvo.shape[2] // 3, float(penalty), _lib.ptr(metrics),
                                      _lib.ptr(d_hvo), 1.0, _lib.ptr(partials), _lib.stream_ptr(hvo.device)), "gt_loss")
        ctx.d_hvo = d_hvo
        ctx.mark_non_differentiable(metrics)
        return metrics[0].clone(), metrics

    @staticmethod
    def backward(ctx, g_loss, _g_metrics):
        d = ctx.d_hvo
        ctx.d_hvo = None
        return (d * g_loss if d is not None else None), None, None


def _packed(prediction):
    h, v, o = prediction
    base = getattr(h, "_groove_hvo", None)
    if base is not None and getattr(v, "_groove_hvo", None) is base and getattr(o, "_groove_hvo", None) is base:
        return base
    return torch.cat((h.float(), v.float(), o.float()), dim=2).contiguous()


def calculate_loss(prediction, y, bce_fn=None, mse_fn=None, hit_loss_penalty=1.0):
    """Same contract as the reference: ``prediction`` = (h logits, v, o); returns
    ``(total_loss 0-dim tensor with grad_fn, hit_accuracy, hit_perplexity, bce_hits, mse_velocities,
    mse_offsets)`` with the last five as Python floats.  ``bce_fn`` / ``mse_fn`` are accepted for
    signature compatibility (train.py:176-179 passes BCEWithLogitsLoss / MSELoss with reduction
    'none'); the fused kernel implements exactly those two losses and any other callable is rejected."""
    for fn, cls, what in ((bce_fn, torch.nn.BCEWithLogitsLoss, "bce_fn"), (mse_fn, torch.nn.MSELoss, "mse_fn")):
        # the kernel IS BCEWithLogitsLoss / MSELoss with reduction 'none' and no weights; anything else would be silently
        # replaced by them, so it is rejected (use the model's autograd path with your own loss function instead)
        if fn is not None and not (isinstance(fn, cls) and fn.reduction == "none" and getattr(fn, "weight", None) is None
                                   and getattr(fn, "pos_weight", None) is None):
            raise ValueError(f"groove_b200 calculate_loss implements {cls.__name__}(reduction='none') for {what} "
                             f"(train.py:176-179); got {fn!r}")
    hvo = _packed(prediction)
    if not hvo.is_cuda:
        raise RuntimeError("groove_b200 calculate_loss runs on CUDA only — there is no CPU fallback")
    if y.shape != hvo.shape or y.shape[1] != 32 or y.shape[2] % 3 != 0:
        raise ValueError(f"y must have shape {tuple(hvo.shape)} = [N, 32, 3 x voices], got {tuple(y.shape)}")
    y = y.to(hvo.device).contiguous().float()
    loss, metrics = _LossFn.apply(hvo.contiguous(), y, float(hit_loss_penalty))
    m = metrics.tolist()                      # ONE device->host read instead of the reference's five .item()
    return loss, m[1], m[2], m[3], m[4], m[5]


# ----------------------------------------------------------------------------------------------
# fused optimizers over the flat parameter vector (torch.optim.SGD / Adam drop-ins)
# ----------------------------------------------------------------------------------------------
class _FusedOptimizer:
    def __init__(self, model: _GrooveBase, lr: float):
        if not isinstance(model, _GrooveBase):
            raise TypeError("fused optimizers work on groove_b200 models")
        self.model = model
        self.grad_scale = 1.0          # set to 1/world by the data-parallel wrapper (all-reduce SUM)
        self.param_groups = [dict(self._defaults(lr), params=list(model.parameters()))]
        self.state = {}

    def zero_grad(self, set_to_none: bool = False):
        # the fused train step overwrites the gradient, the autograd path accumulates into it
        g = self.model._flat.grad
        if g is not None:
            g.zero_()

    def _lr(self):
        return float(self.param_groups[0]["lr"])

    def _packed_groups(self):
        grp = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        grp["params"] = list(range(len(self.model._views)))
        return [grp]

    def _load_groups(self, sd):
        pg = sd["param_groups"]
        if len(pg) != 1 or len(pg[0]["params"]) != len(self.model._views):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        for k, v in pg[0].items():
            if k != "params":
                self.param_groups[0][k] = v


class FusedSGD(_FusedOptimizer):
    """``torch.optim.SGD(params, lr)`` (momentum 0, weight decay 0 — BGT/models/train.py:65-66)."""

    @staticmethod
    def _defaults(lr):
        return dict(lr=lr, momentum=0, dampening=0, weight_decay=0, nesterov=False, maximize=False, foreach=None,
                    differentiable=False, fused=None)

    @torch.no_grad()
    def step(self):
        m = self.model
        g = m._flat.grad
        if g is None:
            return
        lib = _lib.load()
        _lib.check(lib.gt_sgd_step(_lib.ptr(m._flat), _lib.ptr(g), m._flat.numel(), self._lr(), float(self.grad_scale),
                                   _lib.stream_ptr(m._flat.device)), "gt_sgd_step")

    @torch.no_grad()
    def step_peers(self, bufs, world: int, offset_floats: int):
        """The same update with the gradient taken as the rank-ordered SUM over the ``world`` exchange buffers (dp.PeerExchange):
        gradient exchange + optimizer in one kernel over NVLink peer memory.  ``flat_grad()`` receives the sum."""
        m = self.model
        lib = _lib.load()
        _lib.check(lib.gt_sgd_step_peers(_lib.ptr(m._flat), bufs, world, offset_floats, _lib.ptr(m._flat.grad), m._flat.numel(),
                                         self._lr(), float(self.grad_scale), _lib.stream_ptr(m._flat.device)), "gt_sgd_step_peers")

    def state_dict(self):
        return {"state": {}, "param_groups": self._packed_groups()}

    def load_state_dict(self, sd):
        self._load_groups(sd)


class FusedAdam(_FusedOptimizer):
    """``torch.optim.Adam(params, lr)`` with torch defaults betas=(0.9,0.999), eps=1e-8, wd=0."""

    def __init__(self, model, lr):
        super().__init__(model, lr)
        self._m = torch.zeros_like(model._flat.detach())
        self._v = torch.zeros_like(model._flat.detach())
        self._t = 0

    @staticmethod
    def _defaults(lr):
        return dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                    capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)

    @torch.no_grad()
    def step(self):
        m = self.model
        g = m._flat.grad
        if g is None:
            return
        if self._m.device != m._flat.device:
            self._m, self._v = self._m.to(m._flat.device), self._v.to(m._flat.device)
        self._t += 1
        b1, b2 = self.param_groups[0]["betas"]
        lib = _lib.load()
        _lib.check(lib.gt_adam_step(_lib.ptr(m._flat), _lib.ptr(g), _lib.ptr(self._m), _lib.ptr(self._v), m._flat.numel(),
                                    self._lr(), float(b1), float(b2), float(self.param_groups[0]["eps"]), self._t,
                                    float(self.grad_scale), _lib.stream_ptr(m._flat.device)), "gt_adam_step")

    @torch.no_grad()
    def step_peers(self, bufs, world: int, offset_floats: int):
        """See FusedSGD.step_peers."""
        m = self.model
        if self._m.device != m._flat.device:
            self._m, self._v = self._m.to(m._flat.device), self._v.to(m._flat.device)
        self._t += 1
        b1, b2 = self.param_groups[0]["betas"]
        lib = _lib.load()
        _lib.check(lib.gt_adam_step_peers(_lib.ptr(m._flat), bufs, world, offset_floats, _lib.ptr(self._m), _lib.ptr(self._v),
                                          _lib.ptr(m._flat.grad), m._flat.numel(), self._lr(), float(b1), float(b2),
                                          float(self.param_groups[0]["eps"]), self._t, float(self.grad_scale),
                                          _lib.stream_ptr(m._flat.device)), "gt_adam_step_peers")

    def state_dict(self):
        state = {}
        if self._t > 0:
            for i, (p, (o, s)) in enumerate(zip(self.model._views, self.model._offsets)):
                state[i] = {"step": torch.tensor(float(self._t)),
                            "exp_avg": self._m[o:o + s].view(p.shape).clone(),
                            "exp_avg_sq": self._v[o:o + s].view(p.shape).clone()}
        return {"state": state, "param_groups": self._packed_groups()}

    def load_state_dict(self, sd):
        self._load_groups(sd)
        st = sd.get("state", {})
        if st:
            for i, (p, (o, s)) in enumerate(zip(self.model._views, self.model._offsets)):
                e = st[i] if i in st else st[str(i)]
                self._m[o:o + s].copy_(e["exp_avg"].reshape(-1))
                self._v[o:o + s].copy_(e["exp_avg_sq"].reshape(-1))
                self._t = int(float(e["step"]))


DROPOUT_STATE_KEY = "groove_b200_dropout_state"


# ----------------------------------------------------------------------------------------------
# initialize_model  (BGT/models/train.py:43-108)
# ----------------------------------------------------------------------------------------------
def initialize_model(params):
    """Same ``params`` schema as train.py:115-143 -> ``(model, optimizer, epoch)``.

    Extra optional keys (ignored by the reference): ``params['model']['precision']`` in
    {'fp32','bf16'}; ``params['training']['fused_optimizer']`` (default True) selects FusedSGD /
    FusedAdam instead of torch.optim (both work: parameters are real nn.Parameters)."""
    mp, tp, load_model = params["model"], params["training"], params["load_model"]
    if mp["encoder_only"]:
        model = GrooveTransformerEncoder(mp["d_model"], mp["embedding_size_src"], mp["embedding_size_tgt"], mp["n_heads"],
                                         mp["dim_feedforward"], mp["dropout"], mp["num_encoder_layers"], mp["max_len"],
                                         mp["device"])
    else:
        model = GrooveTransformer(mp["d_model"], mp["embedding_size_src"], mp["embedding_size_tgt"], mp["n_heads"],
                                  mp["dim_feedforward"], mp["dropout"], mp["num_encoder_layers"], mp["num_decoder_layers"],
                                  mp["max_len"], mp["device"])
    model.to(mp["device"])
    if "precision" in mp:
        model.set_precision(mp["precision"])
    lr = tp["learning_rate"]
    adam = mp["optimizer"] == "adam"
    if tp.get("fused_optimizer", True):
        optimizer = FusedAdam(model, lr) if adam else FusedSGD(model, lr)
    else:
        optimizer = torch.optim.Adam(model.parameters(), lr=lr) if adam else torch.optim.SGD(model.parameters(), lr=lr)
    epoch = 0

    if load_model is not None:
        checkpoint = None
        if load_model["location"] == "local":
            # highest number found in the names of files with the pattern's extension; the reference
            # only accepts checkpoint numbers > 0 (train.py:88,93) and then fails with a NameError —
            # here the same situation raises FileNotFoundError.
            ext = re.findall(r"\w+", load_model["file_pattern"])[-1]
            best, best_name = 0, None
            for root, _dirs, files in os.walk(load_model["dir"]):
                for name in files:
                    if name.endswith(ext):
                        digits = re.findall(r"\d+", name)
                        if digits and int(digits[-1]) > best:
                            best, best_name = int(digits[-1]), os.path.join(load_model["dir"], name)
            if best_name is None:
                raise FileNotFoundError(f"no checkpoint with epoch > 0 and extension '{ext}' in {load_model['dir']}")
            checkpoint = torch.load(best_name, map_location="cpu", weights_only=False)
        elif load_model["location"] == "wandb":
            if wandb is None:
                raise RuntimeError("wandb is not installed: cannot restore a checkpoint from wandb")
            f = wandb.restore(load_model["file_pattern"].format(load_model["run"], load_model["epoch"]),
                              run_path=load_model["dir"])
            checkpoint = torch.load(f.name, map_location="cpu", weights_only=False)
        else:
            raise ValueError("load_model['location'] must be 'local' or 'wandb'")
        model.load_state_dict(checkpoint["model_state_dict"])
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        epoch = checkpoint["epoch"]
        ds = checkpoint.get(DROPOUT_STATE_KEY)          # absent in reference checkpoints: a fresh stream, like the reference
        if ds is not None:
            model.set_seed(ds["seed"], ds["step"], model._seq0)
    return model, optimizer, epoch


# ----------------------------------------------------------------------------------------------
# train_loop  (BGT/models/train.py:111-207)
# ----------------------------------------------------------------------------------------------
def _wandb_active():
    return wandb is not None and getattr(wandb, "run", None) is not None


def _log(d):
    if _wandb_active():
        wandb.log(d, commit=True)


def _shift(y):
    return torch.cat((torch.zeros_like(y[:, :1]), y[:, :-1]), dim=1)


def _eval_pass(model, loss_fn, bce_fn, mse_fn, inputs, gt, device, encoder_only, penalty, prefix, epoch):
    inputs, gt = inputs.to(device), gt.to(device)
    model.eval()
    with torch.no_grad():
        pred = model(inputs) if encoder_only else model(inputs, _shift(gt))
        loss, acc, ppl, bce, mv, mo = loss_fn(pred, gt, bce_fn, mse_fn, penalty)
    _log({f"{prefix}_loss": loss.item(), f"{prefix}_hit_accuracy": acc, f"{prefix}_hit_perplexity": ppl,
          f"{prefix}_hit_loss": bce, f"{prefix}_velocity_loss": mv, f"{prefix}_offset_loss": mo, "epoch": epoch})
    return loss.item()


def train_loop(dataloader, groove_transformer, loss_fn, bce_fn, mse_fn, opt, epoch, save, device, encoder_only,
               hit_loss_penalty=1, test_inputs=None, test_gt=None, validation_inputs=None, validation_gt=None):
    """One epoch with the reference's per-batch order (zero_grad, H2D, forward, loss, backward, step,
    wandb.log of the same eight keys), checkpoint format and test / validation passes.

    When ``loss_fn`` is this package's ``calculate_loss`` and ``opt`` is a fused optimizer the batch
    body is ONE library call (``gt_train_step``) plus the optimizer kernel and a single 6-float
    device->host read; otherwise the generic autograd path runs (any loss_fn / torch optimizer)."""
    size = len(dataloader.dataset)
    model = groove_transformer
    model.train()
    fused = loss_fn is calculate_loss and isinstance(opt, _FusedOptimizer) and isinstance(model, _GrooveBase)
    loss_value = 0.0
    for batch, (x, y, idx) in enumerate(dataloader):
        opt.zero_grad()
        x = x.to(device, non_blocking=True)
        y = y.to(device, non_blocking=True)
        if fused:
            metrics, _ = model.train_step(x, y, hit_loss_penalty)
            opt.step()
            loss_value, acc, ppl, bce_h, mse_v, mse_o = metrics.tolist()
        else:
            pred = model(x) if encoder_only else model(x, _shift(y))
            loss, acc, ppl, bce_h, mse_v, mse_o = loss_fn(pred, y, bce_fn, mse_fn, hit_loss_penalty)
            loss.backward()
            opt.step()
            loss_value = loss.item()
        _log({"train_loss": loss_value, "train_hit_accuracy": acc, "train_hit_perplexity": ppl, "train_hit_loss": bce_h,
              "train_velocity_loss": mse_v, "train_offset_loss": mse_o, "epoch": epoch, "batch": batch})
        if batch % 100 == 0:
            print("=======")
            print(f"loss: {loss_value:>4f}  [{batch * len(x):>4d}/{size:>4d}]")
            print("hit accuracy:", np.round(acc, 4))
            print("hit perplexity: ", np.round(ppl, 4))
            print("hit bce: ", np.round(bce_h, 4))
            print("velocity mse: ", np.round(mse_v, 4))
            print("offset mse: ", np.round(mse_o, 4))

    if save:
        run_dir = wandb.run.dir if _wandb_active() else os.environ.get("GROOVE_CKPT_DIR", ".")
        run_id = wandb.run.id if _wandb_active() else "local"
        fn = os.path.join(run_dir, "transformer_run_{}_Epoch_{}.Model".format(run_id, epoch))
        ckpt = {"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "loss": loss_value}
        if isinstance(model, _GrooveBase):             # extra key the reference's loader ignores: resume continues the mask stream
            ckpt[DROPOUT_STATE_KEY] = {"seed": model._seed, "step": model._step}
        torch.save(ckpt, fn)
        if _wandb_active():
            wandb.save(fn, base_path=wandb.run.dir)

    if test_inputs is not None and test_gt is not None:
        _eval_pass(model, loss_fn, bce_fn, mse_fn, test_inputs, test_gt, device, encoder_only, hit_loss_penalty, "test", epoch)
    if validation_inputs is not None and validation_gt is not None:
        _eval_pass(model, loss_fn, bce_fn, mse_fn, validation_inputs, validation_gt, device, encoder_only, hit_loss_penalty,
                   "validation", epoch)
    return loss_value
