"""Per-voice evaluator metrics on the GPU — the step right after ``predict()`` in the reference's per-epoch evaluation.

Reference: GrooveEvaluator/GrooveEvaluator/evaluator.py.  ``train.py`` calls ``evaluator.predict`` (evaluator.py:171-186:
``model.predict(inputs, use_thres=True, thres=0.5)`` -> three tensors ``.cpu()``-ed and ``np.concatenate(axis=2)``-ed
into ``_prediction_hvos_array`` [N,32,27]) and then ``get_hits_accuracies`` / ``get_velocity_errors`` /
``get_micro_timing_errors`` (evaluator.py:189-251) on the host with numpy.  Here the prediction never leaves the
device: ``HVOMetrics`` holds the ground truth in HBM, ``add_predictions`` takes the model's output tensors as they are,
and one library call (``gt_eval_metrics``) reduces both arrays to the 3 x (V + 1) numbers, read back in one 120-byte
copy.  Method names, argument (``drum_mapping``) and the nesting / keys of the returned dictionaries are the reference's.

There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import torch

from . import _lib

_KEYS = ("Hits_Accuracy", "Velocity_MSE", "Micro_Timing_MSE")


def hvo_metrics_vector(pred_hvo: torch.Tensor, gt_hvo: torch.Tensor, n_voices: int = 9) -> torch.Tensor:
    """Device tensor of 3 * (n_voices + 1) floats: {hit accuracy, velocity MSE, micro-timing MSE} x {voices..., Overall}."""
    for t, what in ((pred_hvo, "pred_hvo"), (gt_hvo, "gt_hvo")):
        if not isinstance(t, torch.Tensor) or t.dim() != 3 or t.shape[1] != _lib.T_STEPS or t.shape[2] != 3 * n_voices:
            raise ValueError(f"{what} must have shape [N, {_lib.T_STEPS}, {3 * n_voices}], got {tuple(getattr(t, 'shape', ()))}")
        if not t.is_cuda:
            raise RuntimeError("groove_b200 evaluator metrics run on CUDA only — there is no CPU fallback")
    if pred_hvo.shape[0] != gt_hvo.shape[0] or pred_hvo.shape[0] == 0:
        raise ValueError("prediction / ground-truth batch sizes differ or are empty")
    if pred_hvo.device != gt_hvo.device:
        raise RuntimeError("prediction and ground truth are on different devices")
    lib = _lib.load()
    pred = pred_hvo.contiguous().float()
    gt = gt_hvo.contiguous().float()
    n = pred.shape[0]
    nscr = lib.gt_eval_scratch_floats(n, n_voices)
    if nscr < 0:
        raise RuntimeError(lib.gt_last_error().decode())
    scratch = torch.empty(nscr, dtype=torch.float32, device=pred.device)
    out = torch.empty(3 * (n_voices + 1), dtype=torch.float32, device=pred.device)
    _lib.check(lib.gt_eval_metrics(_lib.ptr(pred), _lib.ptr(gt), n, n_voices, _lib.ptr(out), _lib.ptr(scratch),
                                   _lib.stream_ptr(pred.device)), "gt_eval_metrics")
    return out


class HVOMetrics:
    """The metric half of the reference's ``Evaluator`` (evaluator.py:28-339): ground truth [N,32,27] kept on the device,
    predictions added per epoch, the three metric dictionaries computed by one kernel pass."""

    def __init__(self, gt_hvos_array, _identifier: str = "Train", device=None):
        gt = torch.as_tensor(gt_hvos_array, dtype=torch.float32)
        if device is not None:
            gt = gt.to(device)
        if not gt.is_cuda:
            raise RuntimeError("HVOMetrics keeps the ground truth on a CUDA device — pass device='cuda'")
        self._gt_hvos_array = gt.contiguous()
        self._identifier = _identifier
        self._prediction_hvos_array = None
        self._vec = None

    # evaluator.py:171-186 (predict) + :299-303 (add_predictions)
    def add_predictions(self, prediction_hvos_array):
        """Accepts the [N,32,27] array of the reference API or the (h, v, o) tuple ``model.predict`` returns."""
        if isinstance(prediction_hvos_array, (tuple, list)):
            h, v, o = prediction_hvos_array
            if getattr(v, "_groove_hvo", None) is not None and h.dtype == torch.float32:
                pred = v._groove_hvo                      # the packed [N,32,27] tensor the three views share
            else:
                pred = torch.cat((h.float(), v.float(), o.float()), dim=2)
        else:
            pred = torch.as_tensor(prediction_hvos_array, dtype=torch.float32)
        self._prediction_hvos_array = pred.to(self._gt_hvos_array.device).contiguous()
        self._vec = None
        return self

    def predict(self, model, inputs, use_thres=True, thres=0.5):
        """evaluator.py:171-186 with the prediction left on the device."""
        self.add_predictions(model.predict(inputs.to(self._gt_hvos_array.device), use_thres=use_thres, thres=thres))
        return self._prediction_hvos_array

    def _vector(self, n_voices):
        if self._prediction_hvos_array is None:
            raise RuntimeError("no predictions: call add_predictions() / predict() first")
        if self._vec is None or self._vec[0] != n_voices:
            v = hvo_metrics_vector(self._prediction_hvos_array, self._gt_hvos_array, n_voices)
            self._vec = (n_voices, v.cpu().tolist())         # ONE device -> host read for all three dictionaries
        return self._vec[1]

    def _dict(self, which, drum_mapping):
        voices = list(drum_mapping.keys())
        V = len(voices)
        vec = self._vector(V)
        base = _KEYS.index(which) * (V + 1)
        inner = {str(v): vec[base + i] for i, v in enumerate(voices)}
        inner["Overall"] = vec[base + V]
        return {which: {self._identifier: inner}}

    def get_hits_accuracies(self, drum_mapping):          # evaluator.py:189-209
        return self._dict("Hits_Accuracy", drum_mapping)

    def get_velocity_errors(self, drum_mapping):          # evaluator.py:211-231
        return self._dict("Velocity_MSE", drum_mapping)

    def get_micro_timing_errors(self, drum_mapping):      # evaluator.py:233-251
        return self._dict("Micro_Timing_MSE", drum_mapping)
