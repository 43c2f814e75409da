"""Evaluator metrics kernel (gt_eval_metrics) + its host-side mirror (HVOMetrics) against the oracle restatement of
GrooveEvaluator/GrooveEvaluator/evaluator.py:189-251 and the golden outputs of the reference methods themselves."""
import os

import numpy as np
import pytest
import torch

import eval_oracle as E
import groove_oracle as G
from _util import build_model
from transformergrooveinfilling_b200 import HVOMetrics, hvo_metrics_vector

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAPPING = {v: [i] for i, v in enumerate(E.ROLAND_REDUCED_VOICES)}


@pytest.mark.parametrize("n", [1, 7, 64])
def test_against_reference_golden(n):
    g = np.load(os.path.join(GOLD, "eval_metrics.npz"))[f"n{n}"]
    gt, pr = E.det_eval_arrays(n)
    vec = hvo_metrics_vector(torch.from_numpy(pr).cuda(), torch.from_numpy(gt).cuda()).cpu().numpy()
    # hit accuracies are ratios of small integers: exact in fp32 up to the final rounding; MSEs: fp32 summation order
    np.testing.assert_allclose(vec[:10], g[:10], rtol=1e-6)
    np.testing.assert_allclose(vec[10:], g[10:], rtol=2e-5)


@pytest.mark.parametrize("n", [3, 333, 40000])
def test_dictionaries_match_oracle(n):
    gt, pr = E.det_eval_arrays(n) if n < 1000 else _big(n)
    ev = HVOMetrics(gt, _identifier="Test", device="cuda").add_predictions(pr)
    ref = E.eval_metrics(gt, pr, identifier="Test")
    for name, fn in (("Hits_Accuracy", ev.get_hits_accuracies), ("Velocity_MSE", ev.get_velocity_errors),
                     ("Micro_Timing_MSE", ev.get_micro_timing_errors)):
        got = fn(MAPPING)
        assert list(got.keys()) == [name] and list(got[name].keys()) == ["Test"]
        assert list(got[name]["Test"].keys()) == list(ref[name]["Test"].keys())
        for k, v in ref[name]["Test"].items():
            assert abs(got[name]["Test"][k] - v) <= 3e-5 * max(abs(v), 1e-3), (name, k)


def _big(n):
    rng = np.random.default_rng(5)
    hits = (rng.random((n, 32, 9)) < 0.15).astype(np.float32)
    gt = np.concatenate((hits, rng.random((n, 32, 9), dtype=np.float32) * hits, (rng.random((n, 32, 9), dtype=np.float32) - 0.5) * hits), 2)
    ph = (rng.random((n, 32, 9)) < 0.2).astype(np.float32)
    pr = np.concatenate((ph, rng.random((n, 32, 9), dtype=np.float32), rng.random((n, 32, 9), dtype=np.float32) - 0.5), 2)
    return gt, pr.astype(np.float32)


def test_predict_then_metrics_stays_on_device():
    """evaluator.py:171-186 followed by :189-251: model.predict output goes straight into the metric kernel."""
    cfg = G.GrooveCfg(32, 4, 16, 2, 0, 16, 27)
    model, P = build_model(cfg, dropout=0.0)
    x, y = G.det_batch(cfg, 50)
    ev = HVOMetrics(y, device="cuda")
    pred = ev.predict(model, x)
    assert pred.is_cuda and tuple(pred.shape) == (50, 32, 27)
    h, v, o = G.predict_encoder_only(P, cfg, x)
    ref = E.eval_metrics(y.numpy(), np.concatenate((h.numpy().astype(np.float32), v.numpy(), o.numpy()), 2))
    got = ev.get_hits_accuracies(MAPPING)["Hits_Accuracy"]["Train"]
    for k, val in ref["Hits_Accuracy"]["Train"].items():
        assert abs(got[k] - val) < 2e-3          # a logit within fp32 noise of the threshold may flip one of 1600 cells
    gv = ev.get_velocity_errors(MAPPING)["Velocity_MSE"]["Train"]["Overall"]
    assert abs(gv - ref["Velocity_MSE"]["Train"]["Overall"]) < 1e-4


def test_rejects_cpu_and_bad_shapes():
    gt, pr = E.det_eval_arrays(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        hvo_metrics_vector(torch.from_numpy(pr), torch.from_numpy(gt))
    with pytest.raises(ValueError):
        hvo_metrics_vector(torch.from_numpy(pr).cuda()[:, :16], torch.from_numpy(gt).cuda())
