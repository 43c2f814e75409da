"""CPU restatement (numpy) of the reference's per-voice evaluator metrics — TEST INFRASTRUCTURE ONLY.

Follows GrooveEvaluator/GrooveEvaluator/evaluator.py:189-251 (Evaluator.get_hits_accuracies / get_velocity_errors /
get_micro_timing_errors) on the arrays evaluator.py:171-186 builds: ``_gt_hvos_array`` and ``_prediction_hvos_array`` of shape
[n_examples, 32, 3 V] (hits | velocities | offsets), the prediction being np.concatenate(model.predict(...), axis=2).

Pinned: tests/golden/eval_metrics.npz holds the outputs of the UNMODIFIED reference methods (oracle/make_golden_eval.py
executes their source text from /root/reference) on the deterministic arrays of ``det_eval_arrays``;
tests/test_oracle_golden.py checks this restatement against them.  Only tests/ and __graft_entry__.smoke() may import it.
"""
import numpy as np

import groove_oracle as G

# hvo_sequence/hvo_sequence/drum_mappings.py:54-64 (keys only: the evaluator uses the voice names and their order)
ROLAND_REDUCED_VOICES = ("KICK", "SNARE", "HH_CLOSED", "HH_OPEN", "TOM_3_LO", "TOM_2_MID", "TOM_1_HI", "CRASH", "RIDE")


def det_eval_arrays(n: int, tag: int = 31):
    """Deterministic (ground truth, prediction) pair: two independent HVO draws; half of the prediction's hit cells are copied
    from the ground truth so that accuracies are neither 0 nor 1."""
    cfg = G.GrooveCfg(32, 4, 16, 1, 0, 27, 27)
    _, gt = G.det_batch(cfg, n, tag=tag)
    _, pr = G.det_batch(cfg, n, tag=tag + 1)
    gt, pr = gt.numpy().copy(), pr.numpy().copy()
    pr[:, ::2, :] = np.where(np.arange(27)[None, None, :] < 9, gt[:, ::2, :], pr[:, ::2, :])
    return gt.astype(np.float32), pr.astype(np.float32)


def eval_metrics(gt: np.ndarray, pred: np.ndarray, voices=ROLAND_REDUCED_VOICES, identifier="Train"):
    """The three dictionaries of evaluator.py:189-251, same nesting and keys."""
    V = len(voices)
    n = gt.shape[0]
    out = {}
    # evaluator.py:189-209
    g, p = gt[:, :, :V], pred[:, :, :V]
    acc = {v: ((g[:, :, i] == p[:, :, i]).sum(axis=-1) / g.shape[1]).mean() for i, v in enumerate(voices)}
    gf, pf = g.reshape((n, -1)), p.reshape((n, -1))
    acc["Overall"] = ((gf == pf).sum(axis=-1) / gf.shape[-1]).mean()
    out["Hits_Accuracy"] = {identifier: acc}
    # evaluator.py:211-231 and :233-251
    for name, lo in (("Velocity_MSE", V), ("Micro_Timing_MSE", 2 * V)):
        g, p = gt[:, :, lo:lo + V], pred[:, :, lo:lo + V]
        err = {v: (((g[:, :, i] - p[:, :, i]) ** 2).mean(axis=-1)).mean() for i, v in enumerate(voices)}
        gf, pf = g.reshape((n, -1)), p.reshape((n, -1))
        err["Overall"] = (((gf - pf) ** 2).mean(axis=-1)).mean()
        out[name] = {identifier: err}
    return out


def as_vector(metrics: dict, voices=ROLAND_REDUCED_VOICES, identifier="Train") -> np.ndarray:
    """[hits | velocity | micro-timing] x (voices..., Overall) — the layout gt_eval_metrics writes."""
    return np.array([metrics[k][identifier][v] for k in ("Hits_Accuracy", "Velocity_MSE", "Micro_Timing_MSE")
                     for v in tuple(voices) + ("Overall",)], dtype=np.float64)
