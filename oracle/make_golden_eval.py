"""Generate tests/golden/eval_metrics.npz from the UNMODIFIED reference evaluator methods.

GrooveEvaluator's package imports note_seq / bokeh (absent here), so the module cannot be imported; the three metric methods
(GrooveEvaluator/GrooveEvaluator/evaluator.py:189-251) only use numpy and three attributes of ``self``.  This script reads the
reference source file where it lies, extracts those three FunctionDefs with ``ast`` and executes their unmodified text
against a stand-in ``self`` carrying ``_gt_hvos_array`` / ``_prediction_hvos_array`` / ``_identifier``.
Run in the build container only:  python -B oracle/make_golden_eval.py
"""
import ast
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import eval_oracle as E  # noqa: E402

SRC = "/root/reference/GrooveEvaluator/GrooveEvaluator/evaluator.py"
WANTED = ("get_hits_accuracies", "get_velocity_errors", "get_micro_timing_errors")


def reference_methods():
    text = open(SRC).read()
    tree = ast.parse(text)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Evaluator")
    fns = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, SRC, "exec"), ns)
            fns[node.name] = ns[node.name]
    assert set(fns) == set(WANTED)
    return fns


def main():
    fns = reference_methods()
    mapping = {v: [i] for i, v in enumerate(E.ROLAND_REDUCED_VOICES)}
    out = {}
    for n in (1, 7, 64):
        gt, pr = E.det_eval_arrays(n)
        self = types.SimpleNamespace(_gt_hvos_array=gt, _prediction_hvos_array=pr, _identifier="Train")
        res = {}
        for name in WANTED:
            res.update(fns[name](self, mapping))
        out[f"n{n}"] = E.as_vector(res)
    np.savez(os.path.join(os.path.dirname(HERE), "tests", "golden", "eval_metrics.npz"), **out)
    for k, v in out.items():
        print(k, np.round(v, 5))


if __name__ == "__main__":
    main()
