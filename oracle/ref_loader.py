"""Loads the unmodified reference package (oracle/_ref, see build_ref.py) under the private module name ``_groove_reference``
so that it can coexist with this repository's drop-in alias package ``BaseGrooveTransformers`` in one process.
TEST / BENCH INFRASTRUCTURE ONLY — the product never imports this."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NAME = "_groove_reference"


def reference_dir():
    for cand in (os.path.join(HERE, "_ref", "BaseGrooveTransformers"), "/root/reference/BaseGrooveTransformers"):
        if os.path.isfile(os.path.join(cand, "models", "transformer.py")):
            return cand
    return None


def load_reference():
    """-> (package module, directory) or (None, reason).  The reference imports wandb at module import (BGT/models/train.py:3);
    it is only used for logging, which the timed body (train.py:118-141 minus wandb.log) never calls."""
    if NAME in sys.modules:
        return sys.modules[NAME], os.path.dirname(sys.modules[NAME].__file__)
    d = reference_dir()
    if d is None:
        return None, "oracle/_ref is absent (run oracle/build_ref.py where /root/reference exists)"
    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception as e:  # pragma: no cover
        del sys.modules[NAME]
        return None, f"reference import failed: {e!r}"
    return mod, d
