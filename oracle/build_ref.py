"""Recipe for oracle/_ref: the UNMODIFIED reference hot path, made available to the GPU box.

    python oracle/build_ref.py        (run by __graft_entry__.build() whenever /root/reference exists)

The reference is pure Python on top of torch (BGT = BaseGrooveTransformers: models/{io_layers,encoder,decoder,transformer,
utils,train}.py, 488 lines), so "building" it means copying its package tree, byte for byte, from where it lies under
/root/reference into oracle/_ref/ — git-ignored (reference sources never enter this repository's history) but NOT
gpurun-ignored, so the tree travels to the GPU box exactly like the built .so does.  bench.py's reference arm and
cpu_baseline leg then time THE reference (BGT/models/train.py:118-141 body: zero_grad -> forward -> calculate_loss ->
backward -> step) instead of the oracle port; without oracle/_ref they fall back to the port and say so
(cpu_baseline.kind "port").  A manifest with the sha256 of every copied file is written next to the tree so a reader
can check that nothing was edited.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/BaseGrooveTransformers"
DST = os.path.join(HERE, "_ref", "BaseGrooveTransformers")
FILES = ["__init__.py", "models/__init__.py", "models/io_layers.py", "models/encoder.py", "models/decoder.py",
         "models/transformer.py", "models/utils.py", "models/train.py"]


def build_ref(verbose: bool = False) -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)                 # GPU box: use what travelled with the snapshot
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": SRC, "sha256": manifest}, open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} files copied from {SRC}")
    return True


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    sys.exit(0 if ok else 1)
