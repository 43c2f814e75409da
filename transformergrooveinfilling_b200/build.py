"""Builds libgroove_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension
machinery: the library has a plain C ABI and is loaded with ctypes)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgroove_b200.so")
STAMP = LIB + ".stamp"
SOURCES = ["kernels_simt.cu", "peer_opt.cu", "edge32.cu", "edge256.cu", "decode32.cu", "attn_mma.cu", "gemm_tc.cu", "runner.cu", "tc_path.cu", "tc_layers.cu", "tc256.cu", "tc256_path.cu", "tc256_bwd.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]

# Developer builds (e.g. GROOVE_B200_DEV_FLAGS="-DGT_T256_TIMELINE" for the clock64 phase time line of tc256*.cu) go to their own
# library / object directory, so the release library is never replaced by one that allocates its own debug buffers.
_DEV = os.environ.get("GROOVE_B200_DEV_FLAGS", "").split()
if _DEV:
    LIB = os.path.join(HERE, "libgroove_b200_dev.so")
    STAMP = LIB + ".stamp"
    NVCC_FLAGS = NVCC_FLAGS + _DEV


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the groove_b200 CUDA library cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), "include", "groove_b200.h")
    for p in sorted(os.listdir(CSRC)) + [inc]:
        fp = p if os.path.isabs(p) else os.path.join(CSRC, p)
        if fp.endswith((".cu", ".cuh", ".h")):
            h.update(open(fp, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    bdir = os.path.join(HERE, "build_dev" if _DEV else "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    open(STAMP, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
