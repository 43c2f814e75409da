"""tcgen05 / TMEM engine unit test: the stand-alone tile GEMM against torch fp32 on bf16-rounded
operands.  Runs each variant in a SUBPROCESS with a timeout so that a wrong descriptor (illegal
instruction / trap) cannot take the whole test session down."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from transformergrooveinfilling_b200 import _lib
lib = _lib.load()
variant, m, n, k = map(int, sys.argv[1:5])
torch.manual_seed(0)
a = torch.randn(m, k, device="cuda").bfloat16()
b = torch.randn(n, k, device="cuda").bfloat16()
d = torch.zeros(m, n, device="cuda")
ag = a.t().contiguous() if variant & 4 else a      # bit2: A handed over transposed -> MN-major operand
bg = b.t().contiguous() if variant & 8 else b      # bit3: B handed over transposed -> MN-major operand
_lib.check(lib.gt_debug_tc_gemm(ag.data_ptr(), bg.data_ptr(), d.data_ptr(), m, n, k, variant, 0), "tc_gemm")
torch.cuda.synchronize()
ref = a.float() @ b.float().T
err = (d - ref).abs().max().item() / ref.abs().max().item()
print("RELERR", err)
sys.exit(0 if err < 1e-5 else 3)
""" % ROOT


def _run(variant, m, n, k):
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(variant), str(m), str(n), str(k)], capture_output=True, text=True, timeout=180)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.parametrize("m,n,k", [(128, 32, 32), (256, 96, 32), (128, 256, 64), (384, 64, 128), (128, 16, 16), (128, 48, 256)])
def test_tile_gemm(m, n, k):
    rc, out = _run(0, m, n, k)
    if rc != 0:
        rc1, out1 = _run(1, m, n, k)
        pytest.fail(f"variant0 rc={rc}: {out[-400:]}\n--- variant1 (LBO/SBO swapped) rc={rc1}: {out1[-400:]}")


@pytest.mark.parametrize("variant", [4, 8, 12])
@pytest.mark.parametrize("m,n,k", [(128, 32, 128), (256, 48, 64), (128, 96, 32)])
def test_tile_gemm_mn_major_operands(variant, m, n, k):
    """A and/or B consumed through MN-major descriptors (what the fused backward kernel uses to
    contract over the token dimension without transposing anything in shared memory)."""
    rc, out = _run(variant, m, n, k)
    if rc != 0:
        rc1, out1 = _run(variant | 1, m, n, k)
        pytest.fail(f"variant{variant} rc={rc}: {out[-300:]}\n--- with LBO/SBO swapped rc={rc1}: {out1[-300:]}")
