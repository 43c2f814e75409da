"""Multi-GPU equivalence test (NCCL all-reduce form and the peer-memory exchange) — runs only when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_step_equals_single_gpu():
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DP_CHECK fp32" in r.stdout and "DP_CHECK bf16" in r.stdout
    # exchange = 'p2p' (gradient sum over NVLink peer memory inside the optimizer kernel, csrc/peer_opt.cu): SGD and Adam, 4 steps,
    # against the single-GPU run and the NCCL form, replicas bit-identical
    assert r.stdout.count("DP_CHECK p2p") == 3 and "replicas_bit_identical=False" not in r.stdout, r.stdout[-2000:]
