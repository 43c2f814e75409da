"""Clocks per 128 x N x 16 bf16 UMMA from shared memory (canonical no-swizzle operands) for different issue patterns:
blocks of `ks` back-to-back UMMAs; mode bit0 = tcgen05.commit per block, bit1 = mbarrier wait per block, bit2 = one-thread warp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformergrooveinfilling_b200 import _lib
lib = _lib.load()
out = torch.zeros(1, device="cuda")
for n in (64, 192, 256):
    for mode in (0, 1, 3, 4, 5, 7):
        row = []
        for ks in (1, 2, 4, 8, 16):
            lib.gt_debug_umma_rate(n, 3840, ks | (mode << 8), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            row.append(f"{float(out[0]):6.1f}")
        print(f"N={n:3d} mode={mode} (floor {128*n//256:3d}): ks=1,2,4,8,16 ->", " ".join(row))
