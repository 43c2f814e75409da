"""Per-shape throughput of the generic tcgen05 GEMM (gt_debug_gemm, tc=1) at the contraction shapes of one layer:
python tools/gemm_tc_bench.py [d_model] [dim_ff] [n_seq].  Prints TFLOP/s and the fp32 HBM bytes moved (A + C) / time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformergrooveinfilling_b200 import _lib  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n_seq = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
M = n_seq * 32
lib = _lib.load()
dev = "cuda"


def run(name, a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, flags=0, split=0, bytes_moved=0):
    def call():
        _lib.check(lib.gt_debug_gemm(1, a.data_ptr(), sam, sak, b.data_ptr(), sbn, sbk, c.data_ptr(), ldc, m, n, k, flags, 0, 0, 0, 0, 0,
                                     1.0, 0.0, 0, 0, 0, 0, split, 0), name)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:34s} M={m:7d} N={n:4d} K={k:7d}  {ms*1e3:8.1f} us  {2*m*n*k/ms/1e9:7.1f} TFLOP/s  {bytes_moved/ms/1e6:7.1f} GB/s")


for (nm, N, K) in (("linear qkv", 3 * d, d), ("linear out", d, d), ("linear ffn1", F, d), ("linear ffn2", d, F)):
    x, w, c = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.empty(M, N, device=dev)
    run(nm, x, K, 1, w, K, 1, c, N, M, N, K, bytes_moved=4 * (M * K + M * N))
for (nm, N, K) in (("dgrad qkv", d, 3 * d), ("dgrad ffn2 (dH)", F, d), ("dgrad ffn1", d, F)):
    dy, w, c = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev), torch.empty(M, N, device=dev)
    run(nm, dy, K, 1, w, 1, N, c, N, M, N, K, bytes_moved=4 * (M * K + M * N))
for (nm, No, Ni) in (("wgrad qkv", 3 * d, d), ("wgrad out", d, d), ("wgrad ffn1", F, d), ("wgrad ffn2", d, F)):
    dy, x, dw = torch.randn(M, No, device=dev), torch.randn(M, Ni, device=dev), torch.zeros(No, Ni, device=dev)
    run(nm, dy, 1, No, x, 1, Ni, dw, Ni, No, Ni, M, flags=4, split=2048, bytes_moved=4 * (M * No + M * Ni))
