"""CPU restatement of the split-operand contraction behind ``precision = 'fp32_tc'`` (GT_PREC_FP32_TC,
transformergrooveinfilling_b200/csrc/gemm_tc.cu: split_pair / the six passes of the pre-imaged K loop).  Test infrastructure only.

The reference computes every Linear in fp32 (torch.nn.Linear inside BGT/models/io_layers.py:17-22, encoder.py:8-10, decoder.py:8-10,
on whatever BLAS the device has).  The tensor cores take bf16 operands, so the kernel splits each fp32 operand EXACTLY into three
bf16 terms and contracts the six products of weight >= 2^-18; this file restates that arithmetic with torch CPU ops so that the
claim "fp32 results" is pinned without a GPU:

    x0 = bf16(x)            x1 = bf16(x - x0)            x2 = bf16(x - x0 - x1)          (the subtractions are exact in fp32)
    C  = [x0.y0]  +  [x2.y0 + x0.y2 + x1.y1 + x1.y0 + x0.y1]                             (main accumulator + correction accumulator)

What the restatement does NOT model is the tensor core's truncating fp32 accumulation (one ulp towards zero per UMMA); that is a
measured property of the hardware (tests/test_gpu_fp32_tc.py) and the reason for the second accumulator."""
import torch


def bf16(x: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to bf16, kept in the input dtype (like __floats2bfloat162_rn + shift back)."""
    return x.to(torch.bfloat16).to(x.dtype)


def split3(x: torch.Tensor):
    """gemm_tc.cu:split_pair — three bf16 terms whose sum is x exactly (8 + 8 + 8 significand bits)."""
    x = x.float()
    x0 = bf16(x)
    r1 = x - x0
    x1 = bf16(r1)
    r2 = r1 - x1
    x2 = bf16(r2)
    return x0, x1, x2


def split2(x: torch.Tensor):
    """the two-term split that was measured first and rejected (three products, 2^-16)."""
    x = x.float()
    x0 = bf16(x)
    return x0, bf16(x - x0)


# (A image, B image) of pass p, smallest products first — gemm_tc.cu: ia = 0x001102 >> 4p, ib = 0x010120 >> 4p
PASSES = [(2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)]


def split_matmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """C[m, n] = sum_k a[m, k] b[n, k] the way the split-mode kernel computes it: bf16 x bf16 products are exact in fp32, the main
    accumulator takes only the x0.y0 pass, the five corrections share a second fp32 accumulator, the epilogue adds the two."""
    sa, sb = split3(a), split3(b)
    main = torch.zeros(a.shape[0], b.shape[0], dtype=torch.float32)
    corr = torch.zeros_like(main)
    for ia, ib in PASSES:
        prod = sa[ia] @ sb[ib].T                     # fp32 accumulate of exact bf16 x bf16 products
        if (ia, ib) == (0, 0):
            main += prod
        else:
            corr += prod
    return main + corr


def split2_matmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    (a0, a1), (b0, b1) = split2(a), split2(b)
    return a0 @ b0.T + (a1 @ b0.T + a0 @ b1.T)
