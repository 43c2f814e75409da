"""CPU: the arithmetic of precision = 'fp32_tc' (oracle/split_gemm_oracle.py restates gemm_tc.cu's split mode) gives fp32 results.
Pins, without a GPU, the three claims DESIGN.md makes about the mode: the three-term split is exact, six products reach fp32
accuracy, and the two-term split that was tried first does not (the GPU side: tests/test_gpu_fp32_tc.py)."""
import pytest
import torch

import split_gemm_oracle as S


def _rel(c, ref):
    return float((c.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("scale", [1.0, 1e-3, 37.0])
def test_three_term_split_is_exact(scale):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4096, generator=g) * scale
    x[:8] = torch.tensor([0.0, 1.0, -1.0, 1e-30, 3.0e38, -2.5e-7, 0.1, 1.0 + 2 ** -23])
    x0, x1, x2 = S.split3(x)
    for t in (x0, x1, x2):
        assert torch.equal(t, S.bf16(t))                          # every term is a bf16 number
    assert torch.equal((x0.double() + x1.double() + x2.double()).float(), x.float())
    nz = x != 0
    assert float(((x1.abs() / x.abs())[nz]).max()) <= 2.0 ** -8 and float(((x2.abs() / x.abs())[nz]).max()) <= 2.0 ** -16


@pytest.mark.parametrize("m,n,k", [(128, 96, 32), (64, 256, 256), (96, 64, 1024), (33, 40, 72)])
def test_six_products_reach_fp32_accuracy_and_three_do_not(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g)
    ref = a.double() @ b.double().T
    e_fp32 = _rel(a @ b.T, ref)                                   # what the reference's fp32 Linear gives
    e_split3 = _rel(S.split_matmul(a, b), ref)
    e_split2 = _rel(S.split2_matmul(a, b), ref)
    e_bf16 = _rel(S.bf16(a) @ S.bf16(b).T, ref)
    assert e_split3 < 3e-7 and e_split3 < 4 * e_fp32 + 1e-7       # fp32-class: the 1e-4 parity mode
    assert 10 * e_split3 < e_split2 < 1e-4                        # two terms: 2^-16 products (measured on the GPU: 1e-5 gradients)
    assert e_bf16 > 100 * e_split2                                # plain bf16 operands: the 2e-3 mode


def test_pass_table_matches_the_kernel_encoding():
    """gemm_tc.cu packs the (A image, B image) pairs of the six passes into two hex constants."""
    ia = [(0x001102 >> (4 * p)) & 3 for p in range(6)]
    ib = [(0x010120 >> (4 * p)) & 3 for p in range(6)]
    assert list(zip(ia, ib)) == S.PASSES
    assert sorted(S.PASSES) == sorted((i, j) for i in range(3) for j in range(3) if i + j <= 2)    # every product of weight >= 2^-18
    assert S.PASSES[-1] == (0, 0)                                 # the main accumulator's pass comes last
