"""tcgen05 TF32 GEMM (oess_gemm_tf32) against a float64 torch reference.

Tolerance: the tensor core reads fp32 operands as TF32 (10-bit mantissa, 2^-11 relative truncation error per
operand), accumulates in fp32 -> |err| <= 2e-3 * sum_k |a_k||b_k| (stated in include/openess_b200.h)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(M, N, K, bias, seed=0):
    from openess_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    bv = torch.randn(N, device="cuda", generator=g) if bias else None
    out = ops.gemm_tf32(a, b, bv)
    ref = a.double() @ b.double().t()
    if bias:
        ref = ref + bv.double()
    bound = 2e-3 * (a.abs().double() @ b.abs().double().t()) + 1e-6
    err = (out.double() - ref).abs()
    assert bool((err <= bound).all()), f"max err {float(err.max())} vs bound {float(bound.min())} at M={M} N={N} K={K}"
    # not a degenerate pass: the result really carries fp32-accumulated products
    assert float(err.max()) < 0.05 * float(ref.abs().max())


@pytest.mark.parametrize("M,N,K,bias", [
    (128, 128, 32, False),      # one tile, one K block
    (128, 128, 128, False),     # one tile, ring wraps once
    (256, 256, 512, True),      # BN = 256, several M tiles, ring wraps many times
    (1000, 200, 260, True),     # ragged M, N, K (TMA zero fill + masked stores)
    (17600, 256, 2048, True),   # DilationFeatureExtractor decoder conv at 110 x 160 (image_model.py:121-124)
    (300, 64, 36, False),       # BN = 64, K < one block
    (77, 11, 512, True),        # text-embedding conv: K classes = 11 (style_networks.py:165), unaligned N
    (40000, 520, 96, True),     # persistent kernel: 313 x 3 tiles over 148 CTAs (both TMEM accumulators, ring across tiles), ragged N
    (19000, 72, 160, False),    # persistent, BN = 128 with a ragged last N tile (72 = 2 x 32 + 8), 149 tiles: one CTA takes two
])
def test_gemm_tf32(M, N, K, bias):
    _check(M, N, K, bias)


def test_gemm_tf32_exact_on_tf32_representable_inputs():
    """Small integers are exact in TF32 and their products/sums exact in fp32: the result must be bit-exact."""
    from openess_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(-8, 9, (384, 96), device="cuda", generator=g).float()
    b = torch.randint(-8, 9, (160, 96), device="cuda", generator=g).float()
    out = ops.gemm_tf32(a, b)
    assert torch.equal(out, a @ b.t())


def test_gemm_tf32_persistent_exact_with_residual_in_place():
    """Many tiles per CTA, bias + residual aliasing the output (the ViT's x += proj(...)): bit-exact on small integers."""
    from openess_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 128 * 301 + 5, 256, 64
    a = torch.randint(-4, 5, (M, K), device="cuda", generator=g).float()
    b = torch.randint(-4, 5, (N, K), device="cuda", generator=g).float()
    bias = torch.randint(-4, 5, (N,), device="cuda", generator=g).float()
    c = torch.randint(-4, 5, (M, N), device="cuda", generator=g).float()
    ref = a @ b.t() + bias + c
    out = ops.gemm_tf32_ex(a, b, bias, residual=c, out=c)
    assert out.data_ptr() == c.data_ptr() and torch.equal(out, ref)


def test_gemm_tf32_cta_pair_exact_with_an_empty_peer_tile():
    """CTA-pair kernel (256 x 256 tiles, cta_group::2): the last pair's second CTA has no rows at all (M = 256 * 80 + 100), N has
    a ragged last tile, K is not a multiple of the 32-wide block; bit-exact on small integers, with GELU off and a bias."""
    from openess_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 256 * 80 + 100, 520, 72
    a = torch.randint(-4, 5, (M, K), device="cuda", generator=g).float()
    b = torch.randint(-4, 5, (N, K), device="cuda", generator=g).float()
    bias = torch.randint(-4, 5, (N,), device="cuda", generator=g).float()
    assert torch.equal(ops.gemm_tf32_ex(a, b, bias), a @ b.t() + bias)


def test_gemm_tf32_argument_errors():
    from openess_b200 import ops
    from openess_b200._lib import OpenESSB200Error
    a = torch.randn(8, 6, device="cuda")
    b = torch.randn(8, 6, device="cuda")
    with pytest.raises(OpenESSB200Error):
        ops.gemm_tf32(a, b)      # K % 4 != 0
