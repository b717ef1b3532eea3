"""CPU: the arithmetic model behind the engine's parity mode, checked in isolation.

Operands are stored as hi = bf16(x), lo = bf16(x - hi); a product uses A_lo*B_hi + A_hi*B_lo + A_hi*B_hi (the dropped
A_lo*B_lo term is ~2^-18 relative).  These tests pin the error class of that scheme against plain bf16 (fast mode)
and against the fp32 operator the reference uses, on GEMM shapes of the network (K = 64 .. 4608)."""
import pytest
import torch


def _split(x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.double(), lo.double()


@pytest.mark.parametrize("k", [64, 576, 4608])
def test_three_product_split_error_class(k):
    g = torch.Generator().manual_seed(k)
    a = torch.randn(96, k, generator=g)
    b = torch.randn(k, 80, generator=g) / k ** 0.5
    exact = a.double() @ b.double()
    ah, al = _split(a)
    bh, bl = _split(b)
    split3 = al @ bh + ah @ bl + ah @ bh            # what the three MMAs accumulate (fp32 accumulation error aside)
    fast = ah @ bh                                  # nsplit = 1
    fp32 = (a @ b).double()                         # the reference's own operator precision
    scale = exact.abs().max()
    e3, e1, e32 = ((v - exact).abs().max() / scale for v in (split3, fast, fp32))
    # operands carry 16 mantissa bits: relative error of a K-term dot product ~ 2^-17 .. 2^-16, independent of K's
    # size class; plain bf16 is ~2^-9; the 3-product scheme stays within ~an order of magnitude of true fp32
    assert e3 < 2.0 ** -14, e3
    assert e1 > 30 * e3, (e1, e3)
    assert e3 < 64 * max(e32, 2.0 ** -24), (e3, e32)


def test_reconstruction_is_exact_to_16_bits():
    x = torch.randn(100000, generator=torch.Generator().manual_seed(1)) * 7.3
    hi, lo = _split(x)
    rel = ((hi + lo) - x.double()).abs() / x.double().abs().clamp_min(1e-30)
    assert rel.max() < 2.0 ** -16 and rel.median() < 2.0 ** -18
