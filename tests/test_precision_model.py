"""NumPy model of the two fp32-grade tensor-core schemes of the GEMM kernel, checked on the CPU:

  TF32X3       a·b ~ hi_a·hi_b + lo_a·hi_b + hi_a·lo_b                     with hi = trunc_tf32(x), lo = trunc_tf32(x - hi)
  TF32_BF16X2  a·b ~ hi_a·hi_b + bf16(lo_a)·bf16(b) + bf16(a)·bf16(lo_b)   (corrections on the bf16 pipe, half the cost each)

The products are summed in fp64 here, so the numbers isolate the error of the SPLIT itself (the device adds its accumulator's
~2.7e-8 per MMA on top, bounded by the chunked promotion).  They document why both schemes meet the 1e-5 parity bar while a single TF32
pass (error ~7e-4) cannot, and why bf16 — same exponent range as fp32 — is enough for terms that are 2^-11 of the product."""
import numpy as np
import pytest


def trunc_tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_rn(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000          # round to nearest even on the upper 16 bits
    return u.astype(np.uint32).view(np.float32)


def rel(got, ref):
    return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


@pytest.fixture(scope="module")
def operands():
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, (96, 1024)).astype(np.float32)
    B = rng.normal(0, 0.5, (64, 1024)).astype(np.float32)
    return A, B, A.astype(np.float64) @ B.astype(np.float64).T


def test_single_tf32_pass_misses_the_parity_bar(operands):
    A, B, ref = operands
    got = trunc_tf32(A).astype(np.float64) @ trunc_tf32(B).astype(np.float64).T
    assert 1e-4 < rel(got, ref) < 2e-3


def test_tf32x3_split_error(operands):
    A, B, ref = operands
    ah, bh = trunc_tf32(A), trunc_tf32(B)
    al, bl = trunc_tf32(A - ah), trunc_tf32(B - bh)           # the hardware truncates the lo operand as well
    f = lambda x: x.astype(np.float64)
    got = f(ah) @ f(bh).T + f(al) @ f(bh).T + f(ah) @ f(bl).T
    assert rel(got, ref) < 1.5e-6


def test_tf32_plus_two_bf16_corrections_split_error(operands):
    A, B, ref = operands
    ah, bh = trunc_tf32(A), trunc_tf32(B)
    f = lambda x: x.astype(np.float64)
    got = f(ah) @ f(bh).T + f(bf16_rn(A - ah)) @ f(bf16_rn(B)).T + f(bf16_rn(A)) @ f(bf16_rn(B - bh)).T
    e = rel(got, ref)
    assert e < 2e-6, e
    # and it really is the corrections that buy the accuracy: without them the error is the single-pass one
    assert rel(f(ah) @ f(bh).T, ref) > 100 * e


def test_bf16_keeps_the_fp32_exponent_range():
    """lo terms of tiny / huge operands survive bf16 (they would flush or overflow in fp16)."""
    x = np.array([3e-30, 1.5e30, -7e-25], dtype=np.float32)
    lo = x - trunc_tf32(x)
    r = bf16_rn(lo)
    assert np.all(np.isfinite(r)) and np.all((r != 0) == (lo != 0))
    assert np.allclose(r, lo, rtol=2 ** -8)
