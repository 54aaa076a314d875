"""Executable model of the FP16x3 arithmetic, checked against float64 on the CPU.

The default engine of wide layers feeds kind::f16 tensor-core MMAs with operands staged as scaled (hi, lo) FP16
pairs.  This file restates that arithmetic in NumPy exactly as csrc/umma_conv_f16.cu defines it —

    scale  = 2^(15 - e)  with max|v| = f * 2^e, f in [0.5, 1)            (f16_scale)
    hi, lo = fp16(v * scale), fp16(v * scale - hi)                        (f16_split, round to nearest)
    D      = A_hi.[B_hi | B_lo] ; D[:, :n] += A_lo.B_hi ; out = (D[:, :n] + D[:, n:]) / (scale_a * scale_b)
    wgrad: groups of 8 chunks x 32 pixels accumulate on the tensor core, the groups are added in FP32 registers

(FP16 x FP16 products are exact in the FP32 accumulator) — and checks that it meets the FP32 parity bar of the
GPU tests (rtol 1e-4 of the output scale, tests/helpers.py) with two orders of magnitude to spare, on the shapes of
a BASELINE config-2 layer, on a reduction as long as a wgrad's, and on data with a wide dynamic range; a single
FP16 product (no lo terms) must fail the same bar, which proves the check has teeth.  The `-m gpu` counterpart on
the real kernels is tests/test_tensor_core_conv.py::test_f16x3_conv_dynamic_range and
tests/test_full_size_elementwise.py.
"""
import numpy as np

import helpers as H


def f16_scale(x):
    mx = np.float32(np.abs(x).max())
    if not mx > 0:
        return np.float32(1.0)
    _, e = np.frexp(mx)
    return np.float32(np.ldexp(np.float32(1.0), int(np.clip(15 - e, -126, 126))))


def f16_split(x):
    s = f16_scale(x)
    sv = (x.astype(np.float32) * s).astype(np.float32)
    hi = sv.astype(np.float16)
    lo = (sv - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32), s


def f16x3_matmul(a, b, group=None):
    """a [M, K] @ b [K, N] the FP16x3 way; ``group``: K elements accumulated on the tensor core before the partial
    sum is added in FP32 registers (wgrad: 8 chunks x 32 pixels)."""
    a_hi, a_lo, sa = f16_split(a)
    b_hi, b_lo, sb = f16_split(b)
    k = a.shape[1]
    group = group or k
    acc = np.zeros((a.shape[0], b.shape[1]), dtype=np.float32)
    for k0 in range(0, k, group):
        sl = slice(k0, min(k0 + group, k))
        main = a_hi[:, sl] @ b_hi[sl] + a_lo[:, sl] @ b_hi[sl]      # D[:, :n]
        cross = a_hi[:, sl] @ b_lo[sl]                                      # D[:, n:]
        acc += (main + cross).astype(np.float32)
    return acc * np.float32(1.0 / (float(sa) * float(sb)))


def fp16_matmul(a, b):
    """One FP16 product per MAC (hi parts only): what the scheme would be without error compensation."""
    a_hi, _, sa = f16_split(a)
    b_hi, _, sb = f16_split(b)
    return (a_hi @ b_hi) * np.float32(1.0 / (float(sa) * float(sb)))


def scaled_err(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / np.abs(want).max())


def test_split_is_exact_to_22_bits_and_scale_is_a_power_of_two():
    rng = np.random.RandomState(0)
    x = rng.randn(4096).astype(np.float32) * np.float32(3.7e-3)
    hi, lo, s = f16_split(x)
    assert np.log2(float(s)) == np.round(np.log2(float(s)))
    assert 2.0 ** 14 <= np.abs(x * s).max() < 2.0 ** 15
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) / float(s)
    # 11 + 11 significant bits: elements within 2^-7 of the maximum keep a relative error below 2^-21
    big = np.abs(x) >= np.abs(x).max() * 2.0 ** -7
    assert (np.abs(rec[big] - x[big]) <= np.abs(x[big]) * 2.0 ** -21).all()
    # the absolute error never exceeds half an ulp of lo at the top of the range: 2^-9 against a scaled maximum
    # >= 2^14, i.e. 2^-23 of the tensor's largest magnitude
    assert np.abs(rec - x).max() <= np.abs(x).max() * 2.0 ** -23


def test_conv_layer_shaped_contraction_meets_the_fp32_bar():
    # fprop of a 64 -> 64, 3x3 layer as its implicit GEMM: 4096 output pixels x (64 * 9) x 64 filters
    rng = np.random.RandomState(1)
    a = rng.randn(4096, 576).astype(np.float32)
    b = (rng.randn(576, 64) / np.sqrt(576)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    err3, err1 = scaled_err(f16x3_matmul(a, b), want), scaled_err(fp16_matmul(a, b), want)
    assert err3 <= 1e-6, err3                 # measured on the B200: <= 4e-6 including the FP32 epilogue
    assert err3 <= H.RTOL_FP32 / 100
    assert err1 > H.RTOL_FP32, err1           # teeth: plain FP16 inputs do not pass the FP32 bar


def test_wgrad_length_reduction_with_group_promotion():
    # wgrad reduces over the pixels of the batch: 65536 here (64 images of 32 x 32), promoted every 256 pixels
    rng = np.random.RandomState(2)
    x = rng.randn(64, 65536).astype(np.float32)            # [c][pixel]
    dy = (rng.randn(65536, 32) * 1e-3).astype(np.float32)  # [pixel][k]
    want = x.astype(np.float64) @ dy.astype(np.float64)
    err = scaled_err(f16x3_matmul(x, dy, group=256), want)
    assert err <= 2e-6, err


def test_wide_dynamic_range_keeps_absolute_accuracy():
    # log-normal magnitudes spread over e^(+-6): small elements lose relative, never absolute, accuracy
    rng = np.random.RandomState(3)
    a = (rng.randn(1024, 576) * np.exp(rng.uniform(-6, 6, (1024, 576)))).astype(np.float32)
    b = (rng.randn(576, 64) * np.exp(rng.uniform(-6, 6, (576, 64)))).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    err = scaled_err(f16x3_matmul(a, b), want)
    assert err <= H.RTOL_FP32 / 20, err
    # an all-zero operand and a single huge element are handled by the scale
    z = np.zeros((8, 64), np.float32)
    assert np.array_equal(f16x3_matmul(z, b[:64]), np.zeros((8, 64), np.float32))
    spike = a.copy()
    spike[0, 0] = np.float32(1e30)
    assert scaled_err(f16x3_matmul(spike, b), spike.astype(np.float64) @ b.astype(np.float64)) <= H.RTOL_FP32 / 20


# ---- the 3xTF32 engine: shapes FP16x3 does not take, and GEMM ----------------------------------
def tf32_trunc(x):
    """What the tensor core reads from an FP32 word under kind::tf32: the 13 low mantissa bits are ignored."""
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32x3_matmul(a, b):
    """hi = the word as the hardware truncates it, lo = v - trunc(v) (itself read truncated);
    D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi with FP32 accumulation (csrc/umma_common.cuh)."""
    a_hi, b_hi = tf32_trunc(a), tf32_trunc(b)
    a_lo, b_lo = tf32_trunc(a - a_hi), tf32_trunc(b - b_hi)
    return (a_lo @ b_hi + a_hi @ b_lo) + a_hi @ b_hi


def test_tf32x3_model_meets_the_fp32_bar_and_plain_tf32_does_not():
    rng = np.random.RandomState(4)
    a = rng.randn(4096, 288).astype(np.float32)             # a 32 -> 64, 3x3 layer: K = 32 * 9
    b = (rng.randn(288, 64) / np.sqrt(288)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    err3 = scaled_err(tf32x3_matmul(a, b), want)
    err1 = scaled_err(tf32_trunc(a) @ tf32_trunc(b), want)
    assert err3 <= 2e-6, err3                                # measured on the B200: <= 1e-5
    assert err1 > H.RTOL_FP32, err1                          # truncated TF32 inputs alone: the rtol 1e-2 mode
    assert err1 <= H.RTOL_TF32
