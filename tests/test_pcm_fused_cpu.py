"""CPU: the arithmetic the fused int16 ingest / egress kernels rely on (csrc/transform.cu pcm16_scale /
pcm16_from_sample), checked exhaustively against NumPy's evaluation of the reference expressions
(beamformer/utils.py:184-185 load_audio, :193 save_audio)."""
import numpy as np


def _fma32(a, b, c):
    """float32 fused multiply-add, exact: float64 holds a*b exactly (48 bits) and every sum below is exact in float64
    (|rem| is tiny next to q, asserted), so one final rounding to float32 is the fused result."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def test_division_by_32767_without_a_division_is_exact_for_every_int16():
    x = np.arange(-32768, 32768, dtype=np.int64).astype(np.float32)
    ref = x / np.float32(np.iinfo(np.int16).max)                       # load_audio (float32 / 32767)
    r = np.float32(1.0) / np.float32(32767.0)
    q = (x * r).astype(np.float32)
    rem64 = np.float64(x) - np.float64(q) * 32767.0                     # exact: both terms are multiples of 2^-23 below 2^16
    rem = rem64.astype(np.float32)
    assert np.array_equal(rem.astype(np.float64), rem64)                # the fused remainder is exactly representable
    from fractions import Fraction
    out = np.empty_like(q)
    for i in range(x.shape[0]):                                         # final fma rounded once, in exact arithmetic
        v = Fraction(float(rem[i])) * Fraction(float(r)) + Fraction(float(q[i]))
        out[i] = _round_f32(v)
    assert np.array_equal(out, ref)


def _round_f32(fr):
    """round-to-nearest-even of an exact Fraction to float32"""
    if fr == 0:
        return np.float32(0)
    sign = -1 if fr < 0 else 1
    a = abs(fr)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if (a.numerator << max(0, -e)) < (a.denominator << max(0, e)):
        e -= 1
    scaled = a / (2 ** (e - 23)) if e - 23 >= 0 else a * (2 ** (23 - e))      # in [2^23, 2^24)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > 0.5 or (rem == 0.5 and n % 2 == 1):
        n += 1
    return np.float32(sign * float(n) * 2.0 ** (e - 23))


def test_pcm16_egress_expression():
    """save_audio on what Transform.istft returns (float64 holding float32 values): product in double, truncation."""
    rng = np.random.default_rng(3)
    y32 = (rng.standard_normal(100000) * 0.3).astype(np.float32)
    ref = (y32.astype(np.float64) * np.iinfo(np.int16).max).astype(np.int16)        # utils.py:193
    dev = np.clip(np.trunc(y32.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)   # pcm16_from_sample
    inside = np.abs(y32.astype(np.float64) * 32767.0) < 32768
    assert np.array_equal(ref[inside], dev[inside])
    # round trip of load_audio(save_audio(x)) is the identity on int16 data
    pcm = rng.integers(-32768, 32768, size=65536, dtype=np.int16)
    x = pcm.astype(np.float32) / np.float32(32767.0)
    back = (x.astype(np.float64) * 32767).astype(np.int16)
    assert 0.4 < np.mean(back == pcm) < 0.6                              # truncation loses 1 LSB on about half the samples: the reference's behaviour
    assert np.max(np.abs(back.astype(np.int32) - pcm)) <= 1
