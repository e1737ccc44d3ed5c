"""Fractional-delay filter bank -- host-side precompute mirroring
``DistantSpeech/transform/multirate.py:4-51`` (Hann-windowed sinc, L = 81 taps
plus the largest integer delay)."""
import numpy as np


def fractional_delay_filter_bank(delays):
    """delays [chs] in (fractional) samples -> filter bank [filter_len, chs]."""
    delays = np.array(delays, dtype=np.float64)
    delays -= delays.min()
    n_ch = delays.shape[0]
    L = 81
    filter_length = L + int(np.ceil(delays).max())
    di = np.floor(delays).astype(np.int64)
    df = delays - di
    T = np.arange(L)
    window = np.hanning(L)
    bank = np.zeros((n_ch, filter_length))
    for c in range(n_ch):
        bank[c, di[c]: di[c] + L] = window * np.sinc(T - df[c] - (L - 1) / 2)
    return bank.T
