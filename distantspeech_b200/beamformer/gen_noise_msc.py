"""Diffuse-field coherence matrix -- host-side precompute mirroring
``DistantSpeech/beamformer/gen_noise_msc.py:7-28``."""
import numpy as np

from .MicArray import MicArray


def gen_noise_msc(mic: MicArray, nfft=256, Fvv_max=0.9998):
    """Gamma[k, i, j] = sin(x)/x, x = 2 pi f_k d_ij / c (f_0 = 1e-6), diagonal = Fvv_max."""
    M = mic.M
    half_bin = round(nfft / 2 + 1)
    f = np.linspace(0, mic.fs / 2, half_bin)
    f[0] = 1e-6
    diff = mic.mic_loc[:, None, :] - mic.mic_loc[None, :, :]
    dist = np.sqrt(np.sum(diff ** 2, axis=-1))
    Fvv = np.zeros((half_bin, M, M))
    for i in range(M):
        for j in range(M):
            if i == j:
                Fvv[:, i, j] = Fvv_max
            else:
                arg = 2 * np.pi * f * dist[i, j] / mic.c
                Fvv[:, i, j] = np.sin(arg) / arg
    return Fvv
