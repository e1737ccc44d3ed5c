"""Shared constants of the single-channel noise trackers (the public attributes of
``DistantSpeech/noise_estimation/NoiseEstimationBase.py:5-60``).  The recursions themselves run on the device;
this class only carries the tunables the kernels are parameterised with, plus the one host helper
(``smooth_psd``) some notebooks call directly."""
import numpy as np

# attribute -> default (Base :12-31).  alpha_d: noise smoothing, alpha_s: power smoothing, delta_s: presence
# threshold on S / Smin, alpha_p: presence smoothing, b: 3-tap frequency window, L: minimum-search window,
# init_frame: frames of warm-up.
_DEFAULTS = dict(alpha_d=0.95, alpha_s=0.8, delta_s=5, alpha_p=0.2, ell=1, L=125, init_frame=15, frm_cnt=0)


class NoiseEstimationBase(object):
    def __init__(self, nfft=256) -> None:
        self.nfft = nfft
        self.half_bin = int(nfft / 2 + 1)
        self.b = [0.25, 0.5, 0.25]
        for name, value in _DEFAULTS.items():
            setattr(self, name, value)

    def smooth_psd(self, x, previous_x, win, alpha):
        """Frequency smoothing with ``win`` ("same" part of the full convolution) followed by first-order recursive
        smoothing in time (Base :33-51)."""
        half = (len(win) - 1) // 2
        in_freq = np.convolve(x, win)[half: len(x) + half]
        return alpha * previous_x + (1 - alpha) * in_freq

    def estimation(self, X):
        pass
