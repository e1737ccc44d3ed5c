"""Constants and helpers shared by the noise estimators -- mirrors
``DistantSpeech/noise_estimation/NoiseEstimationBase.py:5-60``."""
import numpy as np


class NoiseEstimationBase(object):
    def __init__(self, nfft=256) -> None:
        self.nfft = nfft
        self.half_bin = int(self.nfft / 2 + 1)
        self.alpha_d = 0.95
        self.alpha_s = 0.8
        self.delta_s = 5
        self.alpha_p = 0.2
        self.ell = 1
        self.b = [0.25, 0.5, 0.25]
        self.L = 125
        self.init_frame = 15
        self.frm_cnt = 0

    def smooth_psd(self, x, previous_x, win, alpha):
        """3-tap frequency smoothing + recursive time smoothing (Base :33-51); host helper."""
        w = len(win)
        smoothed_f = np.convolve(x, win)
        smoothed_f_val = smoothed_f[int((w - 1) / 2): int(-((w - 1) / 2))]
        return alpha * previous_x + (1 - alpha) * smoothed_f_val

    def estimation(self, X):
        pass
