"""Delay lines used by the GSC pipelines -- mirrors ``DistantSpeech/beamformer/utils.py``
(DelaySamples :241-274).  Pure buffer shuffling, no arithmetic; inside FDGSC the same
delay lines live in the kernel's shared memory."""
import numpy as np


class DelaySamples(object):
    def __init__(self, data_len, delay, channel=1, dtype=np.float64):
        self.data_len = data_len
        self.n_delay = delay
        self.buffer = np.zeros(((data_len + delay), channel), dtype=dtype)

    def delay(self, x):
        """x (n_samples,) or (n_samples, n_chs) -> the same signal delayed by ``delay`` samples."""
        if len(x.shape) == 1:
            x = x[:, np.newaxis]
        data_len = x.shape[0]
        if self.n_delay == 0:
            return x
        self.buffer[-data_len:, :] = x
        output = self.buffer[:data_len, :].copy()
        self.buffer[: self.n_delay, :] = self.buffer[-self.n_delay:, :]
        return output
