"""Delay lines used by the GSC pipelines -- mirrors ``DistantSpeech/beamformer/utils.py``
(DelaySamples :241-274).  Pure buffer shuffling, no arithmetic; inside FDGSC the same
delay lines live in the kernel's shared memory."""
import ctypes as C

import numpy as np

from .. import _lib as L


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


def pcm16_to_float(pcm):
    """int16 samples -> float32 / 32767 on the device, exactly like load_audio (utils.py:184-185:
    ``astype(float32) / float(iinfo(int16).max)`` -- 32767, not 32768).  NumPy or torch in, same kind out."""
    t = L.require_cuda()
    as_torch = isinstance(pcm, t.Tensor)
    d = (pcm if as_torch else t.as_tensor(np.ascontiguousarray(pcm, dtype=np.int16))).to("cuda").contiguous()
    if d.dtype != t.int16:
        raise ValueError("pcm16_to_float expects int16 samples")
    out = t.empty(d.shape, dtype=t.float32, device="cuda")
    L.check(L.lib().ds_pcm16_to_float_run(d.numel(), L.ptr(d), L.ptr(out), L.stream_ptr()), "ds_pcm16_to_float_run")
    return out if as_torch else out.cpu().numpy()


def float_to_pcm16(audio):
    """float samples -> ``(audio * 32767).astype(int16)`` on the device (save_audio, utils.py:193; the cast
    truncates toward zero like NumPy's).  The product is taken in the input's precision, like NumPy does:
    float64 arrays (what Transform.istft and the pipelines return) multiply in double, float32 arrays in float."""
    t = L.require_cuda()
    as_torch = isinstance(audio, t.Tensor)
    d = (audio if as_torch else t.as_tensor(np.ascontiguousarray(audio))).to("cuda").contiguous()
    if d.dtype not in (t.float32, t.float64):
        d = d.to(t.float64)
    out = t.empty(d.shape, dtype=t.int16, device="cuda")
    if d.dtype == t.float64:
        L.check(L.lib().ds_double_to_pcm16_run(d.numel(), L.ptr(d), L.ptr(out), L.stream_ptr()), "ds_double_to_pcm16_run")
    else:
        L.check(L.lib().ds_float_to_pcm16_run(d.numel(), L.ptr(d), L.ptr(out), L.stream_ptr()), "ds_float_to_pcm16_run")
    return out if as_torch else out.cpu().numpy()


def load_audio(filename: str) -> np.ndarray:
    """utils.py:182-187: read a wav file; int16 data is scaled to float32 (on the device)."""
    from scipy.io import wavfile
    _, audio_data = wavfile.read(filename)
    if audio_data.dtype == np.int16:
        audio_data = pcm16_to_float(audio_data)
    return audio_data


def save_audio(filename: str, audio: np.ndarray, fs=16000):
    """utils.py:190-196: write ``audio`` (Nsamples, Nchannels) as 16-bit wav."""
    from scipy.io import wavfile
    if not filename.endswith(".wav"):
        filename = filename + ".wav"
    wavfile.write(filename, fs, float_to_pcm16(audio))
