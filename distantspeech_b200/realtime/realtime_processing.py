"""Chunked streaming front end -- the processing half of ``DistantSpeech/realtime/realtime_processing.py``
(realtime_processing :9, process :78-84, the capture loop's arithmetic :113-131) without the PyAudio device glue
(SURVEY.md 2 scopes the audio devices out; SURVEY.md 8f.2 asks for the chunk API).

A capture thread hands ``process_pcm`` one interleaved int16 buffer of ``chunk`` frames x ``channels`` channels at a time
(what ``stream.read(CHUNK)`` returns); it is scaled by 1 / 32768 like the reference does here (:118 -- not the 32767 of
``load_audio``), channels 1..4 go through the enhancement object's ``process`` (whose recursive state carries over from
chunk to chunk, so any chunk size that is a multiple of the hop works), the result replaces channel 5 and is returned as
int16 PCM (:129-131).  ``process`` is the reference's method of the same name.
"""
import numpy as np


class realtime_processing(object):
    def __init__(self, EnhancementMehtod=None, angle=0, chunk=1024, channels=6, rate=16000, Recording=False,
                 duplex=False, save_rec_to_file=False):
        self.CHUNK = chunk
        self.CHANNELS = channels
        self.RATE = rate
        self._running = False
        self._frames = []
        self.method = 0
        self.EnhancementMethod = EnhancementMehtod
        self.angle = angle
        self.isRecording = Recording
        self.save_rec_to_file = save_rec_to_file
        self.duplex = duplex

    def process(self, data):
        """data [chunk, 4] float -> enhanced [chunk]   (realtime_processing.py:78-84)."""
        if self.EnhancementMethod is None:
            return data[:, 1]
        output = self.EnhancementMethod.process(data)
        return output["data"]

    def process_pcm(self, data: bytes) -> bytes:
        """One captured buffer: ``chunk`` frames of ``channels`` interleaved little-endian int16 samples -> the bytes the
        reference would play back / record (:113-131): channel 5 of the float frame is overwritten with the enhanced
        signal of channels 1..4; all six channels come back when ``save_rec_to_file`` is set, channel 5 alone otherwise."""
        if self.CHANNELS != 6:
            return data                                              # the reference only touches 6-channel captures (:116)
        samps = np.frombuffer(data, dtype='<i2').astype(np.float32, order='C') / 32768.0
        frame = np.reshape(samps, (self.CHUNK, 6)).copy()
        frame[:, 5] = self.process(frame[:, 1:5])
        if self.save_rec_to_file:
            return (frame * 32768).astype('<i2').tobytes()
        return (frame[:, 5] * 32768).astype('<i2').tobytes()

    def start(self):
        raise RuntimeError("audio-device capture (PyAudio) is outside this package: feed process_pcm from your own capture thread")

    def stop(self):
        self._running = False
