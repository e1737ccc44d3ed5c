"""STFT-domain adaptive filters of ``DistantSpeech/adaptivefilter`` that lie on the GSC path."""
