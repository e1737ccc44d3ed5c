"""distantspeech_b200 -- B200-native (sm_100a) drop-in for the multichannel
enhancement hot path of wangwei2009/DistantSpeech.

Module paths mirror the reference package (``DistantSpeech.x.y`` ->
``distantspeech_b200.x.y``).  All hot-path arithmetic runs in hand-written CUDA
kernels inside ``libds_b200.so`` (C ABI in ``include/ds_b200.h``); PyTorch only
provides device memory and streams.  There is no CPU fallback.
"""
__version__ = "0.1.0"
