"""``DistantSpeech/noise_estimation/mccdr.py`` module path (McCDR :25)."""
from .mcspp import McCDR  # noqa: F401
