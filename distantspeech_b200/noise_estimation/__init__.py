"""``DistantSpeech/noise_estimation/__init__.py`` exports (MCRA2 and McMcra are outside the hot path)."""
from .mcra import NoiseEstimationMCRA  # noqa: F401
from .mcspp_base import McSppBase  # noqa: F401
from .mcspp import McSpp  # noqa: F401
