"""``DistantSpeech/adaptivefilter/SubbandLmsMc.py`` (SubbandLmsMc :13, update :144-191): per-bin NLMS over
``channel`` inputs; the power estimate is divided by the channel count (:175)."""
from .SubbandAF import SubbandAF


class SubbandLmsMc(SubbandAF):
    def __init__(self, filter_len=2, num_bands=512, channel=1, mu=0.1, normalization=True, alpha=0.9, m=2,
                 hop_length=None, input_td=False):
        SubbandAF.__init__(self, filter_len=filter_len, num_bands=num_bands, mu=mu, normalization=normalization,
                           alpha=alpha, m=m, hop_length=hop_length, input_td=input_td, channel=channel)

    def update(self, x_n, d_n, alpha=1e-4, p=None):
        """x_n [samples, channel], d_n [samples] float blocks, p [half_band(, 1)] -> (err block, W)."""
        return self._update_td(x_n, d_n, alpha, p)
