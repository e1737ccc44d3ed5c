"""``DistantSpeech/adaptivefilter/SubbandLMS.py`` (SubbandLMS :12, update :28-84): per-bin NLMS, one input channel."""
from .SubbandAF import SubbandAF


class SubbandLMS(SubbandAF):
    def __init__(self, filter_len=2, num_bands=512, mu=0.1, normalization=True, alpha=0.9, m=2, hop_length=None,
                 input_td=False):
        SubbandAF.__init__(self, filter_len=filter_len, num_bands=num_bands, mu=mu, normalization=normalization,
                           alpha=alpha, m=m, hop_length=hop_length, input_td=input_td, channel=1)

    def update(self, x_n, d_n, alpha=1e-4, p=None):
        """x_n, d_n [samples] float blocks (a multiple of hop_length), p float or [half_band(, 1)] -> (err block, W)."""
        assert x_n.shape == d_n.shape, 'x_n and d_n must be same shape of [samples, ]'
        assert len(x_n.shape) == 1, 'x_n must be shape of [samples, ]'
        return self._update_td(x_n, d_n, alpha, p)
