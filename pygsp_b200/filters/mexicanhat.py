"""Mexican hat wavelet bank (mirror of pygsp/filters/mexicanhat.py:55-84)."""
import numpy as np

from .. import utils
from .filter import Filter


class MexicanHat(Filter):
    r"""One scaling function plus ``Nf - 1`` band-pass wavelets ``x exp(-x)``.

    The band-passes are ``g_i(x) = t_i x exp(-t_i x)`` at log-spaced scales
    ``t_i`` between ``2/lmin`` and ``1/lmax`` (``lmin = lmax / lpfactor``); the
    low-pass is ``1.2 e^{-1} exp(-(x / (0.4 lmin))^4)``.  ``lmin`` and the
    scales are frozen from ``G.lmax`` at construction, as in the reference.
    """

    def __init__(self, G, Nf=6, lpfactor=20, scales=None, normalize=False):
        self.lpfactor = lpfactor
        self.normalize = normalize
        lmin = G.lmax / lpfactor
        if scales is None:
            scales = utils.compute_log_scales(lmin, G.lmax, Nf - 1)
        self.scales = scales
        if len(scales) != Nf - 1:
            raise ValueError("len(scales) should be Nf-1.")

        def scaling(x):
            return 1.2 * np.exp(-1) * np.exp(-((x / 0.4 / lmin) ** 4))

        def wavelet(x, t):
            gain = np.sqrt(t) if normalize else 1
            return gain * (t * x) * np.exp(-t * x)

        kernels = [scaling] + [lambda x, t=t: wavelet(x, t) for t in scales]
        super().__init__(G, kernels)

    def _get_extra_repr(self):
        return dict(lpfactor="{:.2f}".format(self.lpfactor), normalize=self.normalize)
