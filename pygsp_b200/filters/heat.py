"""Heat kernel filter bank (mirror of pygsp/filters/heat.py:102-119)."""
import numpy as np

from .filter import Filter


class Heat(Filter):
    r"""Low-pass heat kernels ``g(x) = min(exp(-scale * x / lmax), 1)``.

    ``G.lmax`` is read when the kernel is *evaluated*, as in the reference, so
    estimate lmax before filtering.  ``normalize=True`` divides by the kernel's
    norm over the exact spectrum ``G.e``, which needs the dense Fourier basis:
    not available in this engine.
    """

    def __init__(self, G, scale=10, normalize=False):
        try:
            iter(scale)
        except TypeError:
            scale = [scale]
        if normalize:
            raise NotImplementedError("normalize=True needs the exact spectrum G.e (dense "
                                      "eigendecomposition), outside the Chebyshev path.")
        self.scale = scale
        self.normalize = normalize

        def heat(x, tau):
            return np.minimum(np.exp(-tau * x / G.lmax), 1)

        super().__init__(G, [lambda x, tau=tau: heat(x, tau) for tau in scale])

    def _get_extra_repr(self):
        scale = "[" + ", ".join("{:.2f}".format(s) for s in self.scale) + "]"
        return dict(scale=scale, normalize=self.normalize)
