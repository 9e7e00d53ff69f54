"""``Filter``: a bank of spectral kernels applied by Chebyshev recurrence on the GPU.

Mirror of the hot-path part of ``pygsp/filters/filter.py``: the container
(:56-69), ``evaluate`` (:112-144), ``filter`` with its shape conventions
(:146-328), ``analyze`` / ``synthesize`` (:330-348), ``localize`` (:350-391)
and the small operators (:83-103).  ``method='exact'`` needs the dense
eigendecomposition and is not part of this engine.
"""
import numpy as np

from .. import _native as nat
from .. import utils
from ..graphs.graph import Graph
from . import approximations

_logger = utils.build_logger(__name__)


class Filter:
    r"""Filter bank defined by kernel functions of the graph frequencies.

    Parameters
    ----------
    G : Graph
    kernels : function or list of functions (one per filter, NumPy array in -> out)
    """

    def __init__(self, G, kernels):
        self.G = G
        single = callable(kernels) and not hasattr(kernels, "__iter__")
        self._kernels = [kernels] if single else kernels
        n_out = len(self._kernels)
        # one input feature, one output feature per kernel (an "analysis" bank)
        self.n_features_in = 1
        self.n_features_out = n_out
        self.shape = (n_out, 1)
        self.Nf = self.n_filters = n_out
        # synthesis as ONE backward recurrence (K SpMMs) instead of the reference's Nf
        # forward recurrences (Nf * K SpMMs); same value, different rounding.  Set to
        # False to reproduce the reference's operation order.
        self.fused_synthesis = True
        # a single filter is evaluated by Clenshaw's backward recurrence: 4 instead of 5
        # passes over the signal block per order (no accumulator block), same value,
        # different rounding.  False = the reference's forward recurrence and operation order.
        self.clenshaw = True

    def _get_extra_repr(self):
        return dict()

    def __repr__(self):
        fields = [("in", self.n_features_in), ("out", self.n_features_out)]
        fields += list(self._get_extra_repr().items())
        return "{}({})".format(type(self).__name__, ", ".join("%s=%s" % kv for kv in fields))

    def __len__(self):
        return self.n_filters

    def __getitem__(self, key):
        return Filter(self.G, self._kernels[key])

    def __add__(self, other):
        if not isinstance(other, Filter):
            return NotImplemented
        return Filter(self.G, self._kernels + other._kernels)

    def __call__(self, x):
        if isinstance(x, Graph):
            return Filter(x, self._kernels)
        return self.evaluate(x)

    def __matmul__(self, other):
        return self.filter(other)

    def evaluate(self, x):
        r"""Frequency response of every kernel at ``x``: shape (Nf, *x.shape)."""
        freqs = np.asanyarray(x)
        response = np.empty((self.Nf,) + freqs.shape)
        for row, g in zip(response, self._kernels):
            row[...] = g(freqs)
        return response

    def filter(self, s, method="chebyshev", order=30):
        r"""Filter signals (analysis or synthesis) -- filter.py:146-328.

        ``s`` is read as (N, N_SIGNALS, N_FEATURES).  A last dimension that is
        neither 1 nor Nf is a signal dimension.  One input feature -> analysis:
        every filter is applied, output (N, N_SIGNALS, Nf).  Nf input features
        -> synthesis: filter i is applied to feature i and the results are
        summed, output (N, N_SIGNALS).  Singleton dimensions are squeezed.
        NumPy in -> NumPy out; CUDA tensors stay on the device.
        """
        torch = nat.require_cuda()
        s = self.G._check_signal(s)
        if method != "chebyshev":
            if method == "exact":
                raise NotImplementedError(
                    "method='exact' needs the dense Fourier basis, which is outside this "
                    "engine's path; use method='chebyshev'.")
            raise ValueError("Unknown method {}.".format(method))

        if s.ndim == 1 or s.shape[-1] not in [1, self.Nf]:
            if s.ndim == 3:
                raise ValueError("Third dimension (#features) should be either 1 or the number "
                                 "of filters Nf = {}, got {}.".format(self.Nf, tuple(s.shape)))
            s = s[..., None]
        n_features_in = s.shape[-1]
        if s.ndim < 3:
            s = s[:, None, :]
        if s.ndim > 3:
            raise ValueError("At most 3 dimensions: #nodes x #signals x #features.")
        n_signals = s.shape[1]
        N = self.G.N

        c = approximations.compute_cheby_coeff(self, m=order)
        c = np.atleast_2d(np.asarray(c, dtype=np.float64))
        if c.shape[1] < 2:
            raise TypeError("The coefficients have an invalid shape")
        L = approximations._laplacian_on_device(self.G)
        view = approximations._GraphView(L)

        if n_features_in == 1:                                   # analysis
            flat = s.reshape(N, n_signals)
            if approximations._is_pinned_block(flat, L):
                # page-locked host block: column chunks, transfers overlapped with the recurrence
                from . import pipeline
                r = pipeline.filter_pinned(L, self.G.lmax, c, flat, clenshaw=self.clenshaw)
                return r.permute(1, 2, 0).squeeze()
            x, _, kind = approximations._as_device_block(view, flat)
            if self.clenshaw and c.shape[0] == 1:
                r = approximations.cheby_clenshaw_device(L, self.G.lmax, c, x)[None]
            else:
                r = approximations.cheby_op_device(L, self.G.lmax, c, x)  # (Nf, N, nsig)
            out = r.permute(1, 2, 0)                                      # (N, nsig, Nf)
        else:                                                    # synthesis
            x, _, kind = approximations._as_device_block(view, s.reshape(N, -1))
            x = x.reshape(N, n_signals, n_features_in)
            if self.fused_synthesis and n_features_in <= 16:
                out = approximations.cheby_clenshaw_device(L, self.G.lmax, c, x.permute(2, 0, 1))
            else:
                out = torch.zeros((N, n_signals), dtype=L.dtype, device=L.device)
                for i in range(n_features_in):
                    xi = x[:, :, i].contiguous()
                    out += approximations.cheby_op_device(L, self.G.lmax, c[i], xi)[0]
            out = out[:, :, None]
        out = out.squeeze()
        return approximations._leave_device(out, kind)

    def analyze(self, s, method="chebyshev", order=30):
        r"""Alias of :meth:`filter` for single-feature input (filter.py:330-336)."""
        if s.ndim == 3 and s.shape[-1] != 1:
            raise ValueError("Last dimension (#features) should be 1, got {}.".format(
                tuple(s.shape)))
        return self.filter(s, method, order)

    def synthesize(self, s, method="chebyshev", order=30):
        r"""Alias of :meth:`filter` for Nf-feature input (filter.py:338-348)."""
        if s.shape[-1] != self.Nf:
            raise ValueError("Last dimension (#features) should be the number of filters "
                             "Nf = {}, got {}.".format(self.Nf, tuple(s.shape)))
        return self.filter(s, method, order)

    def compute_frame(self, **kwargs):
        r"""Matrix of the analysis operator, (N Nf, N): one delta per vertex through
        :meth:`filter` (filter.py:540-603; ``method='chebyshev'`` only, like :meth:`filter`)."""
        if self.G.N > 2000:
            _logger.warning("Creating a big matrix. You should prefer the filter method.")
        s = np.identity(self.G.N)
        return self.filter(s, **kwargs).T.reshape(-1, self.G.N)

    def localize(self, i, **kwargs):
        r"""Kernels localised at vertex ``i``: sqrt(N) g(L) delta_i (filter.py:350-391)."""
        delta = np.zeros(self.G.N)
        delta[i] = 1
        return self.filter(delta, **kwargs) * np.sqrt(self.G.N)
