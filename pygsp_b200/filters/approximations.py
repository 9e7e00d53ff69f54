"""Chebyshev approximation of graph filters, device side.

Mirror of ``pygsp/filters/approximations.py``: ``compute_cheby_coeff`` (:9-55,
host quadrature, K+1 scalars per filter), ``cheby_op`` (:58-114, THE hot loop)
and ``cheby_rect`` (:117-163).  The recurrence runs in ``libgspb200``:
one fused CUDA kernel per order (csrc/cheby.cu) instead of SciPy's
``csr_matvecs`` + NumPy temporaries + fancy-indexed accumulation.
"""
import numpy as np

from .. import _native as nat
from .. import utils

_logger = utils.build_logger(__name__)
PATCH_DEFAULT_DTYPE = None     # engine dtype for reference (scipy) graphs, see patch_pygsp()


@utils.filterbank_handler
def compute_cheby_coeff(f, m=30, N=None, *args, **kwargs):
    r"""Chebyshev coefficients of filter ``i`` of the bank ``f`` on [0, lmax].

    Chebyshev-Gauss quadrature with ``N`` (default ``m + 1``) nodes:
    ``c[o] = 2/N sum_j g(a cos(t_j) + a) cos(o t_j)``, ``t_j = pi (j + 1/2) / N``,
    ``a = lmax / 2``.  Evaluated on the host in float64: it is K+1 numbers and
    must see exactly the ``G.lmax`` the device recurrence is given.
    """
    G = f.G
    i = kwargs.pop("i", 0)
    if not N:
        N = m + 1
    half = G.lmax / 2.0
    theta = np.pi * (np.arange(N) + 0.5) / N
    samples = f._kernels[i](half * np.cos(theta) + half)
    orders = np.arange(m + 1)[:, None]
    return (2.0 / N) * (np.cos(orders * theta[None, :]) @ samples)


def _as_device_block(G, signal):
    """signal -> contiguous (N, nsig) device tensor in the graph's dtype."""
    torch = nat.require_cuda()
    kind = "cuda"
    if not torch.is_tensor(signal):
        signal = torch.from_numpy(np.ascontiguousarray(np.asarray(signal)))
        kind = "numpy"
    elif not signal.is_cuda:
        kind = "pinned" if signal.is_pinned() else "cpu"
    one_d = signal.dim() == 1
    x = signal.to(device=G.device, dtype=G.dtype, non_blocking=True)
    x = x.reshape(x.shape[0], -1).contiguous()
    return x, one_d, kind


def _is_pinned_block(s, L):
    """A contiguous page-locked host tensor of the engine's dtype, big enough to pipeline."""
    torch = nat.require_cuda()
    return (torch.is_tensor(s) and not s.is_cuda and s.dim() == 2 and s.dtype == L.dtype
            and s.is_contiguous() and s.is_pinned() and s.numel() > 0)


def _leave_device(t, kind):
    torch = nat.require_cuda()
    if kind == "cuda":
        return t
    if kind == "numpy":
        return t.cpu().numpy()
    out = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=(kind == "pinned"))
    out.copy_(t, non_blocking=False)
    return out


def _laplacian_on_device(G):
    """(DeviceCSR L, lmax) of any graph object offering .L / .lmax / .N.

    A graph of this package already holds L in HBM.  A *reference* pygsp graph
    (scipy ``G.L``) is uploaded once and the copy is cached on the object, so
    that this function can stand in for ``pygsp.filters.approximations.cheby_op``.
    """
    from ..graphs.csr import DeviceCSR
    L = G.L
    if isinstance(L, DeviceCSR):
        return L
    torch = nat.require_cuda()
    cached = getattr(G, "_gspb200_L", None)
    if cached is None or cached[0] is not L:
        dtype = getattr(G, "_gspb200_dtype", None) or PATCH_DEFAULT_DTYPE or torch.float32
        dev = torch.device("cuda:%d" % torch.cuda.current_device())
        cached = (L, DeviceCSR.from_scipy(L, dtype, dev))
        G._gspb200_L = cached
    return cached[1]


def cheby_op_device(L, lmax, c, x, out=None, work=None):
    """Device-to-device core: x (N, nsig) tensor -> r (Nscales, N, nsig) tensor.

    ``out`` / ``work`` ((Nscales, N, nsig) and (2, N, nsig), contiguous) may be given by callers
    that keep their buffers (the column-chunk pipeline of host signals)."""
    torch = nat.require_cuda()
    c = np.atleast_2d(np.asarray(c, dtype=np.float64))
    nscales, M = c.shape
    if M < 2:
        raise TypeError("The coefficients have an invalid shape")
    n, nsig = x.shape
    c = np.ascontiguousarray(c)
    r = out if out is not None else torch.empty((nscales, n, nsig), dtype=L.dtype, device=L.device)
    if work is None:
        work = torch.empty((2, n, nsig), dtype=L.dtype, device=L.device)
    plan = L.tile_plan(nsig, nscales)
    with torch.cuda.device(L.device):
        nat.call("gsp_cheby_op_" + nat.suffix(L.dtype), nat.i64(n), nat.i64(L.nnz), L.indptr,
                 L.indices, L.data, nat.f64(lmax), c, nat.i32(nscales), nat.i32(M), x,
                 nat.i64(nsig), r, work, plan, nat.stream_ptr(L.device))
    return r


def cheby_clenshaw_device(L, lmax, c, sources, out=None, work=None):
    """sum_i p_i(L) s_i by ONE backward (Clenshaw) recurrence, device to device.

    ``sources``: (nsrc, N, nsig) tensor (or (N, nsig) for a single filter), ``c``:
    (nsrc, M) coefficients.  With one source this is the single-filter filtering of
    :func:`cheby_op_device` at 4 instead of 5 passes over the signal block per order; with
    nsrc = Nf sources it is the *synthesis* of ``Filter.filter`` in K SpMMs instead of the
    reference's Nf * K (filter.py:313-322), because Clenshaw's recurrence is linear in its
    source term: b_k = sum_i c_ik s_i + 2 Lt b_{k+1} - b_{k+2}  (SURVEY.md 8f).
    Returns (N, nsig).
    """
    torch = nat.require_cuda()
    c = np.ascontiguousarray(np.atleast_2d(np.asarray(c, dtype=np.float64)))
    if c.shape[1] < 2:
        raise TypeError("The coefficients have an invalid shape")
    if sources.dim() == 2:
        sources = sources[None]
    nsrc, n, nsig = sources.shape
    if nsrc != c.shape[0]:
        raise ValueError("one coefficient row per source block")
    if nsrc > 16:
        raise ValueError("at most 16 source blocks per call")
    sources = sources.contiguous()
    if out is None:
        out = torch.empty((n, nsig), dtype=L.dtype, device=L.device)
    if work is None:
        work = torch.empty((2, n, nsig), dtype=L.dtype, device=L.device)
    plan = L.tile_plan(nsig, nsrc)
    with torch.cuda.device(L.device):
        nat.call("gsp_cheby_clenshaw_" + nat.suffix(L.dtype), nat.i64(n), nat.i64(L.nnz),
                 L.indptr, L.indices, L.data, nat.f64(lmax), c, nat.i32(nsrc),
                 nat.i32(c.shape[1]), sources, nat.i64(nsig), out, work, plan,
                 nat.stream_ptr(L.device))
    return out


def cheby_op(G, c, signal, **kwargs):
    r"""Chebyshev polynomial of the graph Laplacian applied to a signal block.

    Same contract as the reference (approximations.py:58-114): ``c`` is one
    coefficient vector or an (Nscales, M) array / list of vectors, ``signal``
    is (N,) or (N, Nsig); the result is (Nscales*N,) or (Nscales*N, Nsig) with
    filter-major row blocks.  ``M < 2`` raises TypeError.  NumPy in -> NumPy
    out, CUDA tensor in -> CUDA tensor out.  The arithmetic type is the
    graph's (float32 by default; the reference always computes in float64).
    A single filter is evaluated by Clenshaw's backward recurrence (one pass less over the
    signal block per order, same value, different rounding); ``clenshaw=False`` keeps the
    reference's forward recurrence and operation order.
    """
    if not isinstance(c, np.ndarray):
        c = np.array(c)
    c = np.atleast_2d(c)
    if c.shape[1] < 2:
        raise TypeError("The coefficients have an invalid shape")
    L = _laplacian_on_device(G)
    x, one_d, kind = _as_device_block(_GraphView(L), signal)
    if x.shape[0] != G.N:
        raise ValueError("First dimension must be the number of vertices "
                         "G.N = {}, got {}.".format(G.N, tuple(x.shape)))
    clenshaw = kwargs.get("clenshaw", None)
    if clenshaw and c.shape[0] != 1:
        raise ValueError("clenshaw=True evaluates a single filter")
    if clenshaw is None:
        clenshaw = c.shape[0] == 1
    if clenshaw:
        r = cheby_clenshaw_device(L, G.lmax, c[0], x)
    else:
        r = cheby_op_device(L, G.lmax, c, x)
    r = r.reshape(c.shape[0] * G.N, x.shape[1])
    if one_d:
        r = r.reshape(-1)
    out = _leave_device(r, kind)
    if kind == "numpy" and not isinstance(G.L, type(L)):
        out = out.astype(np.float64, copy=False)     # a reference graph expects float64 back
    return out


class _GraphView:
    def __init__(self, L):
        self.device, self.dtype = L.device, L.dtype


def cheby_rect(G, bounds, signal, **kwargs):
    r"""Ideal band-pass [bounds[0], bounds[1]] by closed-form Chebyshev coefficients.

    Reference: approximations.py:117-163.  The expansion coefficients of the
    rectangle are c_0/2 = (b1-b2)/pi, c_k = 2/(k pi) (sin k b1 - sin k b2) with
    b = arccos(2 bounds / lmax - 1); the recurrence is the one of ``cheby_op``,
    so the same fused kernel is used with these coefficients.
    """
    if not (isinstance(bounds, (list, np.ndarray)) and len(bounds) == 2):
        raise ValueError("Bounds of wrong shape.")
    bounds = np.array(bounds, dtype=np.float64)
    order = int(kwargs.pop("order", 30))
    b1, b2 = np.arccos(2.0 * bounds / G.lmax - 1.0)
    k = np.arange(1, order + 1)
    c = np.empty(order + 1)
    c[0] = 2.0 * (b1 - b2) / np.pi
    c[1:] = 2.0 / (k * np.pi) * (np.sin(k * b1) - np.sin(k * b2))
    return cheby_op(G, c, signal)


def compute_jackson_cheby_coeff(filter_bounds, delta_lambda, m):
    r"""Chebyshev and Jackson-damped coefficients of an ideal band-pass.

    Reference: approximations.py:166-225.  For the band [a, b] inside
    [lambda_min, lambda_max], mapped to [-1, 1]: ``ch[0] = 2/pi (acos a' - acos b')``,
    ``ch[i] = 2/(pi i) (sin(i acos a') - sin(i acos b'))``; the Jackson factors
    ``g_i = ((1 - i/(m+2)) sin(t) cos(i t) + cos(t) sin(i t)/(m+2)) / sin(t)``,
    ``t = pi/(m+2)``, damp the Gibbs oscillations.  Returns ``(ch, ch * g)``; either
    feeds :func:`cheby_op` directly (host code, m+1 numbers).  Unlike the reference
    the caller's ``filter_bounds`` list is not modified.
    """
    lo, hi = float(delta_lambda[0]), float(delta_lambda[1])
    a, b = float(filter_bounds[0]), float(filter_bounds[1])
    if lo > a or hi < b:
        raise ValueError("Bounds of the filter are out of the lambda values")
    if lo > hi:
        raise ValueError("lambda_min is greater than lambda_max")
    half, mid = (hi - lo) / 2, (hi + lo) / 2
    ta, tb = np.arccos((a - mid) / half), np.arccos((b - mid) / half)
    i = np.arange(1, m + 1)
    ch = np.empty(m + 1)
    ch[0] = 2 / np.pi * (ta - tb)
    ch[1:] = 2 / (np.pi * i) * (np.sin(i * ta) - np.sin(i * tb))
    j = np.arange(m + 1)
    t = np.pi / (m + 2)
    damp = ((1 - j / (m + 2)) * np.sin(t) * np.cos(j * t) + np.cos(t) * np.sin(j * t) / (m + 2)) / np.sin(t)
    return ch, ch * damp
