"""Callers of the Chebyshev filtering path in ``pygsp/reduction.py``.

``interpolate`` (reduction.py:150-193), ``pyramid_analysis`` (:384-449) and the direct branch
of ``pyramid_synthesis`` (:504-514): every step of the Kron pyramid is a Chebyshev filter --
the analysis filter ``h`` and the order-100 Green kernel ``1 / (eps + x)`` of the
interpolation -- and runs on the CUDA engine through :class:`pygsp_b200.filters.Filter`.
What is NOT the filtering path stays host-side set-up, as in the reference: building the
multiresolution sequence itself (``graph_multiresolution``: eigenvector-based down-sampling,
Kron reduction, sparsification) and the Schur complement ``K_reg`` of a level
(:func:`kron_reduction`, a sparse direct solve, computed once per level).

A graph of the sequence carries ``G.mr = {'idx': kept vertices of the level above,
'K_reg': ...}`` like the reference's.  Shapes: the reference keeps consistent shapes only for
column-vector signals (a 1-D signal is broadcast to (N, N) at reduction.py:447); here a signal is
(N,) or (N, Nv), every coefficient block is returned 2-D (n_level, Nv), and Nv columns mean Nv
independent signals.  The least-squares synthesis (reduction.py:534-630) is broken in the
reference (NameError at :593) and has no oracle; ``least_squares=True`` raises.
"""
import numpy as np
from scipy import sparse

from . import filters
from . import utils

logger = utils.build_logger(__name__)


def kron_reduction(L, ind):
    """Kron reduction (Schur complement) of a Laplacian-like sparse matrix onto ``ind``
    (reduction.py:309-382, matrix branch).  Host-side, once per pyramid level."""
    from scipy.sparse import linalg
    if hasattr(L, "to_scipy"):
        L = L.to_scipy()
    L = sparse.csr_matrix(L, dtype=np.float64)
    n = L.shape[0]
    ind = np.asarray(ind)
    rest = np.setdiff1d(np.arange(n, dtype=int), ind)
    inner = L[rest][:, rest].tocsc()
    coupling = L[rest][:, ind].tocsc()
    schur = L[ind][:, ind] - L[ind][:, rest].dot(linalg.spsolve(inner, coupling))
    schur = sparse.csr_matrix(schur)
    if np.abs(schur - schur.T).sum() < np.spacing(1) * np.abs(schur).sum():
        schur = (schur + schur.T) / 2.0
    return sparse.csr_matrix(schur)


def _as_columns(s):
    s = np.asarray(s) if not hasattr(s, "is_cuda") else s
    return s.reshape(s.shape[0], -1)


def _filter_columns(g, block, **kwargs):
    """One-filter bank applied to every column of an (N, Nv) block -> (N, Nv)."""
    out = g.filter(block if block.shape[1] > 1 else block[:, 0], **kwargs)
    return out.reshape(block.shape)


def _level_operator(G, reg_eps):
    """(K_reg, Green filter) of a level; cached in G.mr like graph_multiresolution does."""
    mr = getattr(G, "mr", None)
    if mr is None:
        mr = G.mr = {}
    if "green_kernel" not in mr or mr.get("_green_eps") != reg_eps:
        mr["green_kernel"] = filters.Filter(G, lambda x: 1.0 / (reg_eps + x))
        mr["_green_eps"] = reg_eps
    return mr


def interpolate(G, f_subsampled, keep_inds, order=100, reg_eps=0.005, **kwargs):
    r"""Interpolate a graph signal from its samples on ``keep_inds`` (reduction.py:150-193).

    ``alpha = K_reg f`` with ``K_reg`` the Kron reduction of ``L + eps I`` onto the samples
    (taken from ``G.mr['K_reg']`` when the multiresolution set-up stored it), zero-fill, then
    the Green kernel ``1 / (eps + x)`` as an order-``order`` Chebyshev filter on the device.
    Returns (N, Nv).
    """
    keep_inds = np.asarray(keep_inds)
    mr = _level_operator(G, reg_eps)
    K_reg = mr.get("K_reg")
    if K_reg is None or mr.get("_kreg_key") not in (None, (reg_eps, keep_inds.tobytes())):
        K_reg = kron_reduction(G.L.to_scipy().astype(np.float64) + reg_eps * sparse.eye(G.N),
                               keep_inds)
        mr["K_reg"], mr["_kreg_key"] = K_reg, (reg_eps, keep_inds.tobytes())
    sub = _as_columns(np.asarray(f_subsampled, dtype=np.float64))
    if sub.shape[0] != keep_inds.size:
        raise ValueError("f_subsampled must have one row per kept vertex")
    full = np.zeros((G.N, sub.shape[1]))
    full[keep_inds] = K_reg.dot(sub)
    return _filter_columns(mr["green_kernel"], full, order=order, **kwargs)


def _level_filters(h_filters, levels):
    if not isinstance(h_filters, list):
        if callable(h_filters):
            logger.warning("Converting filters into a list.")
            h_filters = [h_filters]
        else:
            raise TypeError("Filters must be a list of functions.")
    if len(h_filters) == 1:
        h_filters = h_filters * levels
    elif len(h_filters) != levels:
        raise ValueError("The number of filters must be one or equal to {}.".format(levels))
    return h_filters


def pyramid_analysis(Gs, f, **kwargs):
    r"""Graph pyramid transform (reduction.py:384-449).

    Per level: low-pass ``h(L) ca_i`` (Chebyshev filter on the device), keep the vertices
    of the next level, interpolate them back (:func:`interpolate`), prediction error
    ``pe_i = ca_i - interpolation``.  ``h_filters``: list of kernels (default
    ``1 / (2x + 1)``); remaining keyword arguments (``order`` ...) go to the filters, as in the
    reference.  Returns ``(ca, pe)``: lists of (n_level, Nv) arrays.
    """
    f = np.asarray(f)
    if f.shape[0] != Gs[0].N:
        raise ValueError("PYRAMID ANALYSIS: The signal to analyze should have the same "
                         "dimension as the first graph.")
    levels = len(Gs) - 1
    h_filters = _level_filters(kwargs.pop("h_filters", lambda x: 1.0 / (2 * x + 1)), levels)
    ca, pe = [_as_columns(f.astype(np.float64))], []
    for i in range(levels):
        idx = np.asarray(Gs[i + 1].mr["idx"])
        s_low = _filter_columns(filters.Filter(Gs[i], h_filters[i]), ca[i], **kwargs)
        ca.append(s_low[idx])
        s_pred = interpolate(Gs[i], ca[i + 1], idx, **kwargs)
        pe.append(ca[i] - s_pred)
    return ca, pe


def pyramid_synthesis(Gs, cap, pe, order=30, **kwargs):
    r"""Signal from its pyramid coefficients, direct method (reduction.py:452-532).

    From the coarsest approximation up: interpolate to the finer level (order-``order`` Green
    kernel filter on the device) and add that level's prediction error.  Returns
    ``(reconstruction, ca)``.
    """
    if bool(kwargs.pop("least_squares", False)):
        raise NotImplementedError(
            "least-squares pyramid synthesis is broken in the reference (reduction.py:593) and "
            "is not part of this engine; use the direct method.")
    kwargs.pop("use_landweber", None)
    reg_eps = float(kwargs.pop("reg_eps", 0.005))
    levels = len(Gs) - 1
    if len(pe) != levels:
        raise ValueError("Gs and pe have different shapes.")
    ca = [_as_columns(np.asarray(cap, dtype=np.float64))]
    for i in range(levels):
        lv = levels - i - 1
        s_pred = interpolate(Gs[lv], ca[i], np.asarray(Gs[lv + 1].mr["idx"]), order=order,
                             reg_eps=reg_eps, **kwargs)
        ca.append(s_pred + _as_columns(np.asarray(pe[lv])))
    return ca[levels], ca
