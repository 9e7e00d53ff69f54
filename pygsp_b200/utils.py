"""Host utilities of the filtering path (mirror of the used part of pygsp/utils.py)."""
import functools
import logging

import numpy as np
from scipy import sparse


def build_logger(name):
    """Per-module logger with the reference's format (pygsp/utils.py:16-31)."""
    logger = logging.getLogger(name)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setLevel(logging.DEBUG)
        handler.setFormatter(logging.Formatter(
            "%(asctime)s:[%(levelname)s](%(name)s.%(funcName)s): %(message)s"))
        logger.setLevel(logging.DEBUG)
        logger.addHandler(handler)
    return logger


def filterbank_handler(func):
    """Call ``func`` once per filter of a bank (pygsp/utils.py:37-53).

    With ``i=`` given, or a single filter, the call goes straight through;
    otherwise the results for i = 0..Nf-1 are collected in a list.
    """
    @functools.wraps(func)
    def wrapper(f, *args, **kwargs):
        if "i" in kwargs or f.Nf <= 1:
            return func(f, *args, **kwargs)
        return [func(f, *args, i=i, **kwargs) for i in range(f.Nf)]
    return wrapper


def compute_log_scales(lmin, lmax, Nscales, t1=1, t2=2):
    """Log-spaced wavelet scales from t2/lmin down to t1/lmax (pygsp/utils.py:312-339)."""
    return np.exp(np.linspace(np.log(t2 / lmin), np.log(t1 / lmax), Nscales))


def symmetrize(W, method="average"):
    """Host-side symmetrisation used by the graph generators (pygsp/utils.py:184-277).

    Only the variants the generators on the path need: 'average' ((W+W^T)/2),
    'maximum', 'tril' and 'triu' (mirror one triangle).
    """
    if W.shape[0] != W.shape[1]:
        raise ValueError("Matrix must be square.")
    if method == "average":
        return (W + W.T) / 2
    if method == "maximum":
        if sparse.issparse(W):
            return W.maximum(W.T)
        return np.maximum(W, W.T)
    if method in ("tril", "triu"):
        tri = getattr(sparse if sparse.issparse(W) else np, method)(W)
        return symmetrize(tri, "maximum")
    raise ValueError("Unknown symmetrization method {}.".format(method))
