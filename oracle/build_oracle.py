"""Compiles oracle/cheby_oracle.c with gcc into oracle/_build/libcheby_oracle.so
(test infrastructure; building the checker is not using it)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcheby_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "cheby_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared",
                               src, "-o", LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def cheby_op(L, lmax, c, signal):
    """float64 C oracle of approximations.cheby_op; L is a scipy CSR matrix."""
    c = np.ascontiguousarray(np.atleast_2d(np.asarray(c, dtype=np.float64)))
    x = np.ascontiguousarray(np.asarray(signal, dtype=np.float64))
    n = L.shape[0]
    x2 = x.reshape(n, -1)
    nsig = x2.shape[1]
    r = np.empty((c.shape[0] * n, nsig))
    Ap = np.ascontiguousarray(L.indptr, dtype=np.int32)
    Aj = np.ascontiguousarray(L.indices, dtype=np.int32)
    Ax = np.ascontiguousarray(L.data, dtype=np.float64)
    P = ctypes.c_void_p
    rc = lib().cheby_op(ctypes.c_int64(n), P(Ap.ctypes.data), P(Aj.ctypes.data), P(Ax.ctypes.data),
                        ctypes.c_double(lmax), P(c.ctypes.data), ctypes.c_int64(c.shape[0]),
                        ctypes.c_int64(c.shape[1]), P(x2.ctypes.data), ctypes.c_int64(nsig),
                        P(r.ctypes.data))
    if rc == -1:
        raise TypeError("The coefficients have an invalid shape")
    if rc != 0:
        raise MemoryError("cheby_oracle")
    return r.reshape((c.shape[0] * n,) + x.shape[1:])


if __name__ == "__main__":
    print(build(force=True))
