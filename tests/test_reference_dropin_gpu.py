"""Drop-in evidence with the REAL reference on the GPU box.

baseline/_ref holds the unmodified PyGSP 0.6.1 (offline `pip install --target`, git-ignored,
shipped with the snapshot).  When it is present: (1) stock `pygsp` objects filtered through
`patch_pygsp()` equal the stock SciPy results; (2) the reference's OWN test file
`pygsp/tests/test_filters.py` passes with its cheby_op rebound to the CUDA engine."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def pygsp_ref():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.isdir(os.path.join(REF, "pygsp")):
        pytest.skip("baseline/_ref (reference install) not present")
    sys.path.insert(0, REF)
    import logging
    import pygsp
    logging.getLogger("pygsp").setLevel(logging.ERROR)
    yield pygsp
    import pygsp_b200
    pygsp_b200.unpatch_pygsp()
    sys.path.remove(REF)


def test_patched_reference_objects(pygsp_ref):
    import torch
    import pygsp_b200
    pygsp = pygsp_ref
    G = pygsp.graphs.Logo()
    G.estimate_lmax()
    s = np.zeros(G.N); s[[20, 30, 1090]] = 1                 # README.rst:68-89
    bank = pygsp.filters.MexicanHat(G, Nf=5)
    heat = pygsp.filters.Heat(G, scale=50)
    block = np.random.default_rng(0).standard_normal((G.N, 7))
    want = [heat.filter(s), bank.filter(block, order=40), bank.filter(bank.filter(block), order=25)]
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-10)):
        pygsp_b200.patch_pygsp(dtype=dtype)
        if hasattr(G, "_gspb200_L"):
            del G._gspb200_L
        got = [heat.filter(s), bank.filter(block, order=40), bank.filter(bank.filter(block), order=25)]
        pygsp_b200.unpatch_pygsp()
        for a, b in zip(got, want):
            assert a.shape == b.shape and a.dtype == np.float64
            assert np.abs(a - b).max() / np.abs(b).max() <= tol
    # this engine's own Graph against the reference's Graph on the same adjacency
    H = pygsp_b200.graphs.Graph(G.W, dtype=np.float64)
    Lr = G.L.tocsr(); Lr.sort_indices()
    Lo = H.L.to_scipy()
    np.testing.assert_array_equal(Lo.indptr, Lr.indptr)
    np.testing.assert_array_equal(Lo.indices, Lr.indices)
    np.testing.assert_allclose(Lo.data, Lr.data, rtol=1e-13)
    assert H.n_edges == G.n_edges and abs(H._get_upper_bound() - G._get_upper_bound()) < 1e-9
    H.estimate_lmax()
    assert abs(H.lmax - G.lmax) / G.lmax < 2e-4              # ARPACK's own run-to-run spread is 1e-5


def test_reference_test_suite_on_cuda_engine(pygsp_ref):
    """pygsp/tests/test_filters.py, unmodified, with approximations.cheby_op -> CUDA (float64)."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), REF, ROOT]))
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "refsuite_plugin",
                          "-p", "no:cacheprovider", "-s",
                          os.path.join(REF, "pygsp", "tests", "test_filters.py")],
                         capture_output=True, text=True, env=env, cwd=REF, timeout=900)
    tail = (out.stdout + out.stderr)[-3000:]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "refsuite.log"), "w") as fh:
        fh.write(out.stdout + "\n--- stderr ---\n" + out.stderr)
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail
    import re
    calls = re.findall(r"GSPB200_ENGINE_CALLS=(\d+)", out.stdout)
    assert calls and int(calls[-1]) > 20, tail                   # the engine really served them
