"""pytest plugin used by test_reference_dropin_gpu.py: runs the REFERENCE's own
test-suite with its cheby_op rebound to the CUDA engine (float64)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

CALLS = {"n": 0}


def pytest_configure(config):
    import torch
    import pygsp.filters.approximations as ref
    import pygsp_b200
    pygsp_b200.patch_pygsp(dtype=torch.float64)
    engine = ref.cheby_op

    def counted(G, c, signal, **kw):
        CALLS["n"] += 1
        return engine(G, c, signal, **kw)
    ref.cheby_op = counted


def pytest_unconfigure(config):
    print("\nGSPB200_ENGINE_CALLS=%d" % CALLS["n"])
