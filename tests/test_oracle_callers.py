"""The oracle's restatement of the callers of the filtering path (reduction.interpolate,
pyramid_analysis / pyramid_synthesis, learning.regression_tikhonov) against fixtures produced
by the real PyGSP 0.6.1 (tests/golden/make_golden_r2.py).  CPU only."""
import numpy as np
import pytest

from conftest import csr_from
from oracle import pygsp_oracle as orc


@pytest.fixture(scope="module")
def pyr(golden):
    z = golden("pyramid")
    levels = int(z["levels"])
    Ws = [csr_from(z, "W%d" % i) for i in range(levels + 1)]
    Ls = [orc.laplacian(W) for W in Ws]
    lmaxs = [float(z["lmax%d" % i]) for i in range(levels + 1)]
    idxs = [z["idx%d" % (i + 1)] for i in range(levels)]
    return z, levels, Ls, lmaxs, idxs


def test_kron_reduction_golden(pyr):
    from scipy import sparse
    z, levels, Ls, lmaxs, idxs = pyr
    for i in range(levels):
        K = orc.kron_reduction(Ls[i] + 0.005 * sparse.eye(Ls[i].shape[0]), idxs[i])
        np.testing.assert_allclose(K.toarray(), z["Kreg%d" % i], rtol=1e-9, atol=1e-11)


def test_interpolate_golden(pyr):
    z, levels, Ls, lmaxs, idxs = pyr
    got = orc.interpolate(Ls[0], lmaxs[0], z["interp_in"], idxs[0])
    np.testing.assert_allclose(got, z["interp_out"], rtol=1e-9, atol=1e-9 * np.abs(z["interp_out"]).max())
    got3 = orc.interpolate(Ls[0], lmaxs[0], z["interp3_in"], idxs[0], order=60)
    np.testing.assert_allclose(got3, z["interp3_out"], rtol=1e-9, atol=1e-9 * np.abs(got3).max())


def test_pyramid_golden(pyr):
    z, levels, Ls, lmaxs, idxs = pyr
    order = int(z["order"])
    h = lambda x: 5.0 / (5 + x)
    ca, pe = orc.pyramid_analysis(Ls, lmaxs, idxs, z["f"], h, order=order)
    for i in range(levels + 1):
        np.testing.assert_allclose(ca[i], z["ca%d" % i], rtol=1e-8, atol=1e-9)
    for i in range(levels):
        np.testing.assert_allclose(pe[i], z["pe%d" % i], rtol=1e-7, atol=1e-8)
    rec, _ = orc.pyramid_synthesis(Ls, lmaxs, idxs, ca[levels], pe, order=order)
    np.testing.assert_allclose(rec, z["reconstruction"], rtol=1e-8, atol=1e-9)
    assert np.linalg.norm(rec - z["f"]) / np.linalg.norm(z["f"]) < 1e-10


def test_tikhonov_golden(golden):
    z = golden("tikhonov")
    L = orc.laplacian(csr_from(z, "W"))
    mask, tau = z["mask"].astype(bool), float(z["tau"])
    got = orc.regression_tikhonov(L, z["measures"], mask, tau)
    np.testing.assert_allclose(got, z["relaxed_exact"], rtol=1e-9, atol=1e-11)
    # the reference's own CG answer is only 1e-5 accurate (tests/test_learning.py:91)
    np.testing.assert_allclose(z["relaxed_reference_cg"], got, atol=1e-5)
    nan_measures = z["signal"].copy()
    nan_measures[~mask] = np.nan
    got0 = orc.regression_tikhonov(L, nan_measures, mask, 0)
    np.testing.assert_allclose(got0, z["constrained_reference"], rtol=1e-9, atol=1e-11)
    got1 = orc.regression_tikhonov(L, nan_measures[:, 0], mask, 0)
    np.testing.assert_allclose(got1, z["constrained_1d_reference"], rtol=1e-9, atol=1e-11)


def test_tikhonov_ring_kat():
    """tests/test_learning.py:11-19 of the reference: harmonic extension on a ring."""
    from scipy import sparse
    n = 8
    idx = np.arange(n)
    W = sparse.csr_matrix((np.ones(2 * n), (np.r_[idx, idx], np.r_[(idx + 1) % n, (idx - 1) % n])),
                          shape=(n, n))
    signal = np.array([0, np.nan, 4, np.nan, 4, np.nan, np.nan, np.nan])
    mask = np.array([True, False, True, False, True, False, False, False])
    got = orc.regression_tikhonov(orc.laplacian(W), signal, mask, 0)
    np.testing.assert_allclose(got, [0, 2, 4, 4, 4, 3, 2, 1])
