"""Callers of the filtering path on the CUDA engine: reduction.interpolate,
pyramid_analysis / pyramid_synthesis and learning.regression_tikhonov against the fixtures of
the real reference and the oracle."""
import numpy as np
import pytest

from conftest import csr_from, relerr_cols
from oracle import pygsp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gsp():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pygsp_b200
    return pygsp_b200


def _pyramid_graphs(gsp, z, dtype):
    levels = int(z["levels"])
    Gs = []
    for i in range(levels + 1):
        G = gsp.graphs.Graph(csr_from(z, "W%d" % i), dtype=dtype)
        G._lmax, G._lmax_method = float(z["lmax%d" % i]), "lanczos"
        G.mr = {}
        if i > 0:
            G.mr["idx"] = z["idx%d" % i]
        Gs.append(G)
    return Gs, levels


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-8), (np.float32, 2e-4)])
def test_interpolate_matches_reference(gsp, golden, dtype, tol):
    """Order-100 Green-kernel filter (values up to 1/eps = 200): float32 keeps 1e-4."""
    z = golden("pyramid")
    Gs, levels = _pyramid_graphs(gsp, z, dtype)
    got = gsp.reduction.interpolate(Gs[0], z["interp_in"], Gs[1].mr["idx"])
    assert got.shape == z["interp_out"].shape
    assert relerr_cols(got, z["interp_out"]) <= tol
    np.testing.assert_allclose(Gs[0].mr["K_reg"].toarray(), z["Kreg0"], rtol=1e-6, atol=1e-8)
    got3 = gsp.reduction.interpolate(Gs[0], z["interp3_in"], Gs[1].mr["idx"], order=60)
    assert relerr_cols(got3, z["interp3_out"]) <= tol


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-7), (np.float32, 5e-4)])
def test_pyramid_analysis_and_synthesis(gsp, golden, dtype, tol):
    z = golden("pyramid")
    Gs, levels = _pyramid_graphs(gsp, z, dtype)
    order = int(z["order"])
    h = [lambda x: 5.0 / (5 + x)]
    ca, pe = gsp.reduction.pyramid_analysis(Gs, z["f"], h_filters=h, order=order)
    scale = np.abs(z["f"]).max()
    for i in range(levels + 1):
        assert ca[i].shape == z["ca%d" % i].shape
        assert np.abs(ca[i] - z["ca%d" % i]).max() <= tol * scale
    for i in range(levels):
        assert np.abs(pe[i] - z["pe%d" % i]).max() <= tol * scale
    rec, cas = gsp.reduction.pyramid_synthesis(Gs, ca[levels], pe, order=order)
    # analysis followed by direct synthesis is the identity whatever the filters' accuracy
    assert np.linalg.norm(rec - z["f"]) / np.linalg.norm(z["f"]) <= (1e-10 if dtype == np.float64 else 1e-5)
    # several signals at once = the same pyramid column by column
    f2 = np.concatenate([z["f"], -2 * z["f"][::-1]], axis=1)
    ca2, pe2 = gsp.reduction.pyramid_analysis(Gs, f2, h_filters=h, order=order)
    assert np.abs(ca2[levels][:, :1] - ca[levels]).max() <= 10 * tol * scale
    with pytest.raises(ValueError):
        gsp.reduction.pyramid_analysis(Gs, z["f"][:-1], h_filters=h)
    with pytest.raises(NotImplementedError):
        gsp.reduction.pyramid_synthesis(Gs, ca[levels], pe, least_squares=True)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-8), (np.float32, 5e-5)])
def test_regression_tikhonov_relaxed_and_constrained(gsp, golden, dtype, tol):
    """tests/test_learning.py:28-95 of the reference: block solve == column-wise solve ==
    exact solution; the measurements are not modified."""
    z = golden("tikhonov")
    G = gsp.graphs.Graph(csr_from(z, "W"), dtype=dtype)
    mask, tau = z["mask"].astype(bool), float(z["tau"])
    measures = z["measures"].copy()
    got = gsp.learning.regression_tikhonov(G, measures, mask, tau=tau)
    np.testing.assert_array_equal(measures, z["measures"])
    assert np.abs(got - z["relaxed_exact"]).max() <= tol * np.abs(z["relaxed_exact"]).max()
    col = gsp.learning.regression_tikhonov(G, measures[:, 2], mask, tau)
    assert col.shape == (G.N,)
    assert np.abs(col - z["relaxed_exact"][:, 2]).max() <= tol * np.abs(z["relaxed_exact"]).max()
    nan_measures = z["signal"].copy()
    nan_measures[~mask] = np.nan
    got0 = gsp.learning.regression_tikhonov(G, nan_measures, mask, tau=0)
    assert np.abs(got0 - z["constrained_reference"]).max() <= tol * np.abs(z["constrained_reference"]).max()
    np.testing.assert_allclose(got0[mask], z["signal"][mask], rtol=1e-6)
    with pytest.raises(ValueError):
        gsp.learning.regression_tikhonov(G, nan_measures, mask[:-1], tau=0)


def test_regression_tikhonov_ring_kat_and_classification(gsp):
    """tests/test_learning.py:11-19 (ring KAT) and the logits wrapper (:98-123)."""
    G = gsp.graphs.Ring(N=8, dtype=np.float64)
    signal = np.array([0, np.nan, 4, np.nan, 4, np.nan, np.nan, np.nan])
    mask = np.array([True, False, True, False, True, False, False, False])
    got = gsp.learning.regression_tikhonov(G, signal, mask, tau=0)
    np.testing.assert_allclose(got, [0, 2, 4, 4, 4, 3, 2, 1], atol=1e-8)
    G = gsp.graphs.Sensor(300, k=6, seed=3, dtype=np.float64)
    rng = np.random.default_rng(0)
    labels = (G.coords[:, 0] > 0.5).astype(int) + (G.coords[:, 1] > 0.5).astype(int)
    mask = rng.uniform(size=G.N) > 0.6
    rec = gsp.learning.classification_tikhonov(G, labels, mask, tau=0)
    assert rec.shape == (G.N, 3)
    ref = orc.regression_tikhonov(orc.laplacian(G.W.to_scipy().astype(np.float64)),
                                  np.eye(3)[np.where(mask, labels, 0)] * mask[:, None], mask, 0)
    np.testing.assert_allclose(rec, ref, atol=1e-6)
    assert (np.argmax(rec, axis=1) == labels).mean() > 0.85


def test_block_cg_large(gsp):
    """1e5 vertices, 64 right-hand sides in one block solve, against the sparse direct solve."""
    import torch
    G = gsp.graphs.Sensor(100_000, k=8, seed=1, order="morton")
    rng = np.random.default_rng(2)
    y = rng.standard_normal((G.N, 64)).astype(np.float32)
    mask = rng.uniform(size=G.N) > 0.3
    got = gsp.learning.regression_tikhonov(G, torch.from_numpy(y).cuda(), mask, tau=2.0)
    assert got.is_cuda and got.shape == (G.N, 64)
    ref = orc.regression_tikhonov(orc.laplacian(G.W.to_scipy().astype(np.float64)), y[:, :4], mask, 2.0)
    assert relerr_cols(got[:, :4].cpu().numpy(), ref) <= 2e-5
