"""Pins oracle/pygsp_oracle.py against fixtures produced by the real PyGSP 0.6.1."""
import numpy as np
import pytest

from conftest import csr_from, relerr_cols
from oracle import pygsp_oracle as orc


def _same_csr(A, B, rtol=1e-13):
    assert A.shape == B.shape
    np.testing.assert_array_equal(A.indptr, B.indptr)
    np.testing.assert_array_equal(A.indices, B.indices)
    np.testing.assert_allclose(A.data, B.data, rtol=rtol, atol=0)


def test_laplacian_kats(golden):
    g = golden("laplacian_kat")
    for name in g["names"]:
        W = orc.canonical_adjacency(g[name + "_A"])
        directed = orc.is_directed(W)
        assert directed == bool(g[name + "_directed"]), name
        assert orc.count_edges(W, directed) == int(g[name + "_n_edges"]), name
        np.testing.assert_allclose(orc.weighted_degree(W, directed), g[name + "_dw"], rtol=1e-14)
        np.testing.assert_allclose(orc.degree(W, directed), g[name + "_d"])
        for lap in ("combinatorial", "normalized"):
            _same_csr(orc.laplacian(W, lap), csr_from(g, name + "_L" + lap[0]))
            ref_bound = float(g[name + "_bound_" + lap[0]])
            got = orc.upper_bound(W, lap)
            if np.isnan(ref_bound):
                assert np.isnan(got), name
            else:
                np.testing.assert_allclose(got, ref_bound, rtol=1e-13, err_msg=name)


def test_reference_kat_values(golden):
    # pygsp/tests/test_graphs.py:195-230 -- the literal matrices
    lap = np.array([[4, -3, 0, -1], [-3, 4, -1, 0], [0, -1, 4, -3], [-1, 0, -3, 4.]])
    for A in ([[0, 3, 0, 1], [3, 0, 1, 0], [0, 1, 0, 3], [1, 0, 3, 0]],
              [[0, 6, 0, 1], [0, 0, 0, 0], [0, 2, 0, 3], [1, 0, 3, 0]]):
        W = orc.canonical_adjacency(A)
        np.testing.assert_allclose(orc.laplacian(W, "combinatorial").toarray(), lap)
        np.testing.assert_allclose(orc.laplacian(W, "normalized").toarray(), lap / 4)
    # test_graphs.py:257-294 -- tight bounds
    assert orc.upper_bound(orc.canonical_adjacency(np.full((10, 10), 2))) == pytest.approx(20)
    bip = [[0, 0, 1, 1], [0, 0, 1, 1], [1, 1, 0, 0], [1, 1, 0, 0]]
    assert orc.upper_bound(orc.canonical_adjacency(bip)) == pytest.approx(4)
    assert orc.lambda_max_exact(orc.laplacian(orc.canonical_adjacency(bip))) == pytest.approx(4)


@pytest.mark.parametrize("fixture,prefix", [("logo", "logo"), ("sensor123", "s"), ("grid13x9", "g")])
def test_laplacian_fixtures(golden, fixture, prefix):
    g = golden(fixture)
    W = csr_from(g, "W")
    for lap in ("combinatorial", "normalized"):
        _same_csr(orc.laplacian(W, lap), csr_from(g, prefix + "_L" + lap[0]))
        np.testing.assert_allclose(orc.upper_bound(W, lap), float(g[prefix + "_bound_" + lap[0]]), rtol=1e-13)


def test_logo_lmax(golden):
    g = golden("logo")
    L = csr_from(g, "logo_Lc")
    assert orc.lambda_max_exact(L) == pytest.approx(float(g["lmax_exact"]), rel=1e-10)
    lo, hi = orc.lmax_lanczos_band(L)
    assert lo <= float(g["lmax_lanczos"]) <= hi          # graph.py:891-899: 13.78 / 13.92 / 18.58
    assert "{:.2f}".format(float(g["lmax_exact"])) == "13.78"
    assert "{:.2f}".format(float(g["logo_bound_c"])) == "18.58"


def test_readme_example(golden):
    g = golden("logo")
    L = csr_from(g, "logo_Lc")
    lmax = float(g["lmax_lanczos"])
    c = orc.cheby_coeff(orc.heat_kernels(lmax, 50), lmax, 30)
    np.testing.assert_allclose(c[0], g["heat50_coeff"], rtol=1e-10, atol=1e-14)
    for spmm in ("scipy", "numpy"):
        y = orc.cheby_op(L, lmax, c, g["readme_signal"], spmm=spmm)
        np.testing.assert_allclose(y, g["readme_filtered"], rtol=1e-11, atol=1e-14)
    y = orc.filter_signal(L, lmax, orc.heat_kernels(lmax, 50), g["readme_signal"])
    np.testing.assert_allclose(y, g["readme_filtered"], rtol=1e-11, atol=1e-14)
    r = orc.cheby_rect(L, lmax, g["rect_bounds"], g["rect_signal"], order=25)
    np.testing.assert_allclose(r, g["rect_filtered"], rtol=1e-10, atol=1e-13)


def test_sensor123(golden):
    g = golden("sensor123")
    L = csr_from(g, "s_Lc")
    lmax = float(g["lmax"])
    heat = orc.heat_kernels(lmax, 10)
    np.testing.assert_allclose(orc.cheby_coeff(heat, lmax, 30)[0], g["heat10_coeff"], rtol=1e-10, atol=1e-14)
    y = orc.filter_signal(L, lmax, heat, g["signal"])
    np.testing.assert_allclose(y, g["heat10_cheb"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(y, g["heat10_exact"], rtol=1e-7)       # test_filters.py:403-417
    # frame of Heat([8, 9]) (test_filters.py:157-168): filter the identity
    k89 = orc.heat_kernels(lmax, [8, 9])
    np.testing.assert_allclose(orc.cheby_coeff(k89, lmax, 30), g["heat89_coeff"], rtol=1e-10, atol=1e-14)
    n = L.shape[0]
    F = orc.cheby_op(L, lmax, orc.cheby_coeff(k89, lmax, 30), np.identity(n))
    np.testing.assert_allclose(F, g["heat89_frame"], rtol=1e-10, atol=1e-13)
    mh = orc.mexican_hat_kernels(lmax, Nf=5)
    c = orc.cheby_coeff(mh, lmax, 40)
    np.testing.assert_allclose(c, g["mh5_coeff"], rtol=1e-11, atol=1e-14)
    for spmm in ("scipy", "numpy"):
        np.testing.assert_allclose(orc.cheby_op(L, lmax, c, g["mh5_block"], spmm=spmm),
                                   g["mh5_cheby_op"], rtol=1e-10, atol=1e-12)
    a = orc.filter_signal(L, lmax, mh, g["mh5_block"], order=40)
    np.testing.assert_allclose(a, g["mh5_analysis"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.filter_signal(L, lmax, mh, a, order=40), g["mh5_synthesis"],
                               rtol=1e-10, atol=1e-12)
    loc = np.zeros(n); loc[7] = 1
    np.testing.assert_allclose(np.sqrt(n) * orc.filter_signal(L, lmax, heat, loc, order=25),
                               g["localize_7"], rtol=1e-10, atol=1e-13)


def test_shape_truth_table(golden):
    g = golden("sensor123")
    L = csr_from(g, "s_Lc")
    lmax = float(g["lmax"])
    mh = orc.mexican_hat_kernels(lmax, Nf=5)
    heat = orc.heat_kernels(lmax, 10)
    for j in range(9):
        x = g["tt%d_in" % j]
        y = orc.filter_signal(L, lmax, mh, x, order=20)
        assert y.shape == g["tt%d_mh5" % j].shape, x.shape
        np.testing.assert_allclose(y, g["tt%d_mh5" % j], rtol=1e-10, atol=1e-12)
        if "tt%d_heat" % j in g:
            y = orc.filter_signal(L, lmax, heat, x, order=20)
            assert y.shape == g["tt%d_heat" % j].shape
            np.testing.assert_allclose(y, g["tt%d_heat" % j], rtol=1e-10, atol=1e-12)
    n = L.shape[0]
    with pytest.raises(ValueError):
        orc.filter_signal(L, lmax, mh, np.zeros((n, 3, 2)))
    with pytest.raises(ValueError):
        orc.filter_signal(L, lmax, mh, np.zeros((n, 3, 1, 1)))
    with pytest.raises(ValueError):
        orc.filter_signal(L, lmax, mh, np.zeros((n + 1,)))
    with pytest.raises(TypeError):
        orc.filter_signal(L, lmax, heat, np.zeros(n), order=0)
    assert orc.filter_signal(L, lmax, heat, np.ones(n), order=1).shape == (n,)


def test_doctest_golden(golden):
    g = golden("doctest")
    W = csr_from(g, "W")
    L = orc.laplacian(W)
    lmax = float(g["lmax"])
    s1 = np.zeros(30); s1[13] = 1
    s1 = orc.filter_signal(L, lmax, orc.heat_kernels(lmax, 3), s1)
    mh = orc.mexican_hat_kernels(lmax, Nf=4)
    s2 = orc.filter_signal(L, lmax, mh, s1)
    assert s2.shape == (30, 4)
    s3 = orc.filter_signal(L, lmax, mh, s2)
    assert "{:.5f}".format(np.linalg.norm(s1 - s3)) == "0.27649"      # filter.py:255-256
    np.testing.assert_allclose(s3, g["s3"], rtol=1e-10, atol=1e-13)
    R = orc.laplacian(csr_from(g, "ringW"))
    rl = float(g["ring_lmax"])
    y = orc.filter_signal(R, rl, orc.heat_kernels(rl, [1, 10, 100]), g["ring_signal"])
    assert y.shape == (60, 10, 3)                                      # filter.py:217-219
    np.testing.assert_allclose(y, g["ring_filtered"], rtol=1e-10, atol=1e-13)


def test_grid_bank(golden):
    g = golden("grid13x9")
    L = csr_from(g, "g_Lc")
    lmax = float(g["lmax"])
    mh = orc.mexican_hat_kernels(lmax, Nf=6)
    np.testing.assert_allclose(orc.cheby_coeff(mh, lmax, 50), g["coeff"], rtol=1e-10, atol=1e-13)
    y = orc.filter_signal(L, lmax, mh, g["signal"], order=50)
    assert relerr_cols(y, g["filtered"]) < 1e-11


def test_c_oracle_matches(golden):
    """oracle/cheby_oracle.c (plain C restatement) == numpy oracle == PyGSP goldens."""
    from oracle import build_oracle
    g = golden("sensor123")
    L = csr_from(g, "s_Lc")
    lmax = float(g["lmax"])
    y = build_oracle.cheby_op(L, lmax, g["mh5_coeff"], g["mh5_block"])
    np.testing.assert_allclose(y, g["mh5_cheby_op"], rtol=1e-10, atol=1e-12)
    y1 = build_oracle.cheby_op(L, lmax, g["heat10_coeff"], g["signal"])
    np.testing.assert_allclose(y1, g["heat10_cheb"], rtol=1e-11, atol=1e-14)
    with pytest.raises(TypeError):
        build_oracle.cheby_op(L, lmax, [1.0], g["signal"])
