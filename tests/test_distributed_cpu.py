"""Host logic of the vertex-partitioned path on CPU: world_size 2 and 3, gloo backend.
The CUDA step kernel is replaced by a NumPy stand-in built on the oracle's CSR product
(tests/dist_worker.py); the partition plan, the halo id exchange, the per-step
all-to-all-v and the boundary/interior split are the product code."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(backend, world, n, nsig, nscales, order, overlap=0, timeout=300,
              worker="dist_worker.py"):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, worker), backend,
                               str(world), str(r), str(port), str(n), str(nsig), str(nscales),
                               str(order), str(overlap)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-1500:])
        assert "rank %d ok" % r in o
    return outs


@pytest.mark.parametrize("world,nsig,nscales", [(2, 3, 1), (3, 4, 2)])
def test_partitioned_cheby_gloo(world, nsig, nscales):
    run_world("gloo", world, n=900, nsig=nsig, nscales=nscales, order=12)


def test_plan_single_rank_is_identity_like():
    from scipy import sparse
    from pygsp_b200 import distributed as gd
    A = sparse.random(50, 50, 0.1, random_state=0, format="csr")
    A = (A + A.T).tocsr()
    plan = gd.HaloPlan(A, gd.even_bounds(50, 1), 0,
                       exchange_ids=lambda ids, rc, rank, P, group: (np.zeros(0, np.int64), np.zeros(1, np.int64)))
    assert plan.n_halo == 0 and plan.n_boundary == 0
    np.testing.assert_array_equal(plan.perm, np.arange(50))
    np.testing.assert_array_equal(plan.indices, A.indices)
    np.testing.assert_array_equal(gd.even_bounds(10, 4), [0, 2, 5, 7, 10])


def test_sensor_strips_equal_global_knn_graph():
    """The per-rank strip generator reproduces the k-NN graph of the union of strips."""
    from scipy import sparse, spatial
    from oracle import pygsp_oracle as orc
    from pygsp_b200.graphs.generators import SensorStrips, laplacian_rows
    P, n_per, k = 3, 4000, 6
    gens = [SensorStrips(r, P, n_per, k=k, seed=5) for r in range(P)]
    tot = [g.distance_sum() for g in gens]
    sigma = sum(t[0] for t in tot) / sum(t[1] for t in tot)
    rows = [g.adjacency_rows(sigma) for g in gens]
    Wd = sparse.vstack(rows).tocsr()
    pts = np.concatenate([g.coords for g in gens])
    D, NN = spatial.cKDTree(pts).query(pts, k=k + 1)
    assert abs(D[:, 1:].mean() - sigma) < 1e-15
    A = sparse.csr_matrix((np.exp(-D[:, 1:].ravel() ** 2 / sigma),
                           (np.repeat(np.arange(P * n_per), k), NN[:, 1:].ravel())),
                          shape=(P * n_per, P * n_per))
    Wg = ((A + A.T) / 2).tocsr()
    Wg.sort_indices()
    np.testing.assert_array_equal(Wd.indptr, Wg.indptr)
    np.testing.assert_array_equal(Wd.indices, Wg.indices)
    np.testing.assert_allclose(Wd.data, Wg.data, rtol=1e-13)
    Lg = orc.laplacian(Wg)
    for r in range(P):
        Lr, _ = laplacian_rows(rows[r], r * n_per)
        ref = Lg[r * n_per:(r + 1) * n_per]
        np.testing.assert_array_equal(Lr.indices, ref.indices)
        np.testing.assert_allclose(Lr.data, ref.data, rtol=1e-12)


@pytest.mark.parametrize("dim,P,n_per,k", [(2, 3, 3000, 6), (3, 2, 6000, 8), (3, 1, 2000, 5)])
def test_knn_slabs_equal_global_knn_graph(dim, P, n_per, k):
    """Config-5 generator (host backend): the slab row blocks are exactly the rows of the k-NN
    graph of the union of all slabs, NNGraph's weights and symmetrisation
    (pygsp/graphs/nngraphs/nngraph.py:213-226,289-297)."""
    from scipy import sparse, spatial
    from oracle import pygsp_oracle as orc
    from pygsp_b200.graphs.generators import KnnSlabs, laplacian_rows
    gens = [KnnSlabs(r, P, n_per, dim=dim, k=k, seed=3, backend="host") for r in range(P)]
    tot = [g.distance_sum() for g in gens]
    sigma = sum(t[0] for t in tot) / sum(t[1] for t in tot)
    rows = [g.adjacency_rows(sigma) for g in gens]
    Wd = sparse.vstack(rows).tocsr()
    pts = np.concatenate([g.coords for g in gens])
    assert pts.shape == (P * n_per, dim) and pts.min() >= 0 and pts.max() < 1
    D, NN = spatial.cKDTree(pts).query(pts, k=k + 1)
    assert abs(D[:, 1:].mean() - sigma) < 1e-14
    A = sparse.csr_matrix((np.exp(-D[:, 1:].ravel() ** 2 / sigma),
                           (np.repeat(np.arange(P * n_per), k), NN[:, 1:].ravel())),
                          shape=(P * n_per, P * n_per))
    Wg = ((A + A.T) / 2).tocsr()
    Wg.sort_indices()
    np.testing.assert_array_equal(Wd.indptr, Wg.indptr)
    np.testing.assert_array_equal(Wd.indices, Wg.indices)
    np.testing.assert_allclose(Wd.data, Wg.data, rtol=1e-13)
    Lg = orc.laplacian(Wg)
    for r in range(P):
        Lr, _ = laplacian_rows(rows[r], r * n_per)
        ref = Lg[r * n_per:(r + 1) * n_per]
        np.testing.assert_array_equal(Lr.indices, ref.indices)
        np.testing.assert_allclose(Lr.data, ref.data, rtol=1e-12)


def test_knn_slabs_reject_thin_slabs():
    from pygsp_b200.graphs.generators import KnnSlabs
    with pytest.raises(ValueError):
        KnnSlabs(0, 8, 50, dim=3, k=16, seed=0, backend="host")


@pytest.mark.parametrize("parts,density", [(3, 0.02), (1, 0.02), (4, 0.3), (2, 0.0)])
def test_halo_plan_from_tensors_equals_host_plan(parts, density):
    """HaloPlan.from_device (torch ops, here on CPU tensors) builds the plan of HaloPlan():
    thin halos, a single part (no halo), a dense block (every row a boundary row), no edges."""
    import torch
    from scipy import sparse
    from pygsp_b200 import distributed as gd
    n = 400
    A = sparse.random(n, n, density, random_state=1, format="csr")
    A = (A + A.T).tocsr()
    A.sort_indices()
    bounds = gd.even_bounds(n, parts)
    fake = lambda h, rc, rank, P, g: (np.zeros(0, dtype=np.int64), np.zeros(P, dtype=np.int64))
    for r in range(parts):
        rows = A[bounds[r]:bounds[r + 1]]
        p1 = gd.HaloPlan(rows, bounds, r, exchange_ids=fake)
        p2 = gd.HaloPlan.from_device(torch.from_numpy(rows.indptr), torch.from_numpy(rows.indices),
                                     torch.from_numpy(rows.data), bounds, r, exchange_ids=fake)
        for name in ("halo_ids", "recv_counts", "perm", "inv_perm", "indptr", "indices", "data"):
            a, b = getattr(p1, name), getattr(p2, name)
            np.testing.assert_array_equal(a, b.numpy() if torch.is_tensor(b) else b, err_msg=name)
        assert (p1.n_boundary, p1.n_true_boundary, p1.n_halo, p1.nnz) == \
            (p2.n_boundary, p2.n_true_boundary, p2.n_halo, p2.nnz)
