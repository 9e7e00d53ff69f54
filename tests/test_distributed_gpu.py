"""Partitioned path on GPUs (NCCL).  The 2-rank test needs >= 2 devices and is skipped
on a single-GPU box; the 1-rank test runs the same code path without a process group."""
import numpy as np
import pytest

from test_distributed_cpu import run_world

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_partitioned_single_rank_matches_graph_path():
    if _gpus() < 1:
        pytest.skip("no CUDA device")
    import torch
    import pygsp_b200 as gsp
    from pygsp_b200 import distributed as gd
    from pygsp_b200.filters import approximations as apx
    G = gsp.graphs.Sensor(20000, k=8, seed=3, order="morton")
    G.estimate_lmax()
    L = G.L.to_scipy()
    plan = gd.HaloPlan(L, gd.even_bounds(G.N, 1), 0)
    op = gd.PartitionedCheby(plan)
    c = np.random.default_rng(0).standard_normal((2, 16)) / np.arange(1, 17)
    x = torch.randn(G.N, 64, device="cuda")
    a = op.cheby_op(G.lmax, c, x)
    b = apx.cheby_op_device(G.L, G.lmax, c, x)
    assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["0", "1", "p2p", "p2p_unfused"])
def test_partitioned_two_ranks(mode):
    """NCCL all-to-all-v (unsplit / split + overlapped) and the NVLink peer-store exchange."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_world("nccl", 2, n=60000, nsig=64, nscales=2, order=20, overlap=mode)


def test_partitioned_four_ranks_p2p():
    if _gpus() < 4:
        pytest.skip("needs 4 GPUs")
    run_world("nccl", 4, n=120000, nsig=64, nscales=1, order=16, overlap="p2p")


@pytest.mark.parametrize("world,dim,n_per,k", [(2, 3, 30000, 16), (2, 2, 20000, 10)])
def test_knn_slabs_two_ranks(world, dim, n_per, k):
    """BASELINE configs[4] shape at test size: per-rank device k-NN generator, device halo plan,
    fused peer-store exchange, against the float64 oracle on the assembled graph."""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    run_world("nccl", world, n=n_per, nsig=32, nscales=dim, order=20, overlap=k,
              worker="dist_worker_slabs.py")


def test_knn_slabs_four_ranks():
    if _gpus() < 4:
        pytest.skip("needs 4 GPUs")
    run_world("nccl", 4, n=40000, nsig=64, nscales=3, order=16, overlap=16,
              worker="dist_worker_slabs.py")


def test_knn_slabs_device_equals_host_backend():
    """One process, the ranks one after the other: the device generator (grid-hash k-NN,
    device symmetrisation, Laplacian rows assembled in HBM) gives the rows of the host one."""
    if _gpus() < 1:
        pytest.skip("no CUDA device")
    import torch
    from pygsp_b200.graphs.generators import (KnnSlabs, laplacian_rows, morton_order,
                                              morton_order_device)
    pts = np.random.default_rng(2).uniform(size=(5000, 3))
    np.testing.assert_array_equal(morton_order_device(torch.from_numpy(pts).cuda()).cpu().numpy(),
                                  morton_order(pts))
    for dim, P, n_per, k in ((3, 3, 20000, 16), (2, 2, 15000, 10), (3, 1, 5000, 7)):
        dev = [KnnSlabs(r, P, n_per, dim=dim, k=k, seed=4) for r in range(P)]
        host = [KnnSlabs(r, P, n_per, dim=dim, k=k, seed=4, backend="host") for r in range(P)]
        td, th = [g.distance_sum() for g in dev], [g.distance_sum() for g in host]
        sigma = sum(t[0] for t in th) / sum(t[1] for t in th)
        assert abs(sum(t[0] for t in td) / sum(t[1] for t in td) - sigma) <= 1e-12 * sigma
        for r in range(P):
            np.testing.assert_array_equal(dev[r].coords.cpu().numpy(), host[r].coords)
            ref, dw_ref = laplacian_rows(host[r].adjacency_rows(sigma), r * n_per)
            for dtype, tol in ((np.float64, 1e-12), (np.float32, 2e-6)):
                ptr, idx, val, dw = dev[r].laplacian_rows_device(sigma, dtype)
                np.testing.assert_array_equal(ptr.cpu().numpy(), ref.indptr)
                np.testing.assert_array_equal(idx.cpu().numpy(), ref.indices)
                np.testing.assert_allclose(val.cpu().numpy(), ref.data, rtol=tol, atol=tol * 1e-2)
                np.testing.assert_allclose(dw.cpu().numpy(), dw_ref, rtol=tol)
