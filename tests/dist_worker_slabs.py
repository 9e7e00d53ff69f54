"""Worker of the multi-GPU test of the per-rank graph generator (BASELINE configs[4] shape).

usage: dist_worker_slabs.py nccl <world> <rank> <port> <n_per> <nsig> <dim> <order> <k>
Every rank builds ITS slab of one k-NN graph on the GPU (graphs.KnnSlabs), plans its halo on
the device (HaloPlan.from_device) and filters; the result is compared with the float64 oracle
on the assembled global graph (host restatement of the same generator).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    backend, world, rank, port, n_per, nsig, dim, order, k = sys.argv[1:10]
    world, rank, n_per, nsig, dim, order, k = map(int, (world, rank, n_per, nsig, dim, order, k))
    import torch
    import torch.distributed as dist
    from scipy import sparse
    from oracle import pygsp_oracle as orc
    from pygsp_b200 import distributed as gd
    from pygsp_b200.graphs.generators import KnnSlabs, laplacian_rows

    torch.cuda.set_device(rank)
    dist.init_process_group(backend, init_method="tcp://127.0.0.1:%s" % port, world_size=world,
                            rank=rank)
    gen = KnnSlabs(rank, world, n_per, dim=dim, k=k, seed=11)
    tot = torch.tensor(gen.distance_sum(), dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    sigma = float(tot[0] / tot[1])
    ptr, idx, val, dw = gen.laplacian_rows_device(sigma)
    # the host restatement of the same generator, every slab: the global graph
    hosts = [KnnSlabs(r, world, n_per, dim=dim, k=k, seed=11, backend="host") for r in range(world)]
    hs = [h.distance_sum() for h in hosts]
    assert abs(sum(t[0] for t in hs) / sum(t[1] for t in hs) - sigma) <= 1e-12 * sigma
    np.testing.assert_allclose(gen.coords.cpu().numpy(), hosts[rank].coords, rtol=0, atol=0)
    W = sparse.vstack([h.adjacency_rows(sigma) for h in hosts]).tocsr()
    L = orc.laplacian(W)
    n = world * n_per
    mine = L[rank * n_per:(rank + 1) * n_per]
    np.testing.assert_array_equal(ptr.cpu().numpy(), mine.indptr)
    np.testing.assert_array_equal(idx.cpu().numpy(), mine.indices)
    np.testing.assert_allclose(val.cpu().numpy(), mine.data, rtol=2e-6, atol=1e-7)
    ref_rows, _ = laplacian_rows(hosts[rank].adjacency_rows(sigma), rank * n_per)
    np.testing.assert_array_equal(ref_rows.indices, mine.indices)

    bounds = gd.even_bounds(n, world)
    plan = gd.HaloPlan.from_device(ptr, idx, val, bounds, rank)
    host_plan = gd.HaloPlan(sparse.csr_matrix((val.cpu().numpy(), idx.cpu().numpy(),
                                               ptr.cpu().numpy()), shape=(n_per, n)), bounds, rank)
    for name in ("halo_ids", "perm", "send_idx", "send_counts", "recv_counts"):
        np.testing.assert_array_equal(getattr(plan, name), getattr(host_plan, name), err_msg=name)
    np.testing.assert_array_equal(plan.indices.cpu().numpy(), host_plan.indices)
    op = gd.PartitionedCheby(plan, dtype=torch.float32, exchange="p2p")
    lmax = op.estimate_lmax()
    lam = orc.lambda_max_exact(L)
    assert lam * (1 - 2e-4) <= lmax / 1.01 <= lam * (1 + 1e-5), (lmax, lam)
    rng = np.random.default_rng(5)
    c = rng.standard_normal((1, order + 1)) / np.arange(1, order + 2) ** 2
    x = rng.standard_normal((n, nsig))
    ref = orc.cheby_op(L.astype(np.float64), lmax, c, x).reshape(1, n, nsig)
    lo, hi = rank * n_per, (rank + 1) * n_per
    xl = torch.from_numpy(x[lo:hi]).to("cuda", torch.float32)
    errs = []
    for clenshaw in (True, False):
        r = op.cheby_op(lmax, c, xl, clenshaw=clenshaw).double().cpu().numpy()
        errs.append(float(np.abs(r - ref[:, lo:hi]).max() / np.abs(ref).max()))
    assert max(errs) <= 1e-5, errs
    # the packed NCCL exchange gives the bits of the peer-store exchange (same kernels, same order)
    op2 = gd.PartitionedCheby(plan, dtype=torch.float32, exchange="nccl")
    a = op.cheby_op(lmax, c, xl, clenshaw=False)
    b = op2.cheby_op(lmax, c, xl, clenshaw=False)
    assert torch.equal(a, b)
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok err=%.2e halo=%d boundary=%d/%d" % (rank, max(errs), plan.n_halo,
                                                        plan.n_true_boundary, plan.n_local))


if __name__ == "__main__":
    main()
