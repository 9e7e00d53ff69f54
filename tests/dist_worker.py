"""Worker of the multi-process tests (spawned by test_distributed_*.py).

usage: dist_worker.py <backend: gloo|nccl> <world> <rank> <port> <n> <nsig> <nscales> <order> <overlap>
Every rank builds the same seeded global graph, keeps its row block, runs the
partitioned cheby_op and compares with the single-process oracle on the full graph.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class NumpyBackend:
    """Stand-in for the CUDA step kernel in the CPU (gloo) tests: same contract,
    arithmetic by the oracle's CSR product.  Test infrastructure only."""
    has_streams = False

    def __init__(self):
        import torch
        self.device = torch.device("cpu")

    def tile_plan(self, op, nsig, nscales):
        return None

    def gather_rows(self, buf, idx, nsig):
        return buf.index_select(0, idx).contiguous()

    def spmm(self, op, x_ext):
        import torch
        from oracle import pygsp_oracle as orc
        return torch.from_numpy(orc.csr_spmm(op.indptr.numpy().astype(np.int64), op.indices.numpy(),
                                             op.data.numpy(), x_ext.numpy()))

    def step(self, op, first, x_cur, x_old, x_new, r, nsig, nscales, ck, c0, coef, plan, rows):
        from oracle import pygsp_oracle as orc
        rb, re = rows
        if re <= rb:
            return
        ptr = op.indptr.numpy().astype(np.int64)
        s, e = ptr[rb], ptr[re]
        acc = orc.csr_spmm(ptr[rb:re + 1] - s, op.indices.numpy()[s:e], op.data.numpy()[s:e],
                           x_cur.numpy())
        alpha, beta, gamma = coef
        xc = x_cur.numpy()[rb:re]
        new = alpha * acc + beta * xc
        if not first:
            new = new + gamma * x_old.numpy()[rb:re]
        x_new.numpy()[rb:re] = new
        for i in range(nscales):
            if first:
                r.numpy()[i, rb:re] = 0.5 * c0[i] * xc + ck[i] * new
            else:
                r.numpy()[i, rb:re] += ck[i] * new


def main():
    backend, world, rank, port, n, nsig, nscales, order, overlap = sys.argv[1:10]
    world, rank, n, nsig, nscales, order = map(int, (world, rank, n, nsig, nscales, order))
    import torch
    import torch.distributed as dist
    from scipy import sparse, spatial
    from oracle import pygsp_oracle as orc
    from pygsp_b200 import distributed as gd

    if backend == "nccl":
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, init_method="tcp://127.0.0.1:%s" % port, world_size=world,
                            rank=rank)
    rng = np.random.default_rng(123)
    pts = rng.uniform(size=(n, 2))
    pts = pts[np.argsort(pts[:, 0], kind="stable")]          # 1-D strips along x
    D, NN = spatial.cKDTree(pts).query(pts, k=7)
    W = sparse.csr_matrix((np.exp(-D[:, 1:].ravel() ** 2 / D[:, 1:].mean()),
                           (np.repeat(np.arange(n), 6), NN[:, 1:].ravel())), shape=(n, n))
    W = ((W + W.T) / 2).tocsr()
    L = orc.laplacian(W)
    lmax = orc.upper_bound(W)
    c = rng.standard_normal((nscales, order + 1)) / np.arange(1, order + 2) ** 2
    x = rng.standard_normal((n, nsig))
    ref = orc.cheby_op(L, lmax, c, x).reshape(nscales, n, nsig)

    bounds = gd.even_bounds(n, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    plan = gd.HaloPlan(L[lo:hi], bounds, rank)
    assert plan.n_boundary % 4 == 0 or plan.n_boundary == plan.n_local
    assert plan.send_idx.size == plan.send_counts.sum()
    assert plan.recv_counts[rank] == 0 and plan.send_counts[rank] == 0
    if plan.send_idx.size:
        assert plan.send_idx.max() < plan.n_boundary       # symmetric pattern: sent rows are boundary rows
    if world > 1:
        assert plan.n_halo > 0

    if backend == "nccl":
        dtype = torch.float32
        if overlap in ("p2p", "p2p_unfused"):
            op = gd.PartitionedCheby(plan, dtype=dtype, exchange="p2p")
            op.fuse_halo = overlap == "p2p"
        else:
            op = gd.PartitionedCheby(plan, dtype=dtype, overlap=(overlap == "1"), exchange="nccl")
        xl = torch.from_numpy(x[lo:hi]).to(device=op.device, dtype=dtype)
        tol = 1e-5
    else:
        dtype = torch.float64
        op = gd.PartitionedCheby(plan, dtype=dtype, backend=NumpyBackend())
        xl = torch.from_numpy(x[lo:hi].copy())
        tol = 1e-12
    # distributed lmax: brackets the true eigenvalue like the single-process estimate
    lam = orc.lambda_max_exact(L)
    est = op.estimate_lmax()
    assert lam * (1 - 2e-4) <= est / 1.01 <= lam * (1 + 1e-5), (est, lam)
    bound = op.estimate_lmax(method="bounds")               # the reference's four bounds, distributed
    assert bound >= lam and abs(bound - orc.upper_bound(W)) <= (1e-5 if backend == "nccl" else 1e-10) * bound
    assert op.estimate_lmax(method="bounds", lap_type="normalized") == 2
    r = op.cheby_op(lmax, c, xl, clenshaw=False).cpu().numpy().astype(np.float64)
    want = ref[:, lo:hi]
    err = np.abs(r - want).max() / np.abs(ref).max()
    assert err <= tol, (rank, err)

    if backend == "nccl":
        # bit-identical to the single-GPU engine on the full graph (same accumulation order)
        from pygsp_b200.filters import approximations as apx
        from pygsp_b200.graphs import DeviceCSR
        Ld = DeviceCSR.from_scipy(L, dtype, op.device)        # the same float32 values of L
        full = apx.cheby_op_device(Ld, lmax, c, torch.from_numpy(x).to(op.device, dtype))
        for _ in range(3):                                   # repeated calls reuse windows / flags
            mine = op.cheby_op(lmax, c, xl, clenshaw=False)
            assert torch.equal(mine, full[:, lo:hi]), float((mine - full[:, lo:hi]).abs().max())
        # single filter: the peer-memory path defaults to the Clenshaw form, bit-identical to the
        # single-GPU Clenshaw evaluation; the NCCL path keeps the forward recurrence
        xd = torch.from_numpy(x).to(op.device, dtype)
        c1 = c[:1]
        fwd1 = apx.cheby_op_device(Ld, lmax, c1, xd)
        cl1 = apx.cheby_clenshaw_device(Ld, lmax, c1, xd)[None]
        want1 = cl1 if overlap in ("p2p", "p2p_unfused") else fwd1
        for _ in range(2):
            mine1 = op.cheby_op(lmax, c1, xl)
            assert torch.equal(mine1, want1[:, lo:hi]), float((mine1 - want1[:, lo:hi]).abs().max())
        assert torch.equal(op.cheby_op(lmax, c1, xl, clenshaw=False), fwd1[:, lo:hi])
        e1 = float((cl1 - fwd1).abs().max() / fwd1.abs().max())
        assert e1 <= 2e-5, e1
        x2 = torch.from_numpy(x[:, :32].copy()).to(op.device, dtype)      # another signal width
        full2 = apx.cheby_op_device(Ld, lmax, c, x2)
        assert torch.equal(op.cheby_op(lmax, c, x2[lo:hi].contiguous(), clenshaw=False),
                           full2[:, lo:hi])
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok err=%.2e halo=%d boundary=%d/%d" % (rank, err, plan.n_halo, plan.n_true_boundary,
                                                        plan.n_local))


if __name__ == "__main__":
    main()
