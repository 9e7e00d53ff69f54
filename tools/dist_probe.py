"""Where does the partitioned step lose time?  (2 GPUs, under torchrun, weak-scaling workload)

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_probe.py

Builds rank q's 1e6-vertex strip exactly like `bench.py --scaling weak` and times, in one process:
  fused     : PartitionedCheby.cheby_op (peer-store exchange fused into the step kernel)
  plain     : the SINGLE-GPU engine on the rank's local matrix padded to a square (halo rows
              empty, no exchange at all) with ordinary torch buffers
  window    : the same, state buffers placed in the IPC-exported peer window
  forcehalo : plain + the halo-capable kernel instantiation (no neighbours)
One JSON line per rank.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    import pygsp_b200 as gsp
    from pygsp_b200 import distributed as gd
    from pygsp_b200.filters import approximations as apx
    from pygsp_b200.graphs import DeviceCSR

    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = dict(bench.WORKLOADS["config2"])
    op, lmax, nnz = bench.build_partitioned_strips(gsp, wl, rank, world, torch, dist)
    p = op.plan
    n, ext, nsig, order = p.n_local, p.n_local + p.n_halo, 64, 30

    class _G:
        pass
    g = _G(); g.lmax = lmax; g.N = n
    c = np.atleast_2d(gsp.filters.compute_cheby_coeff(gsp.filters.Heat(g, scale=50.0), m=order))
    x = torch.randn(n, nsig, device="cuda")
    out = {"rank": rank, "n_local": n, "n_halo": p.n_halo}

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    out["fused_ms"] = timed(lambda: op.cheby_op(lmax, c, x, local_order=True))
    out["fused_forward_ms"] = timed(lambda: op.cheby_op(lmax, c, x, local_order=True, clenshaw=False))
    # the local matrix as a square single-GPU operator: halo rows are empty rows
    ptr = torch.cat([op.indptr, torch.full((p.n_halo,), int(p.nnz), dtype=torch.int32, device="cuda")])
    Lext = DeviceCSR(ptr, op.indices, op.data, (ext, ext))
    xe = torch.randn(ext, nsig, device="cuda")
    o_t = torch.empty(ext, nsig, device="cuda")
    w_t = torch.empty(2, ext, nsig, device="cuda")
    out["plain_ms"] = timed(lambda: apx.cheby_clenshaw_device(Lext, lmax, c, xe, out=o_t, work=w_t))
    win = op._windows[nsig]
    if win.buf_bytes == ext * nsig * 4:
        w_w = gd._wrap(win.base + win.buf_bytes, (2, ext, nsig), torch.float32, op.device)
        win.bufs[0].copy_(xe)
        out["window_ms"] = timed(lambda: apx.cheby_clenshaw_device(Lext, lmax, c, win.bufs[0], out=o_t, work=w_w))
        out["window_src_only_ms"] = timed(lambda: apx.cheby_clenshaw_device(Lext, lmax, c, win.bufs[0], out=o_t, work=w_t))
    else:
        out["window_ms"] = "window buffers are padded (%d vs %d)" % (win.buf_bytes, ext * nsig * 4)
    os.environ["GSPB200_FORCE_HALO"] = "1"
    out["forcehalo_ms"] = timed(lambda: apx.cheby_clenshaw_device(Lext, lmax, c, xe, out=o_t, work=w_t))
    del os.environ["GSPB200_FORCE_HALO"]
    # natural (unpermuted) order for reference: the strip's rows as generated
    out["fused_again_ms"] = timed(lambda: op.cheby_op(lmax, c, x, local_order=True))
    print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
