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
