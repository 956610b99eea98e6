"""World sharding + the one all-reduce of batched system identification, 2 gloo ranks on CPU.

The stepping kernels need a GPU; what is covered here is the host logic every rank runs around them: which worlds a rank
owns, that per-world gradients stay local and that [loss, shared-parameter gradients] are summed with one collective.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsdfsim_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _surrogate(mass, fric, shared):
    """Differentiable stand-in for a rollout loss: per-world parameters + one parameter shared by all worlds."""
    return ((mass * shared[0] - fric) ** 2 + shared[1] * mass).sum()


def _worker(rank, size, port, n_worlds, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(size),
                      LOCAL_RANK=str(rank))
    r, s = D.init(device='cpu')
    assert (r, s) == (rank, size)
    g = torch.Generator().manual_seed(0)
    mass = torch.rand(n_worlds, generator=g, dtype=torch.float64)
    fric = torch.rand(n_worlds, generator=g, dtype=torch.float64)
    m = D.shard(mass, rank, size).clone().requires_grad_(True)
    f = D.shard(fric, rank, size).clone().requires_grad_(True)
    shared = torch.tensor([1.5, -0.3], dtype=torch.float64, requires_grad=True)
    loss = _surrogate(m, f, shared)
    loss.backward()
    tot, (gs,) = D.reduce_loss_and_shared_grads(loss.detach(), [shared.grad])
    t = D.max_over_ranks([float(rank + 1)], 'cpu')
    out[rank] = (float(tot), gs.tolist(), m.grad.tolist(), D.shard_range(n_worlds, rank, size), t[0])
    dist.destroy_process_group()


@pytest.mark.parametrize('n_worlds', [8, 11])
def test_two_ranks_match_single_process(n_worlds):
    size, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(size, port, n_worlds, out), nprocs=size, join=True)
    g = torch.Generator().manual_seed(0)
    mass = torch.rand(n_worlds, generator=g, dtype=torch.float64).requires_grad_(True)
    fric = torch.rand(n_worlds, generator=g, dtype=torch.float64)
    shared = torch.tensor([1.5, -0.3], dtype=torch.float64, requires_grad=True)
    loss = _surrogate(mass, fric, shared)
    loss.backward()
    covered = []
    for rank in range(size):
        tot, gs, gm, (lo, hi), tmax = out[rank]
        assert abs(tot - float(loss)) < 1e-12
        torch.testing.assert_close(torch.tensor(gs, dtype=torch.float64), shared.grad, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(torch.tensor(gm, dtype=torch.float64), mass.grad[lo:hi], rtol=1e-12, atol=1e-12)
        covered += list(range(lo, hi))
        assert tmax == 2.0                      # max over ranks
    assert covered == list(range(n_worlds))     # every world owned exactly once, in order


def test_shard_ranges_partition():
    for n in (1, 7, 4096, 65536):
        for size in (1, 2, 4, 8):
            edges = [D.shard_range(n, r, size) for r in range(size)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(size - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1


def test_single_process_is_identity():
    loss, (g,) = D.reduce_loss_and_shared_grads(torch.tensor(2.0), [torch.ones(3)])
    assert float(loss) == 2.0 and g.tolist() == [1, 1, 1]
