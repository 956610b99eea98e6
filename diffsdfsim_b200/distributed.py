"""World sharding across GPUs: one process per GPU, contiguous blocks of worlds, no data-path collective.

The only exchange of batched system identification / trajectory fitting is one ``all_reduce(SUM)`` per optimisation
iteration of ``[loss, gradients of the parameters shared by all worlds]``; per-world parameters never leave their owner.
Backend: ``nccl`` on the GPU box (NVLink 5 / NVSwitch), ``gloo`` in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init(device=None, backend=None):
    """Initialise the default process group when launched under torchrun; returns (rank, world_size)."""
    rank, _, size = env_rank()
    if size > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if device is not None and torch.device(device).type == 'cuda' else 'gloo')
        kw = {'device_id': torch.device(device)} if backend == 'nccl' else {}
        dist.init_process_group(backend, **kw)
    return rank, size


def shard_range(n_worlds, rank, size):
    """Contiguous block [lo, hi) of the global world index owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_worlds, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, size):
    """Rows of a (n_worlds, ...) tensor owned by ``rank``."""
    lo, hi = shard_range(t.shape[0], rank, size)
    return t[lo:hi]


def reduce_loss_and_shared_grads(loss, shared_grads, group=None):
    """Sum ``loss`` (0-d) and every tensor of ``shared_grads`` over the ranks with ONE all-reduce of one flat buffer.

    Returns (loss, grads) with the same shapes.  A single process returns its inputs unchanged.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return loss, list(shared_grads)
    flat = torch.cat([loss.reshape(1)] + [g.reshape(-1) for g in shared_grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, o = [], 1
    for g in shared_grads:
        out.append(flat[o:o + g.numel()].reshape(g.shape))
        o += g.numel()
    return flat[0], out


def max_over_ranks(values, device, group=None):
    """Element-wise max of a list of floats over the ranks (timings are the max over ranks, never wall clock)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t]
