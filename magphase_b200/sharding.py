"""Utterance sharding across the GPUs of one box.

Utterances are independent units (the reference forks one process per utterance, src/libutils.py:32-63), so
the only collective on the path is the distribution of the WORK LIST: rank 0 enumerates the utterances with
their sizes, assigns them with the longest-processing-time greedy rule and broadcasts the assignment
(NCCL on GPUs, gloo in the CPU tests); every rank then loads / generates only its own shard.  Results stay
with their GPU; the throughput counters are all-reduced (MAX time, SUM units) at the end.
"""
import numpy as np


def lpt_assign(sizes, n_bins):
    """Longest-processing-time greedy assignment: returns owner[i] in [0, n_bins) for every utterance."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind='stable')
    load = np.zeros(n_bins, dtype=np.int64)
    owner = np.zeros(sizes.size, dtype=np.int64)
    for i in order:
        b = int(np.argmin(load))
        owner[i] = b
        load[b] += sizes[i]
    return owner


def scatter_work_list(sizes, device=None):
    """Collective.  `sizes` (per-utterance cost, e.g. samples) is only read on rank 0; every rank gets back
    (my_ids, owner) -- the indices of its own utterances and the full assignment vector."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        owner = lpt_assign(sizes, 1)
        return np.arange(owner.size), owner
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device if device is not None else torch.device('cpu')
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        n[0] = len(sizes)
    dist.broadcast(n, src=0)
    owner_t = torch.zeros(int(n.item()), dtype=torch.int64, device=dev)
    if rank == 0:
        owner_t = torch.from_numpy(lpt_assign(sizes, world)).to(dev)
    dist.broadcast(owner_t, src=0)
    owner = owner_t.cpu().numpy()
    return np.nonzero(owner == rank)[0], owner


def reduce_counters(elapsed, units, device=None):
    """Collective: (max over ranks of every entry of `elapsed`, sum over ranks of every entry of `units`)."""
    import torch
    import torch.distributed as dist
    e = np.asarray(elapsed, dtype=np.float64)
    u = np.asarray(units, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        return e, u
    dev = device if device is not None else torch.device('cpu')
    te, tu = torch.from_numpy(e.copy()).to(dev), torch.from_numpy(u.copy()).to(dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dist.all_reduce(tu, op=dist.ReduceOp.SUM)
    return te.cpu().numpy(), tu.cpu().numpy()
