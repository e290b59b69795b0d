"""The multi-GPU host logic on CPU: world_size-2 gloo process group, work-list scatter + counter reduction."""
import os
import socket

import numpy as np
import torch.multiprocessing as tmp

from magphase_b200.sharding import lpt_assign


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from magphase_b200.sharding import reduce_counters, scatter_work_list
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    sizes = rng.integers(50000, 400000, 37) if rank == 0 else None      # only rank 0 knows the list
    my_ids, owner = scatter_work_list(sizes)
    frames = float(len(my_ids) * 10)
    t, u = reduce_counters([1.0 + rank, 5.0 - rank], [frames, 1.0])
    q.put((rank, my_ids.tolist(), owner.tolist(), t.tolist(), u.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_work_list_scatter_world2():
    world = 2
    port = _free_port()
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ids0, owner0, t0, u0), (r1, ids1, owner1, t1, u1) = res
    assert owner0 == owner1 and len(owner0) == 37
    assert sorted(ids0 + ids1) == list(range(37)) and not set(ids0) & set(ids1)      # a partition
    sizes = np.random.default_rng(0).integers(50000, 400000, 37)
    assert np.array_equal(np.array(owner0), lpt_assign(sizes, 2))                    # rank 0's LPT plan
    loads = [int(sizes[ids0].sum()), int(sizes[ids1].sum())]
    assert abs(loads[0] - loads[1]) <= sizes.max()
    assert t0 == t1 == [2.0, 5.0] and u0 == u1 == [370.0, 2.0]                       # MAX time, SUM units


def test_lpt_balance():
    sizes = np.full(1000, 240000)
    for g in (1, 2, 4, 8):
        owner = lpt_assign(sizes, g)
        counts = np.bincount(owner, minlength=g)
        assert counts.max() - counts.min() <= 1
    rng = np.random.default_rng(1)
    sizes = rng.integers(10000, 1000000, 12500)
    load = np.bincount(lpt_assign(sizes, 8), weights=sizes, minlength=8)
    assert (load.max() - load.min()) / load.mean() < 0.01      # < 1 % imbalance with LPT
