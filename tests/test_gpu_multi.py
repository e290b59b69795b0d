"""Multi-GPU equality: the same work list at world size 1 and 2 gives byte-identical outputs.

The path has no data-plane collective: rank 0 LPT-assigns the utterances, NCCL broadcasts the assignment, every rank runs
its shard through the host API on its own GPU.  Sharding must therefore change nothing but where an utterance is computed.
Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_UTT = 10


def _utts():
    from magphase_b200.synth import synth_utterance
    return [synth_utterance(100 + u, fs=48000, dur_s=0.4 + 0.15 * (u % 4)) for u in range(N_UTT)]


def _run_shard(ids, device):
    """Analysis + synthesis of the utterances `ids` on `device`; the noise of utterance u comes from its own seeded stream,
    so it does not depend on what else is in the batch."""
    os.environ['MPB_DEVICE'] = str(device)
    import magphase_b200.magphase as mp
    utts = _utts()
    sub = [utts[i] for i in ids]
    feats = mp.analysis_compressed_batch([u[0] for u in sub], 48000, [u[1] for u in sub], [u[2] for u in sub], mag_dim=60, phase_dim=45)
    arrs, ns_len = mp.compressed_synthesis_geometry([f[3] for f in feats], [f[0].shape[0] for f in feats], 48000, 4096)
    noise = [np.random.RandomState(900 + int(i)).uniform(-1, 1, n) for i, n in zip(ids, ns_len)]
    ys = mp.synthesis_from_compressed_batch([f[:4] for f in feats], 48000, b_out_hpf=False, l_noise=noise)
    loss = mp.analysis_lossless_batch([u[0] for u in sub], 48000, [u[1] for u in sub], [u[2] for u in sub])
    return {int(i): [np.array(a) for a in f[:5]] + [np.array(y)] + [np.array(l[0])] for i, f, y, l in zip(ids, feats, ys, loss)}


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from magphase_b200.sharding import scatter_work_list
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    sizes = np.array([u[0].size for u in _utts()]) if rank == 0 else None
    my_ids, owner = scatter_work_list(sizes, device=torch.device('cuda', rank))
    out = _run_shard(my_ids, rank)
    q.put((rank, owner.tolist(), out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_outputs_equal_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as tmp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    merged = {}
    for rank, owner, out in res:
        assert set(out) == {i for i, o in enumerate(owner) if o == rank}
        merged.update(out)
    assert sorted(merged) == list(range(N_UTT))
    single = _run_shard(list(range(N_UTT)), 0)
    for i in range(N_UTT):
        for a, b in zip(merged[i], single[i]):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), i
