"""world_size-2 gloo tests (CPU) of the N>1 host logic: flattened gradient all-reduce, batch sharding,
and that two ranks which average their gradients take identical optimiser steps."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from reconfigisp_b200 import dist as D
    r, w, _ = D.init_from_env('gloo')
    assert (r, w) == (rank, world) and D.is_dist()
    # the 41-tensor / 216-float payload of n_step=3, with a None entry like an unused parameter
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(2,), (4,)] + [(15,), (1,), (2,), (1,), (2,), (1,), (3,), (1,), (3,), (3,), (30,), (3,), (5,)] * 3
    grads = [torch.randn(s, generator=g) for s in shapes]
    grads.insert(3, None)
    avg = D.allreduce_mean_flat(grads)
    assert avg[3] is None and sum(t.numel() for t in avg if t is not None) == 216
    # reference: the mean over ranks, tensor by tensor
    for i, (t, a) in enumerate(zip(grads, avg)):
        if t is None:
            continue
        both = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(both, t)
        assert torch.allclose(a, sum(both) / world, atol=1e-7), i
    # identical SGD steps on both ranks after averaging
    p = torch.nn.Parameter(torch.ones(216))
    opt = torch.optim.SGD([p], lr=0.1, momentum=0.9)
    p.grad = torch.cat([t.reshape(-1) for t in avg if t is not None])
    opt.step()
    gathered = [torch.zeros(216) for _ in range(world)]
    dist.all_gather(gathered, p.detach())
    assert torch.equal(gathered[0], gathered[1])
    # in-place average of a contiguous buffer (the tuning step's exchange when the peer-memory path is unavailable), and the
    # peer-memory reducer declining -- on every rank alike -- anything but a single-node NCCL group
    flat = torch.arange(8, dtype=torch.float32) * (rank + 1)
    D.allreduce_mean_(flat)
    assert torch.allclose(flat, torch.arange(8, dtype=torch.float32) * 1.5)
    assert D.P2PAllReduce.create() is None
    sl = D.shard_batch(8, rank, world)
    assert (sl.start, sl.stop) == (4 * rank, 4 * rank + 4)
    out.put((rank, float(p.detach().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]


def test_single_process_is_a_no_op():
    from reconfigisp_b200 import dist as D
    t = [torch.ones(3), None]
    out = D.allreduce_mean_flat(t)
    assert out[0] is t[0] and out[1] is None and not D.is_dist()
    assert D.P2PAllReduce.create() is None


def test_p2p_buffer_layout_is_sized_by_the_library():
    """Host-only entry: slots[2][world][cap] floats + flags[2][world] + epoch + timeout counter."""
    from reconfigisp_b200 import _lib as L
    assert L.size('risp_p2p_buffer_bytes', 8, 1024) == 4 * 2 * 8 * 1024 + 4 * (2 * 8 + 2)
    assert L.size('risp_p2p_buffer_bytes', 9, 1024) == 0 and L.size('risp_p2p_buffer_bytes', 2, 0) == 0
