"""world_size-2 gloo test of the sharded path's host logic (CPU): label-range routing, the single
all-gather of per-shard top-k and its [G][nq][k] layout.  The per-shard searches and the final merge
are played by the oracle / numpy here (the CUDA kernels are covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from embeddinghub_b200 import sharded
    from oracle import oracle as orc

    n_total, d, nq, k = 3000, 16, 40, 7
    base = np.random.default_rng(1234).standard_normal((n_total, d)).astype(np.float32)
    q = np.random.default_rng(4321).standard_normal((nq, d)).astype(np.float32)
    labels = np.arange(n_total, dtype=np.uint64)
    lo, hi = sharded.shard_bounds(n_total, world)[rank:rank + 2]
    mv, ml = sharded.route_rows(base, labels, n_total, world, rank)     # inserts route by label, no collective
    assert ml.min() == lo and ml.max() == hi - 1 and len(ml) == hi - lo
    loc_i, loc_d = orc.bruteforce(mv, q, k, "l2")                        # stand-in for the shard's kernel
    loc_l = ml[loc_i.astype(np.int64)]
    gl, gd = sharded.gather_topk(torch.from_numpy(loc_l.view(np.int64)), torch.from_numpy(loc_d), world)
    assert gl.shape == (world, nq, k)
    # rank g's slice must hold labels of rank g's range
    for g in range(world):
        a, b = sharded.shard_bounds(n_total, world)[g:g + 2]
        x = gl[g].numpy().view(np.uint64)
        assert x.min() >= a and x.max() < b
    # merge (numpy stand-in for merge_topk_kernel) == exact search over the union
    fd = np.transpose(gd.numpy(), (1, 0, 2)).reshape(nq, -1)
    fl = np.transpose(gl.numpy().view(np.uint64), (1, 0, 2)).reshape(nq, -1)
    order = np.argsort(fd, axis=1, kind="stable")[:, :k]
    merged = np.take_along_axis(fl, order, 1)
    ex, _ = orc.bruteforce(base, q, k, "l2")
    assert np.array_equal(merged, ex)
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_sharded_exchange_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get() == "ok"


def test_owner_of_and_bounds():
    from embeddinghub_b200 import sharded

    assert sharded.shard_bounds(10, 3) == [0, 3, 6, 10]
    assert sharded.owner_of([0, 2, 3, 5, 6, 9], 10, 3).tolist() == [0, 0, 1, 1, 2, 2]
