"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo processes each own a contiguous
shard (jubjub_b200.shard_range), compute it (with the oracle standing in for the device, since
there is no GPU here), all-gather, and the gathered buffer must equal the single-process result
in index order -- the property jj_scalar_mul_sharded's in-place ncclAllGather relies on."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jubjub_b200.sharding import equal_shards, gather_offsets, shard_range
    from oracle import binding as ob
    from oracle import model as M

    per = equal_shards(n, world)
    lo, hi = shard_range(n, rank, world)
    assert hi - lo == per and gather_offsets(n, world)[rank] == lo
    g = ob.affine_to_extended(ob.generator())
    t = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 3, per, first=lo))  # inputs by global index
    k = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 2, per, first=lo))
    pts = ob.scalar_mul(np.repeat(g, per, axis=0), t, 1)
    mine = torch.from_numpy(ob.scalar_mul(pts, k, 1).view(np.int64))
    out = torch.empty((n, 20), dtype=torch.int64)
    dist.all_gather_into_tensor(out, mine)  # rank order == index order
    if rank == 0:
        ret["out"] = out.numpy().view(np.uint64).copy()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gather_order_world2(oracle):
    from oracle import model as M

    n, world = 24, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 3, n))
    k = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 2, n))
    want = oracle.scalar_mul(oracle.scalar_mul(np.repeat(g, n, axis=0), t), k)
    assert (ret["out"] == want).all()


def test_shard_range_properties():
    from jubjub_b200.sharding import equal_shards, shard_range
    import pytest

    for n in (0, 1, 7, 16, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert lo == prev and hi >= lo and hi - lo in (n // world, n // world + 1)
                prev = hi
            assert prev == n
    with pytest.raises(ValueError):
        equal_shards(10, 4)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker_ragged(rank, world, port, n, ret):
    """The ragged gather (one broadcast per rank's block, jj_scalar_mul_sharded_n's grouped ncclBroadcast) and the sharded
    sum (local sums, all-gather of the partial sums, sum in rank order: jj_point_sum_sharded), mirrored with gloo."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jubjub_b200.sharding import shard_range
    from oracle import binding as ob
    from oracle import model as M

    lo, hi = shard_range(n, rank, world)
    cnt = hi - lo
    g = ob.affine_to_extended(ob.generator())
    t = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 3, cnt, first=lo))
    k = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 2, cnt, first=lo))
    mine = ob.scalar_mul(ob.scalar_mul(np.repeat(g, cnt, axis=0), t, 1), k, 1)
    out = torch.zeros((n, 20), dtype=torch.int64)
    out[lo:hi] = torch.from_numpy(mine.view(np.int64))
    for r in range(world):  # every rank's block travels from its owner, in place
        a, b = shard_range(n, r, world)
        if b > a:
            blk = out[a:b].contiguous()
            dist.broadcast(blk, src=r)
            out[a:b] = blk
    # sharded sum: local sum -> all-gather of one point per rank -> sum in rank order
    def fold(p):
        acc = ob.identity()
        for i in range(len(p)):
            acc = ob.ext_add(acc, np.ascontiguousarray(p[i:i + 1]))
        return acc
    partial = torch.from_numpy(fold(mine).view(np.int64))
    partials = torch.empty((world, 20), dtype=torch.int64)
    dist.all_gather_into_tensor(partials, partial)
    total = fold(partials.numpy().view(np.uint64))
    ret[rank] = (out.numpy().view(np.uint64).copy(), ob.batch_normalize(total))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_gather_and_sharded_sum_world2(oracle):
    from oracle import model as M

    n, world = 25, 2  # blocks of 12 and 13
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_ragged, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 3, n))
    k = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 2, n))
    want = oracle.scalar_mul(oracle.scalar_mul(np.repeat(g, n, axis=0), t), k)
    acc = oracle.identity()
    for i in range(n):
        acc = oracle.ext_add(acc, np.ascontiguousarray(want[i:i + 1]))
    for r in range(world):
        out, total = ret[r]
        assert (out == want).all(), r                                  # every rank holds the whole batch in index order
        assert (total == oracle.batch_normalize(acc)).all(), r        # and the same sum, equal to the single-process fold
