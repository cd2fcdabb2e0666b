"""jj_scalar_mul_sharded on 2 GPUs: every rank ends with all ranks' results in index order, equal to
the oracle.  Skipped on a single-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`)."""
import os
import sys

import numpy as np
import pytest

from oracle import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, uid_q, n_local, ret, handles, barrier):
    sys.path.insert(0, ROOT)
    import jubjub_b200 as jj
    from oracle import binding as ob

    eng = jj.Engine(rank)
    if rank == 0:
        uid = eng.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=60)
    eng.comm_init(world, rank, uid)
    lo = rank * n_local
    g = ob.affine_to_extended(ob.generator())
    t = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 3, n_local, first=lo))
    k = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 2, n_local, first=lo))
    pts = ob.scalar_mul(np.repeat(g, n_local, axis=0), t, 2)
    for output, width, dtype in (("extended", 20, np.uint64), ("bytes", 32, np.uint8)):
        out_all = eng.empty((world * n_local, width), dtype)
        eng.scalar_mul_sharded(eng.to_device(pts), eng.to_device(k), out_all, output=output)
        ret[(rank, output)] = out_all.download()
    # fused path: P2P stores into every rank's buffer from the kernel epilogue, no ncclAllGather
    out_all = eng.empty((world * n_local, 20), np.uint64)
    handles[rank] = eng.ipc_export(out_all)
    barrier.wait()
    ptrs = [out_all.ptr if r == rank else eng.ipc_open(handles[r]) for r in range(world)]
    eng.set_peer_outputs(ptrs)
    eng.scalar_mul_sharded(eng.to_device(pts), eng.to_device(k), out_all, output="extended")
    ret[(rank, "fused")] = out_all.download()
    barrier.wait()  # nobody frees a buffer a peer may still be writing or reading
    eng.set_peer_outputs(None)
    for r in range(world):
        if r != rank:
            eng.ipc_close(ptrs[r])
    barrier.wait()
    eng.close()


def test_sharded_all_gather_2gpu(oracle):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n_local = 2, 3000
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret, q, handles, barrier = mgr.dict(), mgr.Queue(), mgr.dict(), mgr.Barrier(world)
    procs = [ctx.Process(target=_worker, args=(r, world, q, n_local, ret, handles, barrier)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    n = world * n_local
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 3, n))
    k = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 2, n))
    want_aff = oracle.batch_normalize(oracle.scalar_mul(oracle.scalar_mul(np.repeat(g, n, axis=0), t), k))
    for r in range(world):
        assert (oracle.batch_normalize(ret[(r, "extended")]) == want_aff).all()
        assert (ret[(r, "bytes")] == oracle.affine_to_bytes(want_aff)).all()
        assert (ret[(r, "fused")] == ret[(r, "extended")]).all()  # same kernel, same bits, gathered by P2P stores
