"""jj_scalar_mul_sharded / jj_scalar_mul_sharded_n on 2 GPUs: every rank ends with all ranks' results in index order,
equal to the oracle -- NCCL gather and the fused (NVLink P2P store) gather, ExtendedPoint and 32-byte outputs, equal
and ragged blocks, device and host inputs -- plus the write-after-read ordering of the fused path.  Skipped on a
single-GPU box: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` (scripts/gpu_multi.sh
saves the log under gpurun_out/)."""
import os
import sys

import numpy as np
import pytest

from oracle import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
WORLD = 2
N_EQ = 3000            # units per rank, equal blocks
N_RAGGED = 2 * 3000 + 1  # whole batch, ragged: blocks of 3000 and 3001
WAR_ROUNDS = 6


def _inputs(ob, first, count, kseed):
    g = ob.affine_to_extended(ob.generator())
    t = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 3, count, first=first))
    k = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, kseed, count, first=first))
    return ob.scalar_mul(np.repeat(g, count, axis=0), t, 2), k


def _worker(rank, world, uid_q, ret, handles, barrier):
    sys.path.insert(0, ROOT)
    import jubjub_b200 as jj
    from jubjub_b200 import shard_range
    from oracle import binding as ob

    eng = jj.Engine(rank)
    if rank == 0:
        uid = eng.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=60)
    eng.comm_init(world, rank, uid)
    fmt = {"extended": (20, np.uint64), "bytes": (32, np.uint8), "affine": (8, np.uint64)}

    def run(tag, n_total, output, host_in=False, want_local=False):
        lo, hi = shard_range(n_total, rank, world)
        pts, k = _inputs(ob, lo, hi - lo, M.SEED0 + 2)
        w, dt = fmt[output]
        out_all = buffers.get((n_total, output)) or eng.empty((n_total, w), dt)
        local = np.zeros((hi - lo, w), dt) if want_local else None
        a, b = (pts, k) if host_in else (eng.to_device(pts), eng.to_device(k))
        eng.scalar_mul_sharded_vartime(a, b, out_all, output=output, n_total=n_total, out_local_host=local)
        ret[(rank, tag)] = out_all.download()
        if want_local:
            ret[(rank, tag + "/local")] = local

    buffers = {}
    # ---- NCCL gather (ncclAllGather for equal blocks, grouped ncclBroadcast for ragged ones)
    run("nccl/ext", WORLD * N_EQ, "extended")
    run("nccl/bytes", WORLD * N_EQ, "bytes")
    run("nccl/ragged/ext", N_RAGGED, "extended")
    run("nccl/ragged/bytes", N_RAGGED, "bytes")
    run("nccl/host/ext", N_RAGGED, "extended", host_in=True, want_local=True)
    run("nccl/host/bytes", WORLD * N_EQ, "bytes", host_in=True, want_local=True)
    # the legacy equal-block entry point
    pts, k = _inputs(ob, rank * N_EQ, N_EQ, M.SEED0 + 2)
    o = eng.empty((WORLD * N_EQ, 20))
    eng.scalar_mul_sharded_vartime(eng.to_device(pts), eng.to_device(k), o)
    ret[(rank, "nccl/legacy")] = o.download()

    # ---- sum over the whole (ragged) batch: local sums + all-gather of the partial sums
    lo, hi = shard_range(N_RAGGED, rank, world)
    pts, k = _inputs(ob, lo, hi - lo, M.SEED0 + 2)
    prod = eng.scalar_mul_vartime(eng.to_device(pts), eng.to_device(k))
    ret[(rank, "sum")] = eng.point_sum_sharded(prod, output="bytes").download()

    # ---- fused gather: P2P stores into every rank's buffer from the kernel, no ncclAllGather
    def register(n_total, output):
        w, dt = fmt[output]
        out_all = eng.empty((n_total, w), dt)
        handles[(rank, n_total, output)] = eng.ipc_export(out_all)
        barrier.wait()
        ptrs = [out_all.ptr if r == rank else eng.ipc_open(handles[(r, n_total, output)]) for r in range(world)]
        eng.set_peer_outputs(ptrs)
        buffers[(n_total, output)] = out_all
        return ptrs

    def unregister(ptrs, key):
        eng.sync()
        barrier.wait()  # nobody frees a buffer a peer may still be writing or reading
        eng.set_peer_outputs(None)
        for r in range(world):
            if r != rank:
                eng.ipc_close(ptrs[r])
        buffers.pop(key)
        barrier.wait()

    for n_total, output, tag in ((WORLD * N_EQ, "extended", "fused/ext"), (N_RAGGED, "extended", "fused/ragged/ext"),
                                 (WORLD * N_EQ, "bytes", "fused/bytes"), (N_RAGGED, "bytes", "fused/ragged/bytes"),
                                 (N_RAGGED, "affine", "fused/ragged/affine")):
        ptrs = register(n_total, output)
        run(tag, n_total, output)
        if output == "extended":
            run(tag + "/host", n_total, output, host_in=True, want_local=True)
        unregister(ptrs, (n_total, output))

    # ---- write-after-read ordering of the fused path.  Every round multiplies by another scalar stream and is followed
    # by a consumer of out_all on the context's stream (batch_normalize into a per-round buffer).  Rank 1 is slowed
    # down BEFORE its consumer, so without the leading rendezvous rank 0 would run a round ahead and overwrite rank 1's
    # gathered buffer before rank 1 has read it.  Everything is enqueued asynchronously; checked after one sync.
    n_total = WORLD * N_EQ
    ptrs = register(n_total, "extended")
    out_all = buffers[(n_total, "extended")]
    pts, _ = _inputs(ob, rank * N_EQ, N_EQ, M.SEED0 + 2)
    dp = eng.to_device(pts)
    dks = [eng.to_device(_inputs(ob, rank * N_EQ, N_EQ, 7000 + i)[1]) for i in range(WAR_ROUNDS)]
    consumed = [eng.empty((n_total, 8)) for _ in range(WAR_ROUNDS)]
    big = 1 << 19
    slow_p = eng.scalar_mul_fixed_vartime(ob.generator(), eng.fe_to_bytes("fr", eng.fe_stream("fr", 1, big, device=True)))
    slow_k = eng.fe_to_bytes("fr", eng.fe_stream("fr", 2, big, device=True))
    slow_o = eng.empty((big, 20))
    eng.sync()
    barrier.wait()
    for i in range(WAR_ROUNDS):
        eng.scalar_mul_sharded_vartime(dp, dks[i], out_all, async_=True)
        if rank == 1:  # ~17 ms of unrelated work between the gather and its consumer
            eng.scalar_mul_vartime(slow_p, slow_k, out=slow_o, flags=jj.JJ_ASYNC)
        # the consumer: ordered on the context's stream, returns immediately (JJ_ASYNC)
        eng._call("jj_batch_normalize", [(out_all, 20, np.uint64)], 8, out=consumed[i], flags=jj.JJ_ASYNC)
    eng.sync()
    for i in range(WAR_ROUNDS):
        ret[(rank, f"war/{i}")] = consumed[i].download()
    unregister(ptrs, (n_total, "extended"))
    eng.close()


def test_sharded_all_gather_2gpu(oracle):
    import torch
    import torch.multiprocessing as mp

    from jubjub_b200 import shard_range

    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret, q, handles, barrier = mgr.dict(), mgr.Queue(), mgr.dict(), mgr.Barrier(WORLD)
    procs = [ctx.Process(target=_worker, args=(r, WORLD, q, ret, handles, barrier)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0

    def expected(n_total, kseed=M.SEED0 + 2):
        pts, k = _inputs(oracle, 0, n_total, kseed)
        aff = oracle.batch_normalize(oracle.scalar_mul(pts, k))
        return aff, oracle.affine_to_bytes(aff)

    want = {n: expected(n) for n in (WORLD * N_EQ, N_RAGGED)}
    checked = 0
    for r in range(WORLD):
        for tag, n_total, output in (("nccl/ext", WORLD * N_EQ, "extended"), ("nccl/bytes", WORLD * N_EQ, "bytes"),
                                     ("nccl/ragged/ext", N_RAGGED, "extended"), ("nccl/ragged/bytes", N_RAGGED, "bytes"),
                                     ("nccl/host/ext", N_RAGGED, "extended"), ("nccl/host/bytes", WORLD * N_EQ, "bytes"),
                                     ("nccl/legacy", WORLD * N_EQ, "extended"),
                                     ("fused/ext", WORLD * N_EQ, "extended"), ("fused/ragged/ext", N_RAGGED, "extended"),
                                     ("fused/ext/host", WORLD * N_EQ, "extended"), ("fused/ragged/ext/host", N_RAGGED, "extended"),
                                     ("fused/bytes", WORLD * N_EQ, "bytes"), ("fused/ragged/bytes", N_RAGGED, "bytes"),
                                     ("fused/ragged/affine", N_RAGGED, "affine")):
            got = ret[(r, tag)]
            aff, enc = want[n_total]
            if output == "extended":
                assert (oracle.batch_normalize(got) == aff).all(), (r, tag)
            elif output == "affine":
                assert (got == aff).all(), (r, tag)
            else:
                assert (got == enc).all(), (r, tag)
            if (r, tag + "/local") in ret:  # the host copy of the rank's own block
                lo, hi = shard_range(n_total, r, WORLD)
                assert (ret[(r, tag + "/local")] == got[lo:hi]).all(), (r, tag)
            checked += 1
        # same kernel, same bits: the fused gather must reproduce the NCCL gather byte for byte
        assert (ret[(r, "fused/ext")] == ret[(r, "nccl/ext")]).all()
        assert (ret[(r, "fused/ragged/ext")] == ret[(r, "nccl/ragged/ext")]).all()
    assert checked == WORLD * 14
    # sum_i [k_i] P_i over the whole ragged batch, computed by both ranks: equal to the oracle's sum of its own results
    aff, _ = want[N_RAGGED]
    tot = oracle.affine_to_extended(aff)
    while len(tot) > 1:
        if len(tot) % 2:
            tot = np.concatenate([tot, oracle.identity()])
        h = len(tot) // 2
        tot = oracle.ext_add(np.ascontiguousarray(tot[:h]), np.ascontiguousarray(tot[h:]))
    want_sum = oracle.affine_to_bytes(oracle.batch_normalize(tot))
    for r in range(WORLD):
        assert (ret[(r, "sum")] == want_sum).all(), r
    # write-after-read: the consumer of round i must have seen round i's results for EVERY unit of BOTH blocks
    for i in range(WAR_ROUNDS):
        aff, _ = expected(WORLD * N_EQ, 7000 + i)
        for r in range(WORLD):
            assert (ret[(r, f"war/{i}")] == aff).all(), (i, r)
