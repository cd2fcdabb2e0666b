#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched Jubjub engine.

Metric (BASELINE.json): variable-base scalar-muls / second.  Workload:
  N = 1   BASELINE.json configs[2], the configuration the metric is quoted on: 2^20 variable-base
          `ExtendedPoint * Fr` scalar-muls on one GPU;
  N > 1   BASELINE.json configs[4]: 2^21 units per GPU -- 16 M over 8 GPUs -- with the all-gather of the output
          points inside every step (weak scaling: per-GPU work is fixed for all N > 1).
Points P_i = [t_i] G (full-order), scalars uniform in [0, r), both from the SplitMix64 streams of SURVEY.md
section 8d, generated per rank by index.  One step = one pass of the hot path over that batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--units-log2 L]

N > 1 is launched by torchrun (one rank per GPU); each rank owns a contiguous block and every step ends with all
ranks holding all results (jj_scalar_mul_sharded: all-gather fused into the kernel by NVLink P2P stores, or
ncclAllGather with JJ_GATHER=nccl).  Timing: CUDA events on the engine's stream, barrier + synchronize on both
sides, max over ranks.  After the timed loop EVERY rank checks its gathered buffer against the CPU oracle (a sample
of every rank's block, in index order) and against the other ranks' digests: `parity_check` in the JSON line.
Rank 0 prints ONE JSON line.

`--impl reference` times the reference *algorithm* (bitwise 252-step double-and-add ladder on 4 x u64 Montgomery
limbs) as restated in C in oracle/ -- the reference itself is Rust and cannot be built in this image -- on all host
cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED0 = 0x4A55424A55420001
METRIC = "variable_base_scalar_muls_per_sec"
UNIT = "scalar-muls/s"
# algorithmic work per unit: bytes = 160 B point + 32 B scalar in, 160 B point out;
# IMAD.WIDE.U32 = signed radix-16 window: 249 doublings (4S+3M; the top part of the scalar starts from a table entry, so the
# first pass doubles once), 7 table + 1 + ~58 digit additions (8M), 16 M for to_niels,
# S = 84 and M = 112 multiplier instructions: the minimum of 8x32-bit Montgomery with q's special low limbs.  (The
# shipped kernels issue S = 91, M = 119 -- one extra multiply per reduction row but the last replaces three ALU
# instructions -- so the fraction undercounts the pipe's real occupancy; `imads_issued_per_unit` is ncu's count.)
BYTES_PER_UNIT = 352
IMADS_PER_UNIT = 249 * (4 * 84 + 3 * 112) + (7 + 1 + 62 * 15 / 16) * 8 * 112 + 16 * 112
# from the committed ncu capture of the dominant kernel at 2^20 units (ncu --set full, one launch):
# dram__bytes_read.sum + dram__bytes_write.sum, and IMAD.WIDE thread instructions executed per unit (source page)
NCU_PROFILE = "profiles/r02u_ncu_scalar_mul_default_n1048576.csv"
NCU_DRAM_BYTES_PER_LAUNCH = 226.21e6 + 657.02e6
NCU_IMADS_ISSUED_PER_UNIT = 243955
GEN_RAW = np.array([[0xE4B3D35DF1A7ADFE, 0xCAF55D1B29BF81AF, 0x8B0F03DDD60A8187, 0x62EDCBB8BF3787C8, 0xB, 0, 0, 0]],
                   dtype=np.uint64)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=100):
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        # "under load" = the upper half of the samples (idle samples before/after the region drop out)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def generator_mont(eng):
    u = eng.fe_from_bytes("fq", GEN_RAW[:, :4].view(np.uint8).reshape(1, 32))[0]
    v = eng.fe_from_bytes("fq", GEN_RAW[:, 4:].view(np.uint8).reshape(1, 32))[0]
    return np.concatenate([u, v], axis=1)


def make_inputs(eng, n, first):
    """Device-resident shard: points [t_i]G (extended, 160 B) and canonical scalars (32 B)."""
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 3, n, first=first, device=True))
    pts = eng.scalar_mul_fixed_vartime(generator_mont(eng), t)
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, n, first=first, device=True))
    t.free()
    return pts, k


def pinned(eng, shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    eng._check(eng.lib.jj_host_alloc(eng.ctx, nbytes, C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def oracle_expected(first, count, threads):
    """Normalised affine results (count, 8) of units [first, first + count) of the workload, by the CPU oracle."""
    from oracle import binding as ob
    from oracle import model as M

    g = ob.affine_to_extended(ob.generator())
    t = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 3, count, first=first))
    k = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 2, count, first=first))
    pts = ob.scalar_mul(np.repeat(g, count, axis=0), t, threads)
    return ob.batch_normalize(ob.scalar_mul(pts, k, threads))


def parity_check(eng, out_all, n, world, rank, dist, threads, sample_per_block=1024, runs=8, fmt="extended"):
    """Every rank: (1) a digest of its whole gathered buffer, compared across ranks; (2) `sample_per_block` units of
    EVERY rank's block (runs of consecutive indices at fixed pseudo-random offsets) recomputed by the oracle from the
    input streams by global index -- a misplaced block or a stale buffer fails this, not only a wrong value."""
    from oracle import binding as ob

    host = out_all.download()  # (world * n, 20) uint64, or (world * n, 32) uint8 encodings
    words = host.reshape(-1).view(np.uint64)
    s1 = s2 = np.uint64(0)
    step = 1 << 24
    with np.errstate(over="ignore"):
        for lo in range(0, words.size, step):  # chunked: no multi-GB temporaries
            w = words[lo:lo + step]
            idx = np.arange(lo + 1, lo + 1 + w.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
            s1 = s1 + w.sum(dtype=np.uint64)
            s2 = s2 + (w * idx).sum(dtype=np.uint64)
    digest = [int(s1), int(s2)]
    run_len = sample_per_block // runs
    rng = np.random.RandomState(12345)
    bad, checked = 0, 0
    for r in range(world):
        for s in rng.randint(0, n - run_len, size=runs):
            first = r * n + int(s)
            want = oracle_expected(first, run_len, threads)
            if fmt == "bytes":
                got, want = host[first:first + run_len], ob.affine_to_bytes(want)
            else:
                got = ob.batch_normalize(host[first:first + run_len])
            bad += int((got != want).any(axis=1).sum())
            checked += run_len
    digests = [digest]
    if dist is not None:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    same = all(d == digests[0] for d in digests)
    res = {"rank": rank, "units_checked_vs_oracle": checked, "mismatches": bad, "digest": digest,
           "digest_equal_across_ranks": same}
    allres = [res]
    if dist is not None:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    return {"ok": all(a["mismatches"] == 0 and a["digest_equal_across_ranks"] for a in allres),
            "ranks_checked": len(allres), "units_checked_vs_oracle_per_rank": checked,
            "sample": f"{runs} runs of {run_len} consecutive units in every rank's block, on every rank, normalised "
                      "affine vs oracle/ (inputs regenerated from the streams by global index)",
            "mismatches": sum(a["mismatches"] for a in allres),
            "digest_equal_across_ranks": all(a["digest_equal_across_ranks"] for a in allres),
            "digest": "sum and index-weighted sum (mod 2^64) of all 64-bit words of the gathered buffer"}


def run_reference(args, rank, emit=print):
    """Reference arm: the reference algorithm (C restatement, oracle/) on all host cores."""
    if rank != 0:
        return
    from oracle import binding as ob
    from oracle import model as M

    cores = os.cpu_count() or 1
    logn = workload_log2(args)
    sample = int(os.environ.get("JJ_REF_SAMPLE", str(min(1 << logn, 4096 * cores))))
    g = ob.affine_to_extended(ob.generator())
    t = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 3, sample))
    pts = ob.scalar_mul(np.repeat(g, sample, axis=0), t, cores)
    k = ob.fe_to_bytes(ob.FR, ob.fe_stream(ob.FR, M.SEED0 + 2, sample))
    for _ in range(args.warmup):
        ob.time_scalar_mul(pts, k, cores, reps=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.time_scalar_mul(pts, k, cores, reps=1)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} of 2^{logn} units per step (first {sample} of the same streams), {cores} threads"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (4x64 Montgomery)",
        "data": "synthetic", "config": {"workload": workload_name(args.gpus, logn), "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm (bitwise ladder, src/lib.rs:356-379) as a C restatement; the Rust reference "
                "cannot be built in this image (no cargo/rustc; Fq lives in the un-vendored bls12_381 crate)",
    }))


_FULL_AFFINITY = set()


def bind_host_to_gpu(local):
    """N > 1: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, so that the pinned buffers it
    allocates next (first touch) are local to that GPU's PCIe root -- without it the ranks of the far socket stage every
    chunk across the inter-socket link.  Best effort: returns what was done for the JSON line.  JJ_NUMA_BIND=0 disables."""
    info = {"bound": False}
    if os.environ.get("JJ_NUMA_BIND", "1") == "0":
        info["why"] = "JJ_NUMA_BIND=0"
        return info
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info.update({"gpu": bdf, "numa_node": node, "local_cpus": len(cpus), "allowed_cpus": len(allowed), "usable": len(use)})
        if node < 0 or not use:
            info["why"] = "no NUMA information" if node < 0 else "none of the GPU's local CPUs is in this process's cpuset"
            return info
        if use != allowed:
            _FULL_AFFINITY.update(allowed)  # given back before the CPU-baseline leg
            os.sched_setaffinity(0, use)
        info["bound"] = True
    except Exception as e:  # noqa: BLE001 -- best effort, never fatal
        info["why"] = "%s: %s" % (type(e).__name__, e)
    return info


def workload_log2(args):
    if args.units_log2:
        return args.units_log2
    if os.environ.get("JJ_UNITS_LOG2"):
        return int(os.environ["JJ_UNITS_LOG2"])
    return 20 if args.gpus <= 1 else 21


def workload_name(gpus, logn):
    base = f"2^{logn} variable-base ExtendedPoint*Fr scalar-muls per GPU (P_i=[t_i]G, 252-bit scalars, SplitMix64 streams)"
    if gpus <= 1:
        return base + " -- BASELINE.json configs[2] (1M on 1 B200)"
    total = gpus << logn
    return base + (f"; {total} units over {gpus} GPUs, all-gather of the results every step -- BASELINE.json configs[4] "
                   f"(16M over 8 GPUs = 2^21 per GPU; the same per-GPU batch at every N > 1)")


def timed(eng, fn, reps):
    eng.sync()
    eng.timer_start()
    for _ in range(reps):
        fn()
    return eng.timer_stop() / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--units-log2", type=int, default=0, help="units per GPU (default: 20 at N=1, 21 at N>1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.gpus = max(args.gpus, world)
    # Keep stdout to the single JSON line: anything native libraries print to fd 1 while the bench runs
    # (NCCL's INFO log goes to stdout by default) is sent to stderr; the JSON goes to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        # written straight to the saved descriptor: fd 1 stays pointed at stderr, so whatever NCCL logs while the
        # process group is torn down cannot land after the JSON line
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, emit)
        return
    args.warmup = max(args.warmup, 3)

    import jubjub_b200 as jj

    dist = None
    numa = {"bound": False, "why": "single GPU"}
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        numa = bind_host_to_gpu(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = jj.Engine(local)
    logn = workload_log2(args)
    n = 1 << logn
    pts, k = make_inputs(eng, n, first=rank * n)
    # JJ_OUT=bytes (information only, N > 1): every rank gathers the 32-byte encodings of the results instead of the
    # 160-byte points -- the kernel's fused normalise epilogue + 32-byte P2P stores.  The headline stays ExtendedPoint.
    out_fmt = os.environ.get("JJ_OUT", "extended") if world > 1 else "extended"
    unit_out = 160 if out_fmt == "extended" else 32
    out_all = eng.empty((world * n, 20)) if out_fmt == "extended" else eng.empty((world * n, 32), np.uint8)
    gather = "none"
    if world > 1:
        ids = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(world, rank, ids[0])
        gather = os.environ.get("JJ_GATHER", "p2p")
        if gather == "p2p":
            # fused compute + all-gather: exchange CUDA IPC handles of every rank's gathered buffer; the
            # scalar-mul kernel then stores each result into all of them over NVLink (no ncclAllGather)
            handles = [None] * world
            dist.all_gather_object(handles, eng.ipc_export(out_all))
            try:
                ptrs, ok = [out_all.ptr if r == rank else eng.ipc_open(handles[r]) for r in range(world)], 1
            except jj.JubjubError:
                ptrs, ok = None, 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank must be able to map every peer
            if int(flag.item()) == 1:
                eng.set_peer_outputs(ptrs)
            else:
                gather = "nccl"

    def step():
        if world > 1:
            eng.scalar_mul_sharded_vartime(pts, k, out_all, output=out_fmt, async_=True)
        else:
            eng.scalar_mul_vartime(pts, k, out=out_all, flags=jj.JJ_ASYNC)

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count()
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - launches0
    if dist is not None:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * n * args.steps / (ms * 1e-3)

    # ---- parity of what was just timed: every rank checks the gathered buffer the last step left behind
    cores = os.cpu_count() or 1
    threads = max(1, cores // world)
    parity = parity_check(eng, out_all, n, world, rank, dist, threads, fmt=out_fmt)

    # ---- e2e: HOST (pinned) inputs through the C ABI, results on every rank, own block read back to the host;
    # H2D, kernels, gather and D2H all inside the timed region
    hp, hp_ptr = pinned(eng, (n, 20), np.uint64)
    hk, hk_ptr = pinned(eng, (n, 32), np.uint8)
    ho, ho_ptr = pinned(eng, (n, 20), np.uint64) if out_fmt == "extended" else pinned(eng, (n, 32), np.uint8)
    hp[:] = pts.download()
    hk[:] = k.download()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        if world > 1:
            eng.scalar_mul_sharded_vartime(hp, hk, out_all, output=out_fmt, out_local_host=ho)
        else:
            eng.scalar_mul_vartime(hp, hk, out=ho)

    e2e_step()  # warm-up (staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert ho[-1].any(), "e2e produced no output"
    e2e_same = bool((ho == out_all.download()[rank * n:(rank + 1) * n]).all())  # host copy == own block of the gather
    if dist is not None:
        tmax = torch.tensor([e2e_s, 0.0 if e2e_same else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s, e2e_same = float(tmax[0].item()), float(tmax[1].item()) == 0.0
    # the e2e steps recomputed the same batch: the buffer must still pass the digest comparison
    parity["e2e_host_copy_matches_gather"] = e2e_same
    parity["ok"] = bool(parity["ok"] and e2e_same)

    numa_all = [numa]
    if dist is not None:
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0 only from here: kernel-only duration of the dominant kernel (same launches, no gather), for the roofline
    eng.set_peer_outputs(None)
    if out_fmt != "extended":
        out_all = eng.empty((n, 20))
    kms = timed(eng, lambda: eng.scalar_mul_vartime(pts, k, out=out_all, flags=jj.JJ_ASYNC), args.steps)
    hbm_peak, peak_src = measured_peaks()
    achieved_gbs = BYTES_PER_UNIT * n / (kms * 1e-3) / 1e9
    peak_sampler = ClockSampler(local, period_ms=50).start()
    time.sleep(0.5)  # nvidia-smi needs a moment to start sampling
    imad_peak = max(eng.imad_peak() for _ in range(6))  # ~150 ms each: the clock samples below are taken under this load
    peak_clocks = peak_sampler.stop()
    achieved_imad = IMADS_PER_UNIT * n / (kms * 1e-3)

    # ---- BASELINE config 2: Fq mul / square / add at 2^20 (L2-resident) and 2^26 (HBM-sized)
    fq = {}
    for lg in (20, 26):
        m = 1 << lg
        a = eng.fe_stream("fq", SEED0, m, device=True)
        b = eng.fe_stream("fq", SEED0 + 1, m, device=True)
        o = eng.empty((m, 4))
        row = {}
        for name, fn, nbytes in (("mul", lambda: eng.fe_mul("fq", a, b, out=o, flags=jj.JJ_ASYNC), 96),
                                 ("square", lambda: eng.fe_square("fq", a, out=o, flags=jj.JJ_ASYNC), 64),
                                 ("add", lambda: eng.fe_add("fq", a, b, out=o, flags=jj.JJ_ASYNC), 96)):
            for _ in range(3):
                fn()
            t = timed(eng, fn, 10)
            row[name] = {"gops": m / (t * 1e-3) / 1e9, "GBps": nbytes * m / (t * 1e-3) / 1e9,
                         "frac_of_hbm_peak": nbytes * m / (t * 1e-3) / 1e9 / hbm_peak}
        fq[f"n=2^{lg}"] = row
        for x in (a, b, o):
            x.free()
    fq_mul = {key: v["mul"] for key, v in fq.items()}

    # ---- BASELINE config 4: 2^20 fixed-base scalar-muls through the shared per-window AffineNiels table (no doublings: one
    # mixed addition per window).  Default: 12-bit windows, 4.1 MB table in global memory (L2 / L1 resident), 22 additions;
    # variant 116: 16-bit windows (50 MB, 17 additions); variant 107: round 1's 7-bit windows in shared memory (37 additions)
    m = 1 << 20
    gen = generator_mont(eng)
    kk = k if n == m else eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, m, device=True))
    fo = eng.empty((m, 20))
    fixed = {"units": m}
    for variant, key, table in ((0, "default", "21 windows x 2048 AffineNiels entries (4.1 MB) in global memory, L2/L1 resident"),
                                (116, "w16", "16 windows x 32768 entries (50 MB) in global memory"),
                                (107, "w7_smem", "36 windows x 64 entries (216 KB) staged in shared memory by one TMA bulk copy")):
        eng.set_scalar_mul_variant(variant)
        t0 = time.perf_counter()
        eng.scalar_mul_fixed_vartime(gen, kk, out=fo)  # builds and caches the table of this base
        build_ms = (time.perf_counter() - t0) * 1e3
        eng.scalar_mul_fixed_vartime(gen, kk, out=fo)
        fms = timed(eng, lambda: eng.scalar_mul_fixed_vartime(gen, kk, out=fo), 5)
        fixed[key] = {"ms": fms, "scalar_muls_per_s": m / (fms * 1e-3), "table": table,
                      "first_call_ms_incl_table_build": build_ms}
    eng.set_scalar_mul_variant(0)
    fixed["ms"], fixed["scalar_muls_per_s"] = fixed["default"]["ms"], fixed["default"]["scalar_muls_per_s"]

    # ---- wire-format path (the only one an out-of-crate Rust shim can call): 32-byte encodings + scalars in HOST
    # memory -> decode -> scalar-mul -> normalise -> encode -> 32-byte encodings in HOST memory
    enc_dev = eng.affine_to_bytes(eng.batch_normalize(pts))
    henc, henc_ptr = pinned(eng, (n, 32), np.uint8)
    hout, hout_ptr = pinned(eng, (n, 32), np.uint8)
    hok, hok_ptr = pinned(eng, (n,), np.uint8)  # pageable memory here would serialise the staged chunks
    henc[:] = enc_dev.download()

    def wire_step():
        eng._check(eng.lib.jj_scalar_mul_encoded(eng.ctx, henc.ctypes.data, hk.ctypes.data, hout.ctypes.data,
                                                 hok.ctypes.data, n, jj.JJ_OUT_BYTES))

    wire_step()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        wire_step()
    wire_s = (time.perf_counter() - t0) / e2e_steps
    okd, outd = eng.empty((n, 1), np.uint8), eng.empty((n, 32), np.uint8)

    def wire_dev(flags=0):
        eng._check(eng.lib.jj_scalar_mul_encoded(eng.ctx, enc_dev.ptr, k.ptr, outd.ptr, okd.ptr, n,
                                                 jj.JJ_DEVICE_PTRS | jj.JJ_ASYNC | jj.JJ_OUT_BYTES | flags))

    wire_dev()
    wire_dev_ms = timed(eng, wire_dev, 3)
    # the wire results are the encodings of the headline results: checked here, outside every timed region
    want_enc = eng.affine_to_bytes(eng.batch_normalize(out_all)).download()[:n]
    wire_ok = bool(hok.all() and (hout == want_enc).all() and (outd.download() == want_enc).all())
    wire_dev(jj.JJ_CHECK_SUBGROUP)  # SubgroupPoint::from_bytes semantics: decode + pairing subgroup test + multiply
    wire_sub_ms = timed(eng, lambda: wire_dev(jj.JJ_CHECK_SUBGROUP), 3)
    e2e_wire = {"value": n / wire_s, "unit": UNIT, "h2d_bytes_per_step": n * 64, "d2h_bytes_per_step": n * 33,
                "ms_per_step": wire_s * 1e3, "device_resident_ms": wire_dev_ms,
                "device_resident_with_subgroup_check_ms": wire_sub_ms, "matches_headline_results": wire_ok,
                "note": "jj_scalar_mul_encoded, JJ_OUT_BYTES: AffinePoint::to_bytes encodings + canonical scalars in pinned "
                        "HOST memory -> batch_from_bytes (reads the pinned encodings in place over PCIe; the scalars are "
                        "uploaded behind it), scalar-mul, batch_normalize + to_bytes on the device (stores the encodings in "
                        "place) -> encodings + ok[] in HOST memory; what integration/rust binds (N = 1, rank 0)"}
    parity["wire_path_matches"] = wire_ok
    parity["ok"] = bool(parity["ok"] and wire_ok)

    # ---- CPU baseline: the oracle (reference algorithm, C) on this box's host cores, bounded sample
    if _FULL_AFFINITY:
        os.sched_setaffinity(0, _FULL_AFFINITY)  # all host cores again (N > 1 bound this rank to its GPU's NUMA node)
    sample = min(n, int(os.environ.get("JJ_CPU_SAMPLE", str(12288 * cores))))
    sp, sk = hp[:sample].copy(), hk[:sample].copy()
    from oracle import binding as ob

    cpu_t = ob.time_scalar_mul(sp, sk, cores, reps=1)
    cpu_rate = sample / cpu_t
    cpu_1t = 2048 / ob.time_scalar_mul(sp[:2048], sk[:2048], 1, reps=1)
    cpu_fq = 10_000_000 / ob.time_fe_mul(ob.FQ, 10_000_000, reps=3)

    collective = {"none": "none", "nccl": "ncclAllGather of outputs each step",
                  "p2p": "all-gather fused into the kernel epilogue (NVLink P2P stores), bracketed by two 4-byte NCCL "
                         "rendezvous each step"}[gather]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (8x32-bit Montgomery, IMAD.WIDE.U32)", "data": "synthetic",
        "config": {"workload": workload_name(world, logn), "units_per_gpu": n, "total_units": world * n,
                   "output": "ExtendedPoint (160 B)" if out_fmt == "extended" else "32-byte encodings (normalise + encode fused into the kernel)",
                   "collective": collective,
                   **({"host_numa_binding": numa_all} if world > 1 else {}),
                   "cache": f"inputs+outputs {352 * n / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed); kernel is integer-bound"},
        "e2e": {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": world * n * 192,
                "d2h_bytes_per_step": world * n * unit_out, "ms_per_step": e2e_s * 1e3,
                "note": ("jj_scalar_mul with pinned HOST buffers: chunked H2D, kernel, D2H on two streams" if world == 1 else
                         "jj_scalar_mul_sharded_n on every rank: this rank's inputs from pinned HOST memory (chunked H2D "
                         "overlapping the kernels), results stored into every rank's device buffer (the gather), own "
                         "block read back to HOST memory; wall time incl. barriers, max over ranks")},
        "e2e_wire": e2e_wire,
        "gpu_launches": launches,
        "clocks": clocks,
        "parity_check": parity,
        "roofline": {"bound": "integer multiplier (IMAD.WIDE.U32 issue on the FMA-heavy pipe)", "achieved": achieved_imad / 1e12,
                     "peak": imad_peak / 1e12, "unit": "T thread-ops/s of IMAD.WIDE.U32", "frac": achieved_imad / imad_peak,
                     "imads_per_unit_algorithmic": IMADS_PER_UNIT, "imads_issued_per_unit": NCU_IMADS_ISSUED_PER_UNIT,
                     "frac_issued": NCU_IMADS_ISSUED_PER_UNIT * n / (kms * 1e-3) / imad_peak,
                     "peak_source": "measured live by jj_measure_imad_peak: best of three register-only probes (IMAD.WIDE.U32 chains "
                                    "with register operands / with an immediate multiplier at 64 warps/SM, dependent Fq-product "
                                    "chains at 16 warps/SM), best of 6 calls; not in MEASURED_PEAKS.json; nominal 32 lanes/clk/SM x "
                                    "148 x 1.965 GHz = 9.31", "peak_clocks": peak_clocks,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_PROFILE,
                     "kernel": "k_scalar_mul<512,1,GMEM>", "kernel_ms": kms},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "peak_source": peak_src,
                         "note": "352 algorithmic B/unit: HBM is not the bound of this kernel; DRAM traffic above the "
                                 "algorithmic 3.69e8 B per 2^20 launch is write-back of the L2 window-table scratch"},
        "fq_mul": fq_mul, "fq_ops": fq, "fixed_base": fixed,
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample} units of the same batch, {cores} threads, {cpu_t:.1f} s, "
                                   "reference ladder (C restatement in oracle/)",
                         "single_thread": {"scalar_muls_per_s": cpu_1t, "fq_muls_per_s": cpu_fq,
                                           "sample": "2048 scalar-muls; 1e7 dependent Fq muls, best of 3"}},
    }
    emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
