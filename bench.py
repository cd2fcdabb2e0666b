#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched Jubjub engine.

Metric (BASELINE.json): variable-base scalar-muls / second.  Workload (BASELINE.json configs[2],
the configuration the metric is quoted on): 2^20 variable-base `ExtendedPoint * Fr` scalar-muls per
GPU, points P_i = [t_i] G (full-order), scalars uniform in [0, r), both from the SplitMix64 streams
of SURVEY.md section 8d.  One step = one pass of the hot path over that batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); each rank owns a contiguous shard of 2^20 units
(weak scaling) and every step ends with the NCCL all-gather of all ranks' output points
(jj_scalar_mul_sharded).  Timing: CUDA events on the engine's stream, barrier + synchronize on both
sides, max over ranks.  Rank 0 prints ONE JSON line.

`--impl reference` times the reference *algorithm* (bitwise 252-step double-and-add ladder on
4 x u64 Montgomery limbs) as restated in C in oracle/ -- the reference itself is Rust and cannot be
built in this image -- on all host cores, on a bounded sample of the same workload.
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
LOG_N = 20
METRIC = "variable_base_scalar_muls_per_sec"
UNIT = "scalar-muls/s"
WORKLOAD = "2^20 variable-base ExtendedPoint*Fr scalar-muls per GPU (P_i=[t_i]G, 252-bit scalars, SplitMix64 streams)"
# algorithmic work per unit (DESIGN.md "roofline"): bytes = 160 B point + 32 B scalar in, 160 B point out;
# IMAD.WIDE.U32 = signed radix-16 window: 252 doublings (4S+3M), 7+~59 additions (8M), 16 M for to_niels,
# S = 84 and M = 112 multiplier instructions: the minimum of 8x32-bit Montgomery with q's special low limbs.  (The
# shipped kernels issue S = 91, M = 119 -- one extra multiply per reduction row but the last replaces three ALU instructions,
# DESIGN.md section 5 -- so the fraction below undercounts the pipe's real occupancy: ncu reads 86 %.)
BYTES_PER_UNIT = 352
# DRAM bytes of one 2^20-unit launch of the dominant kernel, from the committed ncu capture
# profiles/r01d_ncu_scalar_mul_default_n1048576.csv (dram__bytes_read.sum + dram__bytes_write.sum)
NCU_DRAM_BYTES_PER_LAUNCH = 235.93e6 + 752.27e6
IMADS_PER_UNIT = 252 * (4 * 84 + 3 * 112) + (7 + 63 * 15 / 16) * 8 * 112 + 16 * 112
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

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

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
    pts = eng.scalar_mul_fixed(generator_mont(eng), t)
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, n, first=first, device=True))
    t.free()
    return pts, k


def pinned(eng, shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    eng._check(eng.lib.jj_host_alloc(eng.ctx, nbytes, C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def cpu_baseline(sample_points, sample_scalars, cores):
    from oracle import binding as ob

    t = ob.time_scalar_mul(sample_points, sample_scalars, cores, reps=1)
    return len(sample_points) / t, t


def run_reference(args, rank, emit=print):
    """Reference arm: the reference algorithm (C restatement, oracle/) on all host cores."""
    if rank != 0:
        return
    from oracle import binding as ob
    from oracle import model as M

    cores = os.cpu_count() or 1
    sample = int(os.environ.get("JJ_REF_SAMPLE", str(min(1 << LOG_N, 2048 * cores))))
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
    desc = f"{sample} of 2^{LOG_N} units per step (first {sample} of the same streams), {cores} threads"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (4x64 Montgomery)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm (bitwise ladder, src/lib.rs:356-379) as a C restatement; the Rust reference "
                "cannot be built in this image (no cargo/rustc; Fq lives in the un-vendored bls12_381 crate)",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # Keep stdout to the single JSON line: anything native libraries print to fd 1 while the bench runs
    # (e.g. NCCL's version banner) is sent to stderr; the JSON goes to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(line, flush=True)

    if args.impl == "reference":
        run_reference(args, rank, emit)
        return
    args.warmup = max(args.warmup, 3)

    import jubjub_b200 as jj

    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner at INFO/VERSION level
        os.environ["NCCL_DEBUG"] = os.environ.get("JJ_NCCL_DEBUG", "WARN")
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = jj.Engine(local)
    n = 1 << LOG_N
    pts, k = make_inputs(eng, n, first=rank * n)
    unit_out = 160
    out_all = eng.empty((world * n, 20))
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
            eng.scalar_mul_sharded(pts, k, out_all, async_=True)
        else:
            eng.scalar_mul(pts, k, out=out_all, flags=jj.JJ_ASYNC)

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

    # ---- e2e: host (pinned) buffers through the C ABI on every rank, H2D + D2H inside the timed region
    eng.set_peer_outputs(None)
    hp, hp_ptr = pinned(eng, (n, 20), np.uint64)
    hk, hk_ptr = pinned(eng, (n, 32), np.uint8)
    ho, ho_ptr = pinned(eng, (n, 20), np.uint64)
    hp[:] = pts.download()
    hk[:] = k.download()
    e2e_steps = max(2, min(args.steps, 5))
    eng.scalar_mul(hp, hk, out=ho)  # warm-up (staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.scalar_mul(hp, hk, out=ho)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert ho[-1].any(), "e2e produced no output"
    if dist is not None:
        tmax = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- kernel-only duration of the dominant kernel (same launches, no gather), for the roofline
    eng.timer_start()
    for _ in range(args.steps):
        eng.scalar_mul(pts, k, out=out_all, flags=jj.JJ_ASYNC)
    kms = eng.timer_stop() / args.steps
    hbm_peak, peak_src = measured_peaks()
    achieved_gbs = BYTES_PER_UNIT * n / (kms * 1e-3) / 1e9
    imad_peak = eng.imad_peak()
    achieved_imad = IMADS_PER_UNIT * n / (kms * 1e-3)

    # ---- secondary metric: Fq mul GOPS (BASELINE config 2: 2^20, L2-resident; and 2^26, HBM-sized)
    fq = {}
    for logn in (20, 26):
        m = 1 << logn
        a = eng.fe_stream("fq", SEED0, m, device=True)
        b = eng.fe_stream("fq", SEED0 + 1, m, device=True)
        o = eng.empty((m, 4))
        for _ in range(3):
            eng.fe_mul("fq", a, b, out=o, flags=jj.JJ_ASYNC)
        reps = 10
        eng.sync()
        eng.timer_start()
        for _ in range(reps):
            eng.fe_mul("fq", a, b, out=o, flags=jj.JJ_ASYNC)
        t = eng.timer_stop() / reps
        fq[f"n=2^{logn}"] = {"gops": m / (t * 1e-3) / 1e9, "GBps": 96 * m / (t * 1e-3) / 1e9,
                             "frac_of_hbm_peak": 96 * m / (t * 1e-3) / 1e9 / hbm_peak}
        for x in (a, b, o):
            x.free()

    # ---- CPU baseline: the oracle (reference algorithm, C) on this box's host cores, bounded sample
    cores = os.cpu_count() or 1
    sample = min(n, int(os.environ.get("JJ_CPU_SAMPLE", str(12288 * cores))))
    sp, sk = hp[:sample].copy(), hk[:sample].copy()
    cpu_rate, cpu_t = cpu_baseline(sp, sk, cores)
    from oracle import binding as ob

    cpu_1t = 2048 / ob.time_scalar_mul(sp[:2048], sk[:2048], 1, reps=1)
    cpu_fq = 10_000_000 / ob.time_fe_mul(ob.FQ, 10_000_000, reps=3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (8x32-bit Montgomery, IMAD.WIDE.U32)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "units_per_gpu": n, "output": "ExtendedPoint (160 B)",
                   "collective": {"none": "none", "nccl": "ncclAllGather of outputs each step",
                                  "p2p": "all-gather fused into the kernel epilogue (NVLink P2P stores) + 4-byte "
                                         "NCCL rendezvous each step"}[gather],
                   "cache": "inputs+outputs 352 MB per GPU > 126 MB L2 (no flush needed); kernel is integer-bound"},
        "e2e": {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": world * n * 192,
                "d2h_bytes_per_step": world * n * unit_out,
                "note": "jj_scalar_mul with pinned HOST buffers on every rank (max wall time over ranks): chunked "
                        "H2D, kernel, D2H on two streams; no gather"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                     "traffic_note": "bytes per launch, ncu capture profiles/r01d_ncu_scalar_mul_default_n1048576.csv; "
                                     "algorithmic bytes per launch = 352 B x 2^20 = 3.69e8; the excess is write-back of the L2 window-table "
                                     "scratch (0.4 % of HBM bandwidth, not re-reads of inputs)",
                     "peak_source": peak_src,
                     "kernel": "k_scalar_mul", "kernel_ms": kms,
                     "note": "dominant kernel moves 352 algorithmic B/unit and is bound by the integer multiplier, "
                             "not HBM: see roofline_int"},
        "roofline_int": {"bound": "imad.wide.u32 issue (fmaheavy pipe)", "achieved": achieved_imad, "peak": imad_peak,
                         "unit": "IMAD.WIDE.U32 thread-ops/s", "frac": achieved_imad / imad_peak,
                         "imads_per_unit": IMADS_PER_UNIT, "peak_source": "measured live (jj_measure_imad_peak)"},
        "fq_mul": fq,
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
