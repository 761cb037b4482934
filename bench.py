#!/usr/bin/env python
"""bench.py — block SpMV throughput of the B200 multiply path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c1|c4]

A "step" is one multiply y = A*x (mul!(y, A, x)) over the whole synthetic matrix.
N = 1 workload (default): configs[1] of BASELINE.json — SymmetricBlockMatrix ComplexF64, 1 M unknowns,
leaves 20–200, half-stored off-diagonals (~12.7 GB in HBM); working set >> L2, so no L2 flush is needed.
N > 1 (torchrun): the same matrix cut into N nnz-balanced block-row slabs, one per rank; every step
all-gathers x over NCCL and each rank writes its own y slice ("scaling": "strong").

Prints ONE JSON line (rank 0). `value` = algorithmic GB/s with operands resident in HBM (CUDA events,
max over ranks); `e2e` = the same metric through the C-ABI host-pointer call (bsm_mul_host) from pinned
host buffers, H2D of x and D2H of y inside the timed region; `roofline` = the dominant kernel against the
measured HBM peak; `cpu_baseline` = the oracle (C restatement of the reference schedule) on the host
cores over a bounded sample.  --impl reference times that CPU restatement alone (Julia is not
installable in this image, so the oracle port stands in for the reference's threaded mul!).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "block SpMV effective HBM GB/s"


# ----------------------------------------------------------------------------- workloads
def workload_spec(name, scale):
    if name == "c2":
        n = max(2000, int(1_000_000 * scale))
        return dict(kind="SymmetricBlockMatrix", dtype="c128", n=n,
                    desc=f"C2 SymmetricBlockMatrix ComplexF64 N={n} leaves 20-200 k_near=6 half-stored, mul!(y,A,x)")
    if name == "c3":
        n = max(2000, int(4_000_000 * scale))
        return dict(kind="VBCRS", dtype="f64", n=n,
                    desc=f"C3 VBCRS Float64 {n} rows blocks 8-64, mul!(y,A,x)")
    if name == "c1":
        return dict(kind="BlockSparseMatrix", dtype="f64", n=10_000,
                    desc="C1 BlockSparseMatrix Float64 10000^2, 2000 blocks 32x32, y=A*x (L2-resident)")
    if name == "c4":
        g = max(4, int(round(64 * np.sqrt(scale))))
        return dict(kind="BlockSparseMatrix", dtype="f32", n=g * 1024, grid=g,
                    desc=f"C4 BlockSparseMatrix Float32 1024^2 blocks 5% of {g}x{g} grid, transpose(A)*x")
    if name == "c5":
        n = max(6400, int(1_000_000 * scale) // 32 * 32)
        return dict(kind="BlockSparseMatrix", dtype="f64", n=n, nblocks=max(1000, int(200_000 * scale)), nrhs=64,
                    desc=f"C5 BlockSparseMatrix Float64 N={n}, {max(1000, int(200_000 * scale))} blocks 32x32, "
                         f"Y = A*X with 64 right-hand sides (SpMM)")
    raise SystemExit(f"unknown workload {name}")


def build_workload(name, scale, rank=0, world=1, threads=8):
    """Returns (host matrix of this rank's slab, op, owned range (lo, hi) or None, x slice bounds per rank)."""
    from bsm_b200 import generators as G
    spec = workload_spec(name, scale)
    if name == "c2":
        if world == 1:
            return G.symmetric_nearfield(seed=2, n=spec["n"], threads=threads), "N", None, None
        S = G.NearfieldStructure(2, spec["n"], 20, 200, 6)
        cuts = S.partition(world)
        A = G.symmetric_nearfield(seed=2, n=spec["n"], threads=threads, leaves=(int(cuts[rank]), int(cuts[rank + 1])))
        rb = S.bounds[cuts]
        return A, "N", (int(rb[rank]), int(rb[rank + 1])), rb
    if name == "c3":
        return G.vbcrs_variable(seed=3, n=spec["n"], threads=threads), "N", None, None
    if name == "c1":
        return G.blocksparse_uniform(seed=1, threads=threads), "N", None, None
    if name == "c5":
        return G.blocksparse_uniform(seed=5, n=spec["n"], nblocks=spec["nblocks"], threads=threads), "N", None, None
    return G.blocksparse_large(seed=4, grid=spec["grid"], threads=threads), "T", None, None


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []

    def _poll(self):
        import pynvml as N
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        while not self._stop:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
                rs = N.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        # NVML polled in-process every ~2 ms (the timed region lasts tens of ms); nvidia-smi -lms as fallback
        try:
            import pynvml as N
            N.nvmlInit()
            self.nvml = N
            self._stop = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            N = self.nvml
            self._stop = True
            self.thread.join(timeout=1.0)
            names = {"hw_slowdown": getattr(N, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, _, r in self.samples))
            sm = [a for a, _, _ in self.samples]
            mx = [b for _, b, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml polled during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_sample(name, scale_hint, steps, warmup, min_seconds=8.0):
    """Times the oracle (C restatement of the reference schedule, all host threads) on a bounded sample of
    the workload. Returns (GB/s, cores, sample description, seconds per multiply)."""
    import bsm_b200 as B
    from bsm_b200 import _lib as L
    from bsm_b200 import generators as G
    from helpers import to_oracle
    from oracle import oracle_np as O

    threads = os.cpu_count() or 1
    nrhs = 1
    if name == "c2":
        n = min(100_000, workload_spec(name, scale_hint)["n"])
        A = G.symmetric_nearfield(seed=2, n=n, threads=min(threads, 16))
        sample = f"same generator at N={n} (~{n / 1e6 * 12.7:.2f} GB ComplexF64), {threads} threads"
        op = "N"
    elif name == "c3":
        n = min(1_000_000, workload_spec(name, scale_hint)["n"])
        A = G.vbcrs_variable(seed=3, n=n, threads=min(threads, 16))
        sample = f"same generator at {n} rows, {threads} threads"
        op = "N"
    elif name == "c1":
        A = G.blocksparse_uniform(seed=1)
        sample, op = f"full C1 matrix, {threads} threads", "N"
    elif name == "c5":
        A = G.blocksparse_uniform(seed=5, n=100_000 // 32 * 32, nblocks=20_000)
        sample, op = f"same generator at N={A.size[0]}, 20000 blocks, 64 right-hand sides as a column loop, {threads} threads", "N"
        nrhs = 64
    else:
        A = G.blocksparse_large(seed=4, grid=16)
        sample, op = f"same generator on a 16x16 grid (13 blocks), {threads} threads", "T"
    work = A.device(device=L.DEVICE_NONE).work(op, nrhs=nrhs)
    OA = to_oracle(A)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(A.size[1]).astype(A.dtype)
    if isinstance(OA, O.OSBM):
        C = O.CSbm(OA, threads)
        run = lambda: C.mul(x, op)
    elif isinstance(OA, O.OVBCRS):
        run = lambda: O.c_mul_vbcrs(OA, x, op, threads=threads)
    else:
        run = lambda: O.c_mul_bsm(OA, x, op, threads=threads)
    if nrhs > 1:      # LinearMaps applies a matrix right-hand side column by column
        one = run
        run = lambda: [one() for _ in range(nrhs)]
    for _ in range(max(1, min(warmup, 2))):
        run()
    times = []
    t_begin = time.perf_counter()
    while len(times) < steps or (time.perf_counter() - t_begin) < min_seconds:
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
        if len(times) >= 200 or (time.perf_counter() - t_begin) > 30:
            break
    sec = float(np.median(times))
    return work["bytes"] / sec / 1e9, threads, sample, sec, work["flops"] / sec / 1e9


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gbs, cores, sample, sec, gflops = cpu_sample(args.workload, args.scale, args.steps, args.warmup)
    spec = workload_spec(args.workload, args.scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
        "gflops": gflops,
        "config": {"workload": spec["desc"], "note": "CPU restatement of the reference schedule (oracle port); "
                   "Julia is not installable in this image"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (development only)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--xchg", default="peer", choices=["peer", "nccl"],
                    help="N > 1, one right-hand side: peer = x read from its owners over NVLink inside the kernels "
                         "(no collective); nccl = all-gather of x, then multiply")
    ap.add_argument("--broadcasts", action="store_true", help="N > 1: grouped in-place broadcasts instead of the all-gather")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: all-gather, then multiply, on one stream")
    ap.add_argument("--op", default=None, choices=["N", "T", "C"], help="override the workload's op (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import bsm_b200 as B
    from bsm_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = workload_spec(args.workload, args.scale)

    t0 = time.time()
    host_threads = max(1, (os.cpu_count() or 8) // max(world, 1))
    A, op, own, rb = build_workload(args.workload, args.scale, rank, world, threads=min(host_threads, 32))
    if args.op:
        op = args.op
    t_gen = time.time() - t0
    t0 = time.time()
    full_work = None
    if world > 1:
        # one process per GPU: nnz-balanced block-row slabs, libbsm_b200's own NCCL communicator
        from bsm_b200.dist import Comm, SlabMatrix
        comm = Comm.from_torch(local)
        comm.set_overlap(not args.no_overlap)
        comm.set_collective(args.broadcasts)
        if rb is None:        # generic partition of the full host matrix (every rank generated it)
            full_work = A.device(device=L.DEVICE_NONE).work(op, nrhs=spec.get("nrhs", 1))
            SM = SlabMatrix(A, comm, ops=(op,), variant=args.variant)
        else:                 # c2: every rank generated only its slab's blocks
            SM = SlabMatrix(A, comm, cuts=rb, variant=args.variant)
        D, own, rb = SM.local, SM.own, SM.cuts
    else:
        D = B.DeviceMatrix(A, device=local, variant=args.variant)
    torch.cuda.synchronize()
    t_pack = time.time() - t0
    tdt = {"c128": torch.complex128, "f64": torch.float64, "f32": torch.float32}[spec["dtype"]]
    nin = A.size[1] if op == "N" else A.size[0]
    nout = A.size[0] if op == "N" else A.size[1]
    nrhs = spec.get("nrhs", 1)
    work = D.work(op, nrhs=nrhs)
    if world > 1 and full_work is not None:
        work = full_work
    elif world > 1:
        # whole-job algorithmic bytes: every stored entry once (slabs duplicate boundary blocks, that is
        # overhead, not work) — computed from the structure on rank 0's formula for the full matrix
        from bsm_b200 import generators as G
        S = G.NearfieldStructure(2, spec["n"], 20, 200, 6)
        sz = S.sizes.astype(np.int64)
        stored = int((sz * sz).sum() + sum(int(sz[i]) * int(sz[S.near[i]].sum()) for i in range(1, S.nl)))
        work = {"bytes": stored * 16.0 + 2 * spec["n"] * 16.0, "flops": 8.0 * (2 * stored - int((sz * sz).sum())),
                "index_table_bytes": 0.0}

    g = torch.Generator(device="cpu").manual_seed(1234)
    if nrhs == 1:
        x_host = torch.randn(nin, dtype=tdt, generator=g).pin_memory()
        y_host = torch.empty(nout, dtype=tdt).pin_memory()
        y_dev = torch.zeros(nout, dtype=tdt, device=dev)
    else:   # column-major (nin x nrhs) / (nout x nrhs)
        x_host = torch.randn((nrhs, nin), dtype=tdt, generator=g).pin_memory().t()
        y_host = torch.empty((nrhs, nout), dtype=tdt).pin_memory().t()
        y_dev = torch.zeros((nrhs, nout), dtype=tdt, device=dev).t()
    x_full = x_host.to(dev) if nrhs == 1 else x_host.t().to(dev).t()

    peer = world > 1 and nrhs == 1 and args.xchg == "peer"
    if peer:
        try:
            xs = comm.alloc(nin, np.dtype({"c128": np.complex128, "f64": np.float64, "f32": np.float32}[spec["dtype"]]))
            xs.copy_(x_full)
            ok = 1
        except Exception as exc:       # no peer access between the GPUs of this box: NCCL path
            print(f"[rank {rank}] peer mode unavailable ({exc}); falling back to the NCCL all-gather", file=sys.stderr)
            ok = 0
        t_ok = torch.tensor([ok], device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        peer = bool(t_ok.item())
    if peer:
        x_full = xs

        def step():
            SM.mul_peer(op, x_full, y_dev)   # bsm_mul_dist_peer: flag barrier, multiply reading x over NVLink, flag barrier
    elif world > 1:
        def step():
            SM.mul(op, x_full, y_dev)   # bsm_mul_dist: NCCL all-gather of the x slabs (in place), then the slab multiply
    else:
        def step():
            D.mul(op, x_full, y_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()          # ranks finish generating / packing at different times
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l2_resident = work["bytes"] <= 4 * 126e6
    barrier()
    # single GPU: the per-kernel event brackets (bsm_set_profiling, a ring of 64 triples inside bsm_mul) are
    # recorded IN the timed region, so roofline.achieved is the dominant kernel's average over these very steps
    inline_prof = world == 1 and not l2_resident
    k_ms = f_ms = None
    if inline_prof:
        D.set_profiling(True)
    if not l2_resident:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        if inline_prof:
            k_ms, f_ms = D.profile()
            D.set_profiling(False)
    else:
        # working set fits the 126 MB L2: flush it (256 MB write) before every timed iteration and time each
        # multiply with its own pair of events, so the blocks really come from HBM
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush.zero_()
            a.record()
            step()
            b.record()
        barrier()
        ms_total = float(sum(a.elapsed_time(b) for a, b in evs))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps

    # per-kernel device time of the dominant kernel (CUDA events inside bsm_mul, same stream); N > 1 and the
    # L2-flushed small case measure it in a separate loop on the rank-local handle
    D.set_profiling(True)
    main_ms, fin_ms = [], []
    for _ in range(0 if inline_prof else max(5, min(args.steps, 20))):
        # the events of the LAST multiply of a back-to-back burst: the kernel is timed in steady state (clocks up,
        # no host synchronisation in front of it), like the steps of the timed region
        for k in range(1 if l2_resident else 4):
            if l2_resident:
                flush.zero_()
            D.mul(op, x_full, y_dev)
        a, b = D.profile()
        main_ms.append(a)
        fin_ms.append(b)
    D.set_profiling(False)
    if not inline_prof:
        k_ms, f_ms = float(np.mean(main_ms)), float(np.mean(fin_ms))
    fin_ms = [f_ms]
    local_work = D.work(op, nrhs=nrhs)

    # end to end through the host-pointer C-ABI call: pinned host x → H2D → multiply → D2H → host y
    e2e = None
    if world == 1:
        xh, yh = (x_host.numpy(), y_host.numpy()) if nrhs == 1 else (x_host.t().numpy().T, y_host.t().numpy().T)
        for _ in range(2):
            D.mul(op, xh, yh)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            D.mul(op, xh, yh)
        e2e_s = (time.perf_counter() - t0) / args.steps
        e2e = {"value": work["bytes"] / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size()),
               "d2h_bytes_per_step": int(y_host.numel() * y_host.element_size())}
    else:
        # N > 1: each rank copies its x slice in and its y slice out every step
        rows = own[1] - own[0]
        if nrhs == 1:
            xs_h = x_host[own[0]:own[1]].clone().pin_memory()
            ys_h = torch.empty(rows, dtype=tdt).pin_memory()
        else:   # slab rows of all columns: contiguous pinned buffers, strided placement done on the device
            xs_h = x_host[own[0]:own[1]].t().contiguous().pin_memory()
            ys_h = torch.empty((nrhs, rows), dtype=tdt).pin_memory()
            xs_d, ys_d = torch.empty_like(xs_h, device=dev), torch.empty_like(ys_h, device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            if nrhs == 1:
                x_full[own[0]:own[1]].copy_(xs_h, non_blocking=True)
            else:
                xs_d.copy_(xs_h, non_blocking=True)
                x_full[own[0]:own[1]].copy_(xs_d.t())
            step()
            if nrhs == 1:
                ys_h.copy_(y_dev[own[0]:own[1]], non_blocking=True)
            else:
                ys_d.copy_(y_dev[own[0]:own[1]].t())
                ys_h.copy_(ys_d, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": work["bytes"] / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(nin * nrhs * x_host.element_size()),
               "d2h_bytes_per_step": int(nout * nrhs * y_host.element_size())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = local_work["bytes"] / (k_ms * 1e-3) / 1e9
    stats = D.plan_stats(op)
    kernel_name = max(stats["bytes"], key=stats["bytes"].get)
    if kernel_name == "sym_fused_tma_kernel" and args.variant == 2:
        kernel_name = "sym_fused_kernel"
    tensor = None
    if nrhs >= 8 and stats["spmm"] and args.variant != 1:
        kernel_name = "spmm_dmma_kernel"
        # FP64 tensor-pipe denominator: cuBLAS DGEMM measured here (MEASURED_PEAKS.json holds no FP64 figure)
        a64 = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        b64 = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a64, b64)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = 1e9
        for _ in range(5):
            ev[0].record()
            torch.matmul(a64, b64)
            ev[1].record()
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        dgemm_tf = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
        ach_tf = local_work["flops"] / (k_ms * 1e-3) / 1e12
        tensor = {"bound": "tensor", "achieved": ach_tf, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach_tf / dgemm_tf,
                  "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (best of 5)"}
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists() and args.scale == 1.0 and world == 1 and args.variant == 0 and not args.op:
        ent = json.loads(tfile.read_text()).get(args.workload)
        if ent and ent["kernel"] == kernel_name:
            traffic = ent["dram_bytes_per_launch"]     # dram__bytes_read.sum + write.sum, ncu --set full (profiles/)
    line = {
        "metric": METRIC, "value": work["bytes"] / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
        "gflops": work["flops"] / (ms_step * 1e-3) / 1e9,
        "config": {"workload": spec["desc"], "op": op, "l2": "working set larger than L2 (no flush needed)"
                   if not l2_resident else "working set fits L2: L2 flushed (256 MB write) before every timed iteration, "
                                           "each multiply timed by its own CUDA events",
                   "variant": {0: "auto", 1: "gather", 2: "fused", 3: "color", 4: "fused_tma"}[args.variant],
                   "parallelism": (f"block-row slabs x{world}, " +
                                   ("x read from its owners' peer-mapped arrays over NVLink inside the kernels (no collective)"
                                    if peer else f"NCCL all-gather of x ({'sequential' if args.no_overlap else 'overlapped with the rank-local slices'})"))
                   if world > 1 else "single GPU",
                   "algorithmic_bytes": work["bytes"], "flops": work["flops"],
                   "gen_s": round(t_gen, 1), "pack_s": round(t_pack, 1), "plan": stats},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms,
                     "finalize_ms": float(np.mean(fin_ms)), "peak_source": peak_src,
                     "bytes_per_launch": local_work["bytes"]},
        "e2e": e2e,
        # our kernels per step: the multiply's own launches (+ the four flag-barrier kernels of peer mode)
        "gpu_launches": int(args.steps * ((D.launch_count(op) if nrhs == 1 or tensor is None else 1) + (4 if peer else 0))),
        "clocks": clocks,
    }
    if tensor is not None:
        line["roofline_tensor"] = tensor
    if world == 1 and not args.no_cpu_baseline:
        gbs, cores, sample, sec, gflops = cpu_sample(args.workload, args.scale, 3, 1)
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample,
                                "ms_per_multiply": sec * 1e3}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
