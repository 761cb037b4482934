#!/usr/bin/env python
"""bench.py — block SpMV / SpMM throughput of the B200 multiply path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1..c5]

A "step" is one multiply y = op(A)*x (mul!(y, A, x)) over the whole synthetic matrix.
Default workload: configs[1] of BASELINE.json ("c2") — SymmetricBlockMatrix ComplexF64, 1 M unknowns, leaves
20–200, half-stored off-diagonals (~12.8 GB in HBM; working set >> L2, no flush needed). With the default
workload the line also carries `also`: compact results of c3 (VBCRS Float64, 4 M rows) and c5 (BlockSparseMatrix
Float64 x 64 right-hand sides), the two other configurations BASELINE.json names for 1/2/4/8 GPUs.

N > 1 (torchrun, one process per GPU):
  c2, c3   nnz-balanced block-row slabs, x kept sharded in peer-mapped arrays; the multiply kernels fetch the x
           entries they need from their owners over NVLink and carry both barriers themselves (no collective, no
           extra launch) — bsm_mul_dist_peer;
  c5       2-D grid of ranks: block-row slabs of A x column groups of the 64 right-hand sides (1 x 2, 2 x 2,
           2 x 4 for N = 2, 4, 8: dist.rhs_grid), no exchange at all.
"scaling": "strong" (the total work is fixed as N grows).

ONE JSON line (rank 0). `value` = algorithmic GB/s with operands resident in HBM (CUDA events, max over ranks);
algorithmic bytes = stored entries * s + (inputs + outputs) * s * nrhs, half-stored blocks counted once, index
tables NOT counted (SURVEY §8d "without" figure — identical for the reference arm). `parity` = rel ||dy||/||y|| of
the GPU result against the C oracle on the SAME full-size matrix and x (every rank checks its own y slice; the run
fails above the tolerance). `e2e` = the same metric through the host-pointer C-ABI call (bsm_mul_host /
bsm_mul_dist_peer_host) from pinned host buffers, H2D of x and D2H of y inside the timed region. `roofline` = the
dominant kernel against the measured HBM peak. `cpu_baseline` / --impl reference = the oracle (C restatement of the
reference schedule; Julia cannot be installed in this image) on the host cores, same matrix, same x; the reference
arm never loads libbsm_b200.so.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "block SpMV effective HBM GB/s"
ALSO = ("c3", "c5")
TOL = {"f32": 1e-5, "f64": 1e-12, "c128": 1e-12}
NPDT = {"c128": np.complex128, "f64": np.float64, "f32": np.float32}


# ----------------------------------------------------------------------------- workloads
def workload_spec(name, scale, hard=""):
    if name == "c2":
        n = max(2000, int(1_000_000 * scale))
        tag = {"": "", "permuted": " PERMUTED (arbitrary unsorted index vectors)",
               "scattered": " SCATTERED near leaves (drawn from all lower leaves)",
               "permuted+scattered": " PERMUTED + SCATTERED"}[hard]
        return dict(kind="SymmetricBlockMatrix", dtype="c128", n=n, op="N",
                    desc=f"C2 SymmetricBlockMatrix ComplexF64 N={n} leaves 20-200 k_near=6 half-stored{tag}, mul!(y,A,x)")
    if name == "c3":
        n = max(2000, int(4_000_000 * scale))
        return dict(kind="VBCRS", dtype="f64", n=n, op="N", desc=f"C3 VBCRS Float64 {n} rows blocks 8-64, mul!(y,A,x)")
    if name == "c1":
        return dict(kind="BlockSparseMatrix", dtype="f64", n=10_000, op="N",
                    desc="C1 BlockSparseMatrix Float64 10000^2, 2000 blocks 32x32, y=A*x (L2-resident)")
    if name == "c4":
        g = max(4, int(round(64 * np.sqrt(scale))))
        return dict(kind="BlockSparseMatrix", dtype="f32", n=g * 1024, grid=g, op="T",
                    desc=f"C4 BlockSparseMatrix Float32 1024^2 blocks 5% of {g}x{g} grid, transpose(A)*x")
    if name == "c5":
        n = max(6400, int(1_000_000 * scale) // 32 * 32)
        nb = max(1000, int(200_000 * scale))
        return dict(kind="BlockSparseMatrix", dtype="f64", n=n, nblocks=nb, nrhs=64, op="N",
                    desc=f"C5 BlockSparseMatrix Float64 N={n}, {nb} blocks 32x32, Y = A*X with 64 right-hand sides (SpMM)")
    raise SystemExit(f"unknown workload {name}")


def build_workload(name, scale, rank=0, world=1, threads=8, hard=""):
    """Returns (host matrix — for c2 on N > 1 only this rank's slab —, slab row bounds per rank or None)."""
    from bsm_b200 import generators as G
    spec = workload_spec(name, scale, hard)
    if name == "c2":
        kw = dict(permuted="permuted" in hard, scattered="scattered" in hard)
        if world == 1:
            return G.symmetric_nearfield(seed=2, n=spec["n"], threads=threads, **kw), None
        S = G.NearfieldStructure(2, spec["n"], 20, 200, 6, scattered=kw["scattered"])
        cuts = S.partition(world)
        A = G.symmetric_nearfield(seed=2, n=spec["n"], threads=threads, leaves=(int(cuts[rank]), int(cuts[rank + 1])), **kw)
        return A, S.bounds[cuts]
    if name == "c3":
        return G.vbcrs_variable(seed=3, n=spec["n"], threads=threads), None
    if name == "c1":
        return G.blocksparse_uniform(seed=1, threads=threads), None
    if name == "c5":
        return G.blocksparse_uniform(seed=5, n=spec["n"], nblocks=spec["nblocks"], threads=threads), None
    return G.blocksparse_large(seed=4, grid=spec["grid"], threads=threads), None


def host_work(A, op, nrhs=1):
    """Algorithmic bytes / flops of one multiply from the HOST container alone (no library call): stored entries
    once (half-stored symmetric blocks counted once), x read once, y written once, no index tables."""
    s = np.dtype(A.dtype).itemsize
    if hasattr(A, "offdiagonals"):
        d = sum(int(b.size) for b in A.diagonals)
        o = sum(int(b.size) for b in A.offdiagonals)
        stored, applied = d + o, d + 2 * o
    else:
        stored = applied = sum(int(b.size) for b in A.blocks)
    nin = A.size[1] if op == "N" else A.size[0]
    nout = A.size[0] if op == "N" else A.size[1]
    return {"bytes": float(stored * s + (nin + nout) * s * nrhs),
            "flops": (8.0 if np.dtype(A.dtype).kind == "c" else 2.0) * applied * nrhs}


def c2_whole_job_work(spec, scattered=False):
    from bsm_b200 import generators as G
    S = G.NearfieldStructure(2, spec["n"], 20, 200, 6, scattered=scattered)
    sz = S.sizes.astype(np.int64)
    d = int((sz * sz).sum())
    o = sum(int(sz[i]) * int(sz[S.near[i]].sum()) for i in range(1, S.nl))
    return {"bytes": float((d + o) * 16 + 2 * spec["n"] * 16), "flops": 8.0 * (d + 2 * o)}


def make_oracle(A, threads):
    """Pre-marshalled C-oracle multiply of the host matrix: f(x, op) -> y (TEST INFRASTRUCTURE, CPU)."""
    from helpers import to_oracle
    from oracle import oracle_np as O
    OA = to_oracle(A)
    if isinstance(OA, O.OSBM):
        C = O.CSbm(OA, threads)
    elif isinstance(OA, O.OVBCRS):
        C = O.CVbcrs(OA, threads)
    else:
        C = O.CBsm(OA, threads)
    return lambda x, op: C.mul(x, op)


def host_x(n, nrhs, dtype, seed=1234):
    """The right-hand side every arm uses (NumPy, column-major n x nrhs, or a vector)."""
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    shape = (n,) if nrhs == 1 else (nrhs, n)
    x = rng.standard_normal(shape)
    if dt.kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    x = x.astype(dt)
    return x if nrhs == 1 else x.T       # (n, nrhs) view with column-major storage


# ----------------------------------------------------------------------------- clocks
_SAMPLER_SRC = r"""
import sys, time
import pynvml as N
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
print("max", N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM), flush=True)
while True:
    try:
        print(time.time(), N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), reasons(h), flush=True)
    except Exception:
        pass
    time.sleep(0.0005)
"""


class ClockSampler:
    """NVML polled every ~0.5 ms by a SEPARATE process (a thread in this process competes with the launch loop for the
    GIL and stalls it); only the samples stamped inside the timed region are kept."""

    def __init__(self, index=0):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None

    def launch(self):
        import subprocess
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.first = self.proc.stdout.readline()      # "max <MHz>": the sampler is up
        except Exception:
            self.proc = None

    def start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        time.sleep(0.01)
        self.proc.terminate()
        out = self.proc.stdout.read().splitlines()
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        sm, rs, all_sm = [], 0, []
        for ln in out:
            f = ln.split()
            if len(f) != 3:
                continue
            try:
                ts, clk, r = float(f[0]), float(f[1]), int(f[2])
            except ValueError:
                continue
            all_sm.append(clk)
            if self.t0 <= ts <= self.t1:
                sm.append(clk)
                rs |= r
        try:
            mx = float(self.first.split()[1])
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else (float(np.median(all_sm)) if all_sm else None), "sm_max_mhz": mx,
                "reasons": sorted(k for k, b in bits.items() if rs & b), "samples": len(sm),
                "source": "nvml polled every ~0.5 ms by a separate process; samples stamped inside the timed region"}


# ----------------------------------------------------------------------------- CPU arm (oracle)
def cpu_time(run, steps, warmup, budget_s):
    for _ in range(max(1, warmup)):
        run()
    times, t_begin = [], time.perf_counter()
    while len(times) < steps and (not times or time.perf_counter() - t_begin < budget_s):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    return float(np.median(times)), len(times)


def cpu_arm(name, scale, hard, A, steps, warmup, budget_s):
    """Times the C oracle (all host threads) on the SAME matrix and x as the GPU arm. For c5 a step is a bounded
    sample: 8 of the 64 right-hand sides (LinearMaps applies a matrix right-hand side column by column), and
    the GB/s are those of the sample."""
    spec = workload_spec(name, scale, hard)
    threads = os.cpu_count() or 1
    op, nrhs = spec["op"], spec.get("nrhs", 1)
    ncols = min(nrhs, 8)
    nin = A.size[1] if op == "N" else A.size[0]
    x = host_x(nin, nrhs, A.dtype)
    orc = make_oracle(A, threads)
    if nrhs == 1:
        run = lambda: orc(x, op)
        sample = f"the full workload ({spec['desc']}), same x, {threads} threads"
    else:
        cols = [np.ascontiguousarray(x[:, j]) for j in range(ncols)]
        run = lambda: [orc(c, op) for c in cols]
        sample = f"the full matrix, {ncols} of the {nrhs} right-hand sides as a column loop, {threads} threads"
    sec, done = cpu_time(run, steps, warmup, budget_s)
    w = host_work(A, op, ncols if nrhs > 1 else 1)
    return {"gbs": w["bytes"] / sec / 1e9, "gflops": w["flops"] / sec / 1e9, "cores": threads, "sample": sample,
            "sec": sec, "steps_timed": done}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's own multiply (oracle port; the reference is Julia and
    cannot be installed here) on the host cores, same config as the GPU arm. Rank 0 only; libbsm_b200.so is never
    loaded by this arm."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    spec = workload_spec(args.workload, args.scale, args.hard)
    A, _ = build_workload(args.workload, args.scale, 0, 1, threads=min(os.cpu_count() or 8, 32), hard=args.hard)
    r = cpu_arm(args.workload, args.scale, args.hard, A, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["gbs"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": r["steps_timed"], "warmup": args.warmup, "ms_per_step": r["sec"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic", "gflops": r["gflops"],
        "config": {"workload": spec["desc"], "op": spec["op"],
                   "note": "CPU restatement of the reference schedule (oracle port, kind=port): Julia is not installable "
                           "in this image; " + r["sample"]},
        "cpu_baseline": {"value": r["gbs"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["gbs"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm
class Ctx:
    pass


def dgemm_peak(torch, dev):
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
    return 2 * 4096 ** 3 / (best * 1e-3) / 1e12


def run_workload(cx, name, primary):
    """One workload on the B200 arm. Returns the JSON line (rank 0) or None (other ranks); raises on parity failure
    only after the line was built (the caller prints it, then exits non-zero)."""
    import torch
    import torch.distributed as dist
    import bsm_b200 as B
    from bsm_b200 import _lib as L
    args, rank, world, local, dev = cx.args, cx.rank, cx.world, cx.local, cx.dev
    hard = args.hard if name == "c2" else ""
    spec = workload_spec(name, args.scale, hard)
    op = args.op or spec["op"]
    if args.nrhs and "nrhs" in spec:
        spec["nrhs"] = args.nrhs
        spec["desc"] = spec["desc"].replace("64 right-hand sides", f"{args.nrhs} right-hand sides")
    nrhs = spec.get("nrhs", 1)
    npdt = np.dtype(NPDT[spec["dtype"]])
    tdt = {"c128": torch.complex128, "f64": torch.float64, "f32": torch.float32}[spec["dtype"]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    host_threads = max(1, (os.cpu_count() or 8) // max(world, 1))
    A, rb = build_workload(name, args.scale, rank, world, threads=min(host_threads, 32), hard=hard)
    t_gen = time.time() - t0
    nin = A.size[1] if op == "N" else A.size[0]
    nout = A.size[0] if op == "N" else A.size[1]
    rhs_split = world > 1 and nrhs > 1          # c5 on N GPUs: 2-D grid of row slabs x column groups, no exchange
    GM = None

    t0 = time.time()
    SM = comm = None
    if rhs_split:
        from bsm_b200.dist import GridSplitMatrix
        grid = tuple(int(v) for v in args.rhs_grid.split("x")) if args.rhs_grid else None
        GM = GridSplitMatrix(A, cx.comm, nrhs, op, grid, variant=args.variant)
        D, own = GM.local, GM.out_rows
        work = host_work(A, op, nrhs)
    elif world > 1:
        from bsm_b200.dist import SlabMatrix
        comm = cx.comm
        if rb is None:        # generic partition of the full host matrix (every rank generated it)
            SM = SlabMatrix(A, comm, ops=(op,), variant=args.variant)
        else:                 # c2: every rank generated only its slab's blocks
            SM = SlabMatrix(A, comm, cuts=rb, variant=args.variant)
        D, own = SM.local, SM.own
        work = c2_whole_job_work(spec, "scattered" in hard) if rb is not None else host_work(A, op, nrhs)
    else:
        D = B.DeviceMatrix(A, device=local, variant=args.variant, plan_hints=args.plan_hints)
        own = (0, nout)
        work = host_work(A, op, nrhs)
    torch.cuda.synchronize()
    t_pack = time.time() - t0

    # ---- operands: the same x on every arm and every rank
    xh = host_x(nin, nrhs, npdt)
    j0, j1 = GM.cols if rhs_split else (0, nrhs)
    if nrhs == 1:
        x_full = torch.from_numpy(xh).to(dev)
        y_dev = torch.zeros(nout, dtype=tdt, device=dev)
    else:                     # this rank's column group, column-major
        x_full = torch.from_numpy(np.ascontiguousarray(xh[:, j0:j1].T)).to(dev).t()
        y_dev = torch.zeros((j1 - j0, nout), dtype=tdt, device=dev).t()
    peer = SM is not None and nrhs == 1 and args.xchg == "peer"
    if peer:
        xs = comm.alloc(nin, npdt)
        xs.copy_(x_full)
        x_full = xs
        step = lambda: SM.mul_peer(op, x_full, y_dev)
    elif SM is not None:
        step = lambda: SM.mul(op, x_full, y_dev)   # bsm_mul_dist: NCCL all-gather of the x slabs, then the slab multiply
    else:
        step = lambda: D.mul(op, x_full, y_dev)

    barrier()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and primary:
        sampler.launch()
    l2_resident = work["bytes"] <= 4 * 126e6 and not args.warm_l2
    # single GPU: the per-kernel event brackets are recorded IN the timed region. N > 1: a multiply lasts tens of
    # microseconds and three event records per step would leave gaps in the stream, so the kernel is timed in a
    # separate loop right after the timed region
    inline_prof = not l2_resident and world == 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if inline_prof:
        D.set_profiling(True)     # per-kernel event brackets inside bsm_mul, recorded IN the timed region
    barrier()
    if rank == 0 and primary:
        sampler.start()
    if not l2_resident:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
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
    clocks = sampler.stop() if (rank == 0 and primary) else None
    if inline_prof:
        k_ms, f_ms = D.profile()
        D.set_profiling(False)
    elif l2_resident:
        D.set_profiling(True)
        ks, fs = [], []
        for _ in range(max(5, min(args.steps, 20))):
            flush.zero_()
            step()
            a, b = D.profile()
            ks.append(a)
            fs.append(b)
        D.set_profiling(False)
        k_ms, f_ms = float(np.mean(ks)), float(np.mean(fs))
    else:
        D.set_profiling(True)
        barrier()
        for _ in range(args.steps):
            step()
        barrier()
        k_ms, f_ms = D.profile()
        D.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms_total, k_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, k_ms_max = float(t[0].item()), float(t[1].item())
    else:
        k_ms_max = k_ms
    ms_step = ms_total / args.steps

    # ---- parity at the benchmarked size: this rank's y slice against the C oracle on the same matrix and x
    y_dev.zero_()
    step()
    torch.cuda.synchronize()
    tol = TOL[spec["dtype"]]
    orc = make_oracle(A, host_threads)
    if nrhs == 1:
        yo = orc(xh, op)[own[0]:own[1]]
        yg = y_dev[own[0]:own[1]].cpu().numpy()
        if npdt == np.float32:      # the 1e-5 bound is against a Float64 evaluation
            from helpers import oracle_mul
            yo = oracle_mul(A, xh, op, threads=host_threads, f64=True)[own[0]:own[1]]
        err = float(np.linalg.norm(yg.astype(np.complex128) - yo) / max(np.linalg.norm(yo), 1e-300))
        checked = f"rows {own[0]}:{own[1]} of y" if world > 1 else "all of y"
    else:
        pc = list(range(j0, j1))[:: max(1, (j1 - j0) // 4)][:4]       # up to 4 of this rank's columns
        yg = y_dev.t().cpu().numpy()
        num = den = 0.0
        for j in pc:
            yo = orc(np.ascontiguousarray(xh[:, j]), op)[own[0]:own[1]]
            num += float(np.linalg.norm(yg[j - j0, own[0]:own[1]] - yo) ** 2)
            den += float(np.linalg.norm(yo) ** 2)
        err = float(np.sqrt(num / max(den, 1e-300)))
        checked = f"columns {pc} of Y" + (f", rows {own[0]}:{own[1]}" if own != (0, nout) else "")
    del orc
    if world > 1:
        t = torch.tensor([err], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err = float(t.item())
    parity = {"rel_err": err, "tol": tol, "op": op, "ok": bool(err <= tol),
              "checked": f"{checked} on every rank vs the C oracle (oracle/) on the same full-size matrix and x; max over ranks"}

    # ---- end to end through the host-pointer C-ABI call: pinned host x -> H2D -> multiply -> D2H -> host y
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    if SM is None:
        if nrhs == 1:
            xt, xa = pinned(xh)
            yt, ya = pinned(np.empty(nout, npdt))
            run_e2e = lambda: D.mul(op, xa, ya)
        else:
            xt, xa = pinned(xh[:, j0:j1].T)
            yt, ya = pinned(np.empty((j1 - j0, nout), npdt))
            run_e2e = lambda: D.mul(op, xa.T, ya.T)
        # whole-job bytes: on a grid the R ranks of a column each copy that column group of X in
        h2d, d2h = nin * nrhs * npdt.itemsize * (GM.grid[0] if GM is not None else 1), nout * nrhs * npdt.itemsize
    elif peer:
        xt, xa = pinned(xh[own[0]:own[1]])
        yt, ya = pinned(np.empty(own[1] - own[0], npdt))
        run_e2e = lambda: SM.mul_peer_host(op, xa, x_full, y_dev, ya)
        h2d, d2h = nin * npdt.itemsize, nout * npdt.itemsize
    else:
        xt, xa = pinned(xh[own[0]:own[1]])
        yt = torch.empty(own[1] - own[0], dtype=tdt).pin_memory()

        def run_e2e():
            x_full[own[0]:own[1]].copy_(xt, non_blocking=True)
            step()
            yt.copy_(y_dev[own[0]:own[1]], non_blocking=True)
            torch.cuda.synchronize()
        h2d, d2h = nin * npdt.itemsize, nout * npdt.itemsize
    for _ in range(2):
        run_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_e2e()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": work["bytes"] / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "call": "bsm_mul_host" if SM is None else ("bsm_mul_dist_peer_host" if peer else "torch copies + bsm_mul_dist")}

    # ---- solver loop through the operator: CG / COCG kept on the device, x sharded on N > 1
    solver = None
    if name == "c2" and primary and not args.no_solver and op == "N":
        solver = run_solver(cx, A, D, SM, own, host_threads)

    local_work = host_work(A, op, j1 - j0) if nrhs > 1 else host_work(A, op, 1)
    if GM is not None and GM.grid[0] > 1:
        lw = D.work(op, nrhs=j1 - j0)
        local_work = {"bytes": lw["bytes"] - lw["index_table_bytes"], "flops": lw["flops"]}
    if SM is not None:      # this rank's slab: what its kernel streams
        lw = D.work(op, nrhs=1)
        local_work = {"bytes": lw["bytes"] - lw["index_table_bytes"], "flops": lw["flops"]}
    stats = D.plan_stats(op)
    launches = D.launch_count(op)
    if peer:
        torch.cuda.synchronize()
        comm.free(xs)
    if rank != 0:
        return None, parity["ok"]

    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = local_work["bytes"] / (k_ms * 1e-3) / 1e9
    kernel_name = max(stats["bytes"], key=stats["bytes"].get)
    if kernel_name == "sym_fused_tma_kernel" and args.variant == 2:
        kernel_name = "sym_fused_kernel"
    tensor = None
    if nrhs >= 8 and stats["spmm"] and args.variant != 1:
        kernel_name = "spmm_dmma_kernel" if args.variant == 2 else stats.get("spmm_kernel", "spmm_dmma_kernel")
        dgemm_tf = dgemm_peak(torch, dev)
        ach_tf = local_work["flops"] / (k_ms * 1e-3) / 1e12
        tensor = {"bound": "tensor", "achieved": ach_tf, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach_tf / dgemm_tf,
                  "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (best of 5)"}
        launches = 1
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists() and args.scale == 1.0 and world == 1 and args.variant == 0 and not args.op and not hard:
        ent = json.loads(tfile.read_text()).get(name)
        if ent and ent["kernel"] == kernel_name:
            traffic = ent["dram_bytes_per_launch"]     # dram__bytes_read.sum + write.sum, ncu --set full (profiles/)
    if world == 1:
        par = "single GPU"
    elif rhs_split:
        R, C = GM.grid
        par = (f"{R} x {C} grid: {R} block-row slab(s) of A x {C} column group(s) of the {nrhs} right-hand sides "
               f"({nrhs // C} per rank), no exchange")
    elif peer:
        par = (f"block-row slabs x{world}, x sharded in peer-mapped arrays, fetched from its owners over NVLink inside "
               "the multiply kernels, both barriers inside the kernels (no collective, no extra launch)")
    else:
        par = f"block-row slabs x{world}, NCCL all-gather of x ({'sequential' if args.no_overlap else 'overlapped with the rank-local slices'})"
    line = {
        "metric": METRIC, "value": work["bytes"] / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
        "gflops": work["flops"] / (ms_step * 1e-3) / 1e9,
        "config": {"workload": spec["desc"], "op": op,
                   "l2": ("working set fits L2: L2 flushed (256 MB write) before every timed iteration, each multiply timed "
                          "by its own CUDA events") if l2_resident else
                         ("WARM L2 on request (--warm-l2): the working set fits L2 and is NOT flushed between iterations (the "
                          "solver-loop case)" if (args.warm_l2 and work["bytes"] <= 4 * 126e6) else
                          "working set larger than L2 (no flush needed)"),
                   "variant": {0: "auto", 1: "gather", 2: "fused", 3: "color", 4: "fused_tma"}[args.variant],
                   "parallelism": par, "algorithmic_bytes": work["bytes"], "flops": work["flops"],
                   "gen_s": round(t_gen, 1), "pack_s": round(t_pack, 1), "plan": stats},
        "parity": parity,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms, "kernel_ms_max_over_ranks": k_ms_max,
                     "finalize_ms": f_ms, "peak_source": peak_src, "bytes_per_launch": local_work["bytes"]},
        "e2e": e2e,
        "gpu_launches": int(args.steps * launches),
        "clocks": clocks,
    }
    if tensor is not None:
        line["roofline_tensor"] = tensor
    if solver is not None:
        line["solver"] = solver
    if world == 1 and primary and not args.no_cpu_baseline:
        r = cpu_arm(name, args.scale, hard, A, 3, 1, budget_s=20.0)
        line["cpu_baseline"] = {"value": r["gbs"], "unit": "GB/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"], "ms_per_multiply": r["sec"] * 1e3}
    return line, parity["ok"]


def run_solver(cx, A, D, SM, own, host_threads, shift=400.0):
    """COCG (bsm_cg / bsm_cg_dist) on the C2 operator made well conditioned by a diagonal shift (new VALUES, same
    structure: bsm_update_values re-uploads without re-planning). Reports wall time per iteration — multiply, fused
    vector kernels, the all-reduces of the dot products on N > 1, one host synchronisation per 8 iterations — and the
    residual of the returned x recomputed with the C oracle on this rank's rows."""
    import torch
    import torch.distributed as dist
    rank, world, dev = cx.rank, cx.world, cx.dev
    for d in A.diagonals:
        d[np.diag_indices(d.shape[0])] += shift
    t0 = time.perf_counter()
    D.update_values(A)
    torch.cuda.synchronize()
    t_update = time.perf_counter() - t0
    n = A.size[0]
    rng = np.random.default_rng(4321)
    bh = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
    b = torch.from_numpy(bh).to(dev)
    def solve(maxit):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if SM is None:
            x, iters, rel = D.cg(b, rtol=1e-10, maxit=maxit)
        else:
            x = torch.zeros_like(b)
            iters, rel = SM.cg(b, x, rtol=1e-10, maxit=maxit)
        torch.cuda.synchronize()
        return x, iters, rel, time.perf_counter() - t0

    # the set-up of a solve (work vectors; on N > 1 the collective allocation of the peer-mapped search direction) is
    # timed apart from the iterations: one solve cut off after 8 iterations, one to convergence
    solve(1)                      # warm-up: first-use costs (IPC mappings, allocator) stay out of both timed solves
    # the set-up time varies by tens of milliseconds between solves (collective allocation): best of three for each
    sec8 = sec = float("inf")
    for _ in range(3):
        _, it8, _, t8 = solve(8)
        x, iters, rel, tf = solve(200)
        sec8, sec = min(sec8, t8), min(sec, tf)
    sec_it = (sec - sec8) / max(iters - it8, 1) if iters > it8 else sec / max(iters, 1)
    xo = x.clone()
    if world > 1:          # every rank owns a slab of x: assemble the full vector for the oracle check
        xo[:own[0]] = 0
        xo[own[1]:] = 0
        xr = torch.view_as_real(xo)
        dist.all_reduce(xr, op=dist.ReduceOp.SUM)
        t = torch.tensor([sec, sec_it], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, sec_it = float(t[0].item()), float(t[1].item())
    res = make_oracle(A, host_threads)(xo.cpu().numpy(), "N")[own[0]:own[1]] - bh[own[0]:own[1]]
    num, den = float(np.linalg.norm(res) ** 2), float(np.linalg.norm(bh[own[0]:own[1]]) ** 2)
    if world > 1:
        t = torch.tensor([num, den], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        num, den = float(t[0].item()), float(t[1].item())
    return {"method": "COCG (bsm_cg%s), diagonal shift %g" % ("_dist" if SM is not None else "", shift), "iterations": iters,
            "relres_reported": rel, "relres_oracle": float(np.sqrt(num / den)), "ms_per_iteration": sec_it * 1e3,
            "solve_ms_end_to_end": sec * 1e3, "update_values_s": round(t_update, 2),
            "note": "ms_per_iteration = (solve to convergence - solve cut off after 8 iterations, best of 3 each) / (iterations - 8): "
                    "multiply + fused vector kernels + dot-product all-reduces + one host synchronisation per 8 iterations"}


def compact(line):
    keep = ("value", "unit", "n_gpus", "ms_per_step", "gflops", "dtype", "parity", "e2e", "gpu_launches")
    out = {k: line[k] for k in keep if k in line}
    out["workload"] = line["config"]["workload"]
    out["parallelism"] = line["config"]["parallelism"]
    out["roofline"] = {k: line["roofline"][k] for k in ("achieved", "peak", "frac", "kernel", "kernel_ms")}
    if "roofline_tensor" in line:
        out["roofline_tensor"] = {k: line["roofline_tensor"][k] for k in ("achieved", "peak", "frac")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (development only)")
    ap.add_argument("--hard", default="", choices=["", "permuted", "scattered", "permuted+scattered"],
                    help="c2 only: arbitrary unsorted index vectors and / or near leaves drawn from the whole leaf range")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--xchg", default="peer", choices=["peer", "nccl"],
                    help="N > 1, one right-hand side: peer = x read from its owners over NVLink inside the kernels "
                         "(no collective); nccl = all-gather of x, then multiply")
    ap.add_argument("--broadcasts", action="store_true", help="N > 1: grouped in-place broadcasts instead of the all-gather")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: all-gather, then multiply, on one stream")
    ap.add_argument("--op", default=None, choices=["N", "T", "C"], help="override the workload's op (development)")
    ap.add_argument("--plan-hints", type=int, default=0, help="bsm_options.plan_hints (development)")
    ap.add_argument("--warm-l2", action="store_true", help="small workloads: do NOT flush L2 between the timed iterations "
                                                           "(the matrix stays L2-resident, as inside a solver loop); the line says so")
    ap.add_argument("--rhs-grid", default="", help="c5 on N > 1: RxC grid of row slabs x column groups (default: the most "
                                                   "square grid with C >= R)")
    ap.add_argument("--nrhs", type=int, default=0, help="c5: number of right-hand sides (development; default 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="default workload only: skip the c3 / c5 companion results")
    ap.add_argument("--no-solver", action="store_true", help="c2: skip the CG / COCG solve through the operator")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    cx = Ctx()
    cx.args = args
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(cx.local)
    cx.dev = torch.device("cuda", cx.local)
    cx.comm = None
    if cx.world > 1:
        dist.init_process_group("nccl", device_id=cx.dev)
        from bsm_b200.dist import Comm
        cx.comm = Comm.from_torch(cx.local)     # libbsm_b200's own communicator (peer-mapped arrays, NCCL by dlopen)
        cx.comm.set_overlap(not args.no_overlap)
        cx.comm.set_collective(args.broadcasts)

    names = [args.workload]
    if args.workload == "c2" and not args.no_also and not args.hard and args.scale == 1.0 and not args.op and args.variant == 0:
        names += list(ALSO)
    lines, ok = [], True
    for i, nm in enumerate(names):
        line, good = run_workload(cx, nm, primary=(i == 0))
        ok = ok and good
        lines.append(line)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if cx.rank == 0:
        line = lines[0]
        if len(lines) > 1:
            line["also"] = [compact(l) for l in lines[1:]]
        print(json.dumps(line))
    if cx.world > 1:
        dist.destroy_process_group()
    if not ok:
        print("PARITY FAILURE: a GPU result differs from the oracle beyond the tolerance", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
