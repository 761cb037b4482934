#!/usr/bin/env python
"""Timing probe of the sharded multiply (development tool, run under torchrun on N GPUs):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/peer_probe.py --workload c3

For the block-row slabs of the workload it times back-to-back multiplies (CUDA events, per rank and max over ranks) in
peer mode with the in-kernel barriers as shipped, with the entry wait / exit wait / both switched off
(bsm_dist_set_debug — x does not change between the multiplies, so the result stays valid), on the NCCL all-gather
path, and the rank-local kernel alone (plain bsm_mul on the slab handle with a fully replicated x)."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--sweep", default="", help="ENVVAR=v1,v2,...: rebuild the slab handle under every value of a tuning "
                                                "variable (BSM_TUNE_SPLIT_DIV, BSM_TUNE_WITEMS_PER_SLOT) and time it")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bsm_b200.dist import Comm, SlabMatrix
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm.from_torch(local)
    spec = bench.workload_spec(args.workload, args.scale)
    A, rb = bench.build_workload(args.workload, args.scale, rank, world, threads=max(1, (os.cpu_count() or 8) // world))
    op = spec["op"]
    SM = SlabMatrix(A, comm, ops=(op,)) if rb is None else SlabMatrix(A, comm, cuts=rb)
    n = A.size[0]
    npdt = np.dtype(bench.NPDT[spec["dtype"]])
    xh = bench.host_x(n, 1, npdt)
    x_rep = torch.from_numpy(xh).to(dev)
    xs = comm.alloc(n, npdt)
    xs.copy_(x_rep)
    y = torch.zeros_like(x_rep)

    def timed(fn, steps):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(o.item()), 4) for o in out]

    res = {"workload": spec["desc"], "n_gpus": world, "own_rows": list(map(int, SM.own))}
    if args.sweep:
        var, vals = args.sweep.split("=")
        res["sweep"] = {}
        for v in vals.split(","):
            os.environ[var] = v
            S2 = SlabMatrix(A, comm, ops=(op,)) if rb is None else SlabMatrix(A, comm, cuts=rb)
            res["sweep"][f"{var}={v}"] = {"peer": max(timed(lambda: S2.mul_peer(op, xs, y), args.steps)),
                                          "local_kernel_only": max(timed(lambda: S2.local.mul(op, x_rep, y), args.steps)),
                                          "slices": S2.local.plan_stats(op)["slices"], "warp_items": S2.local.plan_stats(op)["warp_items"]}
            del S2
        os.environ.pop(var, None)
    for name, flags in (("peer", 0), ("peer_no_entry_wait", 1), ("peer_no_exit_wait", 2), ("peer_no_waits", 3),
                        ("peer_release_signals", 16), ("peer_every_arrival_waits_sys", 4), ("peer_timed", 8)):
        comm.set_debug(flags)
        comm.debug_read()
        res[name] = timed(lambda: SM.mul_peer(op, xs, y), args.steps)
        if flags & 8:
            t = comm.debug_read()
            cnt = max(int(t[3]), 1)
            mine = {"entry_wait_us": t[0] / cnt / 1e3, "first_to_last_arrival_us": t[1] / cnt / 1e3,
                    "exit_wait_us": t[2] / cnt / 1e3, "multiplies": int(t[3])}
            allt = [None] * world
            dist.all_gather_object(allt, mine)
            res["barrier_timers_per_rank"] = allt
    comm.set_debug(0)
    res["nccl_allgather"] = timed(lambda: SM.mul(op, x_rep, y), args.steps)
    res["local_kernel_only"] = timed(lambda: SM.local.mul(op, x_rep, y), args.steps)
    D = SM.local
    D.set_profiling(True)
    for _ in range(10):
        SM.mul_peer(op, xs, y)
    k, f = D.profile()
    D.set_profiling(False)
    res["peer_kernel_ms_this_rank"] = [round(k, 4), round(f, 4)]
    res["plan"] = D.plan_stats(op)
    res["launches"] = D.launch_count(op)
    allres = [None] * world
    dist.all_gather_object(allres, {"rank": rank, "kernel_ms": res["peer_kernel_ms_this_rank"],
                                    "bytes": D.work(op)["bytes"], "slices": res["plan"]["slices"]})
    if rank == 0:
        res["per_rank"] = allres
        print(json.dumps(res))
    torch.cuda.synchronize()
    comm.free(xs)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
