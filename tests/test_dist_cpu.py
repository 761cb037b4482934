"""Host logic of the multi-GPU path on CPU: two gloo ranks cut a matrix into block-row slabs
(partition.slab_cuts / extract_slab), pack their slab with the slab-restricted packer (host-only handle),
execute their plan with the NumPy plan interpreter, exchange the x slices with an all-gather, and together
must reproduce the oracle. The NCCL communicator itself (bsm_dist_*) needs GPUs and is covered by the
-m gpu tests; here its entry points are only checked for presence and argument validation."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _worker(rank, world, port, kind, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bsm_b200 as B
    from bsm_b200 import _lib as L
    from bsm_b200 import generators as G
    from bsm_b200.partition import extract_slab, slab_cuts
    from helpers import oracle_mul
    from plan_interp import run_plan

    if kind == "sbm":
        A = G.symmetric_nearfield(seed=41, n=4000, k_near=3, leaf_min=10, leaf_max=60)
    elif kind == "vbcrs":
        A = G.vbcrs_variable(seed=42, n=6000)
    else:
        A = G.blocksparse_uniform(seed=43, n=3200, nblocks=500, bs=32)
    n = A.size[0]
    ops = ("N", "T", "C")
    cuts = slab_cuts(A, world)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    S = extract_slab(A, lo, hi, ops)
    D = S.device(device=L.DEVICE_NONE, own_rows=(lo, hi), own_cols=(lo, hi))
    rng = np.random.default_rng(5)                      # same x on every rank ...
    x_true = rng.standard_normal(n).astype(A.dtype)
    if np.dtype(A.dtype).kind == "c":
        x_true = x_true + 1j * rng.standard_normal(n)
    out = {}
    for op in ops:
        # ... but each rank only "owns" its slice: everything else arrives through the all-gather
        x = np.full(n, np.nan, x_true.dtype)
        x[lo:hi] = x_true[lo:hi]
        xr = torch.from_numpy(x.view(np.float64) if x.dtype.kind == "c" else x)
        k = 2 if x.dtype.kind == "c" else 1
        # uneven slabs: one in-place broadcast per rank, exactly what bsm_dist_allgather_rows groups in NCCL
        for r in range(world):
            dist.broadcast(xr[int(cuts[r]) * k:int(cuts[r + 1]) * k], src=r)
        xg = xr.numpy()
        xg = xg.view(np.complex128) if x.dtype.kind == "c" else xg
        y = np.zeros(n, xg.dtype)
        run_plan(S, D, op, xg, y=y, own=(lo, hi), in_own=(lo, hi))   # local slices must not need gathered x
        sl = D.table(L.TAB_SLICE, (0 if op == "N" else 1) + 2)
        nloc = int(np.sum((sl["flags"] & 16) == 0))
        assert 0 < nloc < len(sl), "a slab should have both rank-local and remote slices"
        ys = [None] * world
        dist.all_gather_object(ys, (lo, hi, y[lo:hi]))
        full = np.zeros(n, y.dtype)
        for a, b, part in ys:
            full[a:b] = part
        ref = oracle_mul(A, x_true, op)
        out[op] = float(np.linalg.norm(full - ref) / np.linalg.norm(ref))
    stored = D.stored_entries()
    tot = [None] * world
    dist.all_gather_object(tot, stored)
    if rank == 0:
        q.put((out, [int(c) for c in cuts], tot, A.device(device=L.DEVICE_NONE).stored_entries()))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["sbm", "vbcrs", "bsm"])
def test_two_rank_slabs_reproduce_the_oracle(kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"sbm": 0, "vbcrs": 1, "bsm": 2}[kind]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, cuts, stored, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(e < 1e-12 for e in err.values()), err
    assert cuts[0] == 0 and 0 < cuts[1] < cuts[2]
    # balanced: no rank holds much more than its share (boundary blocks are duplicated; a general
    # BlockSparseMatrix serving op N and op T keeps a block if its rows OR its columns touch the slab)
    assert max(stored) < (0.85 if kind == "bsm" else 0.65) * total + 1e4, (stored, total)


def test_slab_cuts_balance_and_alignment():
    import bsm_b200 as B
    from bsm_b200 import generators as G
    from bsm_b200.partition import row_costs, slab_cuts
    V = G.vbcrs_variable(seed=44, n=50000)
    cuts = slab_cuts(V, 8)
    assert cuts[0] == 0 and cuts[-1] == 50000 and np.all(np.diff(cuts) > 0)
    c = np.concatenate([[0], np.cumsum(row_costs(V))])
    share = np.diff(c[cuts]) / c[-1]
    assert share.max() < 1.15 / 8 and share.min() > 0.85 / 8
    starts = set(int(r) - 1 for r in V.rowindices)
    assert all(int(p) in starts for p in cuts[1:-1])      # cuts fall on block-row boundaries


def test_dist_entry_points_validate_arguments():
    from bsm_b200 import _lib as L
    lib = L.lib()
    assert lib.bsm_dist_allgather_rows(None, 1, None, 0, 1, None, None) == -1
    assert lib.bsm_mul_dist(None, None, 0, None, None, 1, None, 0, None, 0, 1, None, None) == -1
    assert lib.bsm_dist_destroy(None) == 0


def test_non_square_operator_slabs_rows_and_columns_cut_separately():
    """A non-square BlockSparseMatrix: rows (outputs of A x, inputs of A' x) and columns are partitioned separately;
    the slab-restricted plans of both "ranks", replayed by the plan interpreter, together give the full products."""
    import bsm_b200 as B
    from bsm_b200 import _lib as L
    from bsm_b200.partition import extract_slab, slab_cuts
    from helpers import oracle_mul
    from plan_interp import run_plan
    rng = np.random.default_rng(11)
    nr, nc = 900, 1700
    blocks, rows, cols = [], [], []
    for _ in range(120):
        m, n = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        r0, c0 = int(rng.integers(1, nr - m + 2)), int(rng.integers(1, nc - n + 2))
        blocks.append(np.asfortranarray(rng.standard_normal((m, n))))
        rows.append(np.arange(r0, r0 + m, dtype=np.int64))
        cols.append(np.arange(c0, c0 + n, dtype=np.int64))
    A = B.BlockSparseMatrix(blocks, rows, cols, (nr, nc))
    world = 2
    rcuts, ccuts = slab_cuts(A, world, "N"), slab_cuts(A, world, "T")
    assert rcuts[-1] == nr and ccuts[-1] == nc
    for op in ("N", "T"):
        nin, nout = (nc, nr) if op == "N" else (nr, nc)
        x = rng.standard_normal(nin)
        full = np.zeros(nout)
        for rank in range(world):
            rlo, rhi = int(rcuts[rank]), int(rcuts[rank + 1])
            clo, chi = int(ccuts[rank]), int(ccuts[rank + 1])
            S = extract_slab(A, rlo, rhi, ("N", "T"), cols=(clo, chi))
            D = S.device(device=L.DEVICE_NONE, own_rows=(rlo, rhi), own_cols=(clo, chi))
            own = (rlo, rhi) if op == "N" else (clo, chi)
            in_own = (clo, chi) if op == "N" else (rlo, rhi)
            y = np.zeros(nout)
            run_plan(S, D, op, x, y=y, own=own, in_own=in_own)
            assert not np.any(y[:own[0]]) and not np.any(y[own[1]:])
            full[own[0]:own[1]] = y[own[0]:own[1]]
        ref = oracle_mul(A, x, op)
        assert np.linalg.norm(full - ref) / np.linalg.norm(ref) < 1e-13, op


def test_rhs_grid_prefers_column_groups_and_divides_the_right_hand_sides():
    from bsm_b200.dist import rhs_grid
    assert [rhs_grid(n, 64) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    assert rhs_grid(8, 4) == (2, 4) and rhs_grid(4, 2) == (2, 2) and rhs_grid(6, 9) == (2, 3)
    with pytest.raises(ValueError):
        rhs_grid(4, 3)        # no grid whose column groups divide 3 right-hand sides
