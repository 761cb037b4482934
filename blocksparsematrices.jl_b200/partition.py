"""Block-row slab partition of a block matrix across the GPUs of one box.

Nothing in the reference does this (BlockSparseMatrices.jl is single-process): output rows are
independent given all of x, so the matrix is cut into `nparts` contiguous row slabs balanced by the bytes
streamed for their outputs; rank r packs only the blocks that contribute to its slab, holds a full-length
x (replicated every multiply by an all-gather of the slab slices over NCCL/NVLink) and writes only
y[cuts[r]:cuts[r+1]]. A half-stored symmetric block whose two uses land on two ranks is stored on both.

    cuts = slab_cuts(A, nparts)                      # 0-based row boundaries, len nparts+1
    S    = extract_slab(A, cuts[r], cuts[r+1])       # host container with the blocks rank r needs
    D    = DeviceMatrix(S, own_rows=(lo, hi), own_cols=(lo, hi))

The slab restriction itself (dropping / clipping contributions, straddling segments through the gather
lists) is done by the packer (bsm_options.own_row_* / own_col_*, csrc/pack.cpp).
"""
from __future__ import annotations

import numpy as np

from .host import BlockSparseMatrix, SymmetricBlockMatrix, VariableBlockCompressedRowStorage


def _uses(A, op="N"):
    """Yields (output index vector (1-based) or (start, len) range, entries) for every contribution of
    y = op(A) x."""
    if isinstance(A, SymmetricBlockMatrix):
        for d, idx in zip(A.diagonals, A.diagonalindices):
            yield idx, d.size
        for o, r, c in zip(A.offdiagonals, A.rowindices, A.colindices):
            yield r, o.size
            yield c, o.size
    elif isinstance(A, VariableBlockCompressedRowStorage):
        nbr = len(A.rowptr) - 1
        for br in range(nbr):
            for b in range(A.rowptr[br] - 1, A.rowptr[br + 1] - 1):
                m, n = A.blocks[b].shape
                if op == "N":
                    yield (int(A.rowindices[br]), m), m * n
                else:
                    yield (int(A.colindices[b]), n), m * n
    else:
        for b, r, c in zip(A.blocks, A.rowindices, A.colindices):
            yield (r if op == "N" else c), b.size


def row_costs(A, op="N") -> np.ndarray:
    """Stored entries streamed for every output row of y = op(A) x (a block's entries are spread evenly
    over the outputs it feeds): the weight the slabs balance."""
    nout = A.size[0] if op == "N" else A.size[1]
    cost = np.zeros(nout + 1, np.float64)
    for out, entries in _uses(A, op):
        if isinstance(out, tuple):
            s, ln = out
            if ln > 0:
                cost[s - 1] += entries / ln        # difference array over the range
                cost[s - 1 + ln] -= entries / ln
        elif len(out):
            idx = np.asarray(out) - 1
            lo, hi = int(idx.min()), int(idx.max()) + 1
            if hi - lo == len(idx):                # contiguous (possibly permuted inside): range update
                cost[lo] += entries / len(idx)
                cost[hi] -= entries / len(idx)
            else:
                np.add.at(cost, idx, entries / len(idx))
                np.add.at(cost, idx + 1, -entries / len(idx))
    return np.cumsum(cost)[:nout]


def _free_boundaries(A, op, nout) -> np.ndarray:
    """free[p] is True when no output index vector of a directly written segment straddles row p, so a
    cut at p keeps every such segment on one rank (no partial sums through the gather lists)."""
    cover = np.zeros(nout + 2, np.int64)
    if isinstance(A, SymmetricBlockMatrix):
        sets = list(A.diagonalindices) + list(A.rowindices)      # the leaf segments own their rows
    elif isinstance(A, VariableBlockCompressedRowStorage):
        sets = [o for o, _ in _uses(A, op)]
    else:
        sets = A.rowindices if op == "N" else A.colindices
    for out in sets:
        if isinstance(out, tuple):
            lo, hi = out[0] - 1, out[0] - 1 + out[1]
        elif len(out):
            lo, hi = int(np.min(out)) - 1, int(np.max(out))
        else:
            continue
        if hi - lo > 1:               # interior boundaries lo+1 .. hi-1 are straddled
            cover[lo + 1] += 1
            cover[hi] -= 1
    return np.cumsum(cover)[:nout + 1] == 0


def slab_cuts(A, nparts: int, op: str = "N") -> np.ndarray:
    """0-based boundaries of `nparts` contiguous output slabs with (nearly) equal row_costs, snapped to
    the nearest boundary no directly written segment straddles."""
    nout = A.size[0] if op == "N" else A.size[1]
    c = np.concatenate([[0.0], np.cumsum(row_costs(A, op))])
    free = np.flatnonzero(_free_boundaries(A, op, nout))
    cuts = [0]
    for p in range(1, nparts):
        ideal = int(np.searchsorted(c, c[-1] * p / nparts))
        if len(free):
            k = int(np.searchsorted(free, ideal))
            cand = [free[j] for j in (k - 1, k) if 0 <= j < len(free)]
            best = min(cand, key=lambda q: abs(q - ideal))
            if abs(best - ideal) <= max(64, nout // (8 * nparts)):
                ideal = int(best)
        cuts.append(min(max(ideal, cuts[-1]), nout))
    cuts.append(nout)
    return np.asarray(cuts, np.int64)


def _touches(idx, lo, hi) -> bool:
    if isinstance(idx, tuple):
        return idx[0] - 1 < hi and idx[0] - 1 + idx[1] > lo
    idx = np.asarray(idx)
    return bool(np.any((idx > lo) & (idx <= hi)))       # 1-based values against the 0-based [lo, hi)


def extract_slab(A, lo: int, hi: int, ops=("N",), cols=None):
    """Host container holding exactly the blocks that contribute to outputs [lo, hi) of op(A) x for the
    given ops (block data is shared with A, nothing is copied). Sizes and index vectors are unchanged, so
    x stays full length and the slab can be handed to DeviceMatrix(..., own_rows=(lo, hi), own_cols=(lo, hi)).
    cols = (clo, chi): the output range of the transposed ops when it differs from the row range (non-square
    operators: rows and columns are partitioned separately)."""
    clo, chi = (lo, hi) if cols is None else (int(cols[0]), int(cols[1]))
    if isinstance(A, SymmetricBlockMatrix):
        dk = [i for i, idx in enumerate(A.diagonalindices) if _touches(idx, lo, hi)]
        ok = [i for i, (r, c) in enumerate(zip(A.rowindices, A.colindices))
              if _touches(r, lo, hi) or _touches(c, lo, hi)]
        return SymmetricBlockMatrix([A.diagonals[i] for i in dk], [A.diagonalindices[i] for i in dk],
                                    [A.offdiagonals[i] for i in ok], [A.rowindices[i] for i in ok],
                                    [A.colindices[i] for i in ok], A.size)
    if isinstance(A, VariableBlockCompressedRowStorage):
        keep, rs, cs = [], [], []
        for br in range(len(A.rowptr) - 1):
            r0 = int(A.rowindices[br])
            for b in range(A.rowptr[br] - 1, A.rowptr[br + 1] - 1):
                m, n = A.blocks[b].shape
                c0 = int(A.colindices[b])
                if ("N" in ops and _touches((r0, m), lo, hi)) or \
                        (("T" in ops or "C" in ops) and _touches((c0, n), clo, chi)):
                    keep.append(A.blocks[b])
                    rs.append(r0)
                    cs.append(c0)
        if not keep:
            raise ValueError("empty slab: a VBCRS needs at least one block")
        return VariableBlockCompressedRowStorage(keep, rs, cs, A.size)
    k = [i for i, (r, c) in enumerate(zip(A.rowindices, A.colindices))
         if ("N" in ops and _touches(r, lo, hi)) or (("T" in ops or "C" in ops) and _touches(c, clo, chi))]
    return BlockSparseMatrix([A.blocks[i] for i in k], [A.rowindices[i] for i in k],
                             [A.colindices[i] for i in k], A.size)
