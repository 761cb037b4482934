"""Seeded synthetic block patterns of the shapes BASELINE.json names.

Every generator returns a host container (host.py) whose blocks are column-major views into ONE
flat buffer, so a 12 GB matrix costs 12 GB of host memory, not 24. Sizes are parameters: the tests
use scaled-down instances of the same generators, bench.py the full ones.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .host import BlockSparseMatrix, SymmetricBlockMatrix, VariableBlockCompressedRowStorage


def fill_normal(buf: np.ndarray, seed: int, threads: int = 8, chunk: int = 1 << 24) -> None:
    """Fills a float32/float64 buffer with standard normals, chunk by chunk from independent child
    seeds (deterministic for a given (seed, chunk), independent of the thread count)."""
    n = buf.size
    nchunks = (n + chunk - 1) // chunk
    seeds = np.random.SeedSequence(seed).spawn(max(nchunks, 1))

    def work(i):
        lo, hi = i * chunk, min(n, (i + 1) * chunk)
        np.random.default_rng(seeds[i]).standard_normal(out=buf[lo:hi], dtype=buf.dtype)

    if nchunks <= 1 or threads <= 1:
        for i in range(nchunks):
            work(i)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(nchunks)))


def _values(total: int, dtype, seed: int, threads: int = 8) -> np.ndarray:
    """`total` standard-normal entries of `dtype` (complex: real and imaginary parts normal/sqrt(2))."""
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        raw = np.empty(2 * total, np.float64)
        fill_normal(raw, seed, threads)
        raw *= np.sqrt(0.5)
        return raw.view(np.complex128)
    raw = np.empty(total, dtype)
    fill_normal(raw, seed, threads)
    return raw


def _views(flat: np.ndarray, shapes) -> list:
    out, p = [], 0
    for m, n in shapes:
        out.append(flat[p:p + m * n].reshape((m, n), order="F"))
        p += m * n
    return out


def tiling(total: int, lo: int, hi: int, rng) -> np.ndarray:
    """Random tile sizes ~U{lo..hi} summing exactly to `total`; returns the 0-based boundaries."""
    est = int(total / ((lo + hi) / 2) * 1.2) + 16
    sizes = rng.integers(lo, hi + 1, est)
    cum = np.cumsum(sizes)
    while cum[-1] < total:
        sizes = np.concatenate([sizes, rng.integers(lo, hi + 1, est)])
        cum = np.cumsum(sizes)
    k = int(np.searchsorted(cum, total))
    bounds = np.concatenate([[0], cum[:k], [total]])
    return np.unique(bounds).astype(np.int64)


# ------------------------------------------------------------------------------------------- C1 / C5
def blocksparse_uniform(seed=1, n=10_000, nblocks=2_000, bs=32, dtype=np.float64, permuted=False, threads=8):
    """C1 (n=10,000, 2,000 blocks of 32x32) and C5 (n=1,000,000, ~200k blocks): equal square blocks on
    distinct random aligned slots of an n x n matrix. permuted=True renumbers rows and columns by random
    permutations, which turns every index vector into an arbitrary (non-contiguous, unsorted) one."""
    rng = np.random.default_rng(seed)
    slots = n // bs
    nblocks = min(nblocks, slots * slots)
    # distinct (block row, block column) slots
    picks = rng.choice(slots * slots, nblocks, replace=False) if slots * slots < (1 << 31) else \
        np.unique(rng.integers(0, slots * slots, int(nblocks * 1.1)))[:nblocks]
    picks = rng.permutation(picks)
    br, bc = picks // slots, picks % slots
    flat = _values(len(picks) * bs * bs, dtype, seed + 1000, threads)
    blocks = _views(flat, [(bs, bs)] * len(picks))
    base = np.arange(1, bs + 1, dtype=np.int64)
    if permuted:
        prow = rng.permutation(n).astype(np.int64) + 1
        pcol = rng.permutation(n).astype(np.int64) + 1
        rows = [prow[r * bs:(r + 1) * bs] for r in br]
        cols = [pcol[c * bs:(c + 1) * bs] for c in bc]
    else:
        rows = [base + r * bs for r in br]
        cols = [base + c * bs for c in bc]
    return BlockSparseMatrix(blocks, rows, cols, (n, n))


# ------------------------------------------------------------------------------------------- C2
class NearfieldStructure:
    """Structure (no values) of the C2 pattern: leaf boundaries and the near-leaf list of every leaf."""

    def __init__(self, seed, n, leaf_min, leaf_max, k_near, scattered=False):
        rng = np.random.default_rng(seed)
        self.seed, self.n, self.k_near = seed, n, k_near
        self.bounds = tiling(n, leaf_min, leaf_max, rng)
        self.nl = len(self.bounds) - 1
        self.sizes = np.diff(self.bounds)
        self.near = [np.zeros(0, np.int64)]
        for i in range(1, self.nl):
            # banded (default): near leaves among the 2*k_near leaves before i; scattered: anywhere below i, so the
            # column sets of neighbouring blocks share nothing (no x reuse, no locality for a slab partition)
            lo = 0 if scattered else max(0, i - 2 * k_near)
            k = min(i - lo, k_near)
            if scattered and i > 4 * k_near:
                pick = np.unique(rng.integers(0, i, 2 * k_near))[:k]
                while len(pick) < k:
                    pick = np.unique(np.concatenate([pick, rng.integers(0, i, k_near)]))[:k]
                self.near.append(np.sort(pick))
            else:
                self.near.append(np.sort(rng.choice(np.arange(lo, i), k, replace=False)))
        self.rng = rng

    def leaf_cost(self) -> np.ndarray:
        """Stored entries streamed for the outputs of every leaf (diagonal + forward off-diagonal +
        transposed contributions landing on the leaf): the weight the slab partition balances."""
        sz = self.sizes.astype(np.int64)
        cost = sz * sz
        for i in range(1, self.nl):
            cost[i] += sz[i] * sz[self.near[i]].sum()
            cost[self.near[i]] += sz[i] * sz[self.near[i]]
        return cost

    def partition(self, nparts: int) -> np.ndarray:
        """Leaf boundaries of `nparts` contiguous slabs balanced by leaf_cost (prefix-sum split)."""
        c = np.cumsum(self.leaf_cost())
        cuts = [0]
        for p in range(1, nparts):
            cuts.append(int(np.searchsorted(c, c[-1] * p / nparts)) + 1)
        cuts.append(self.nl)
        return np.asarray(cuts, np.int64)


def _fill_block(buf: np.ndarray, seed_key, dtype, symmetric=False):
    rng = np.random.default_rng(seed_key)
    if np.dtype(dtype).kind == "c":
        raw = buf.view(np.float64)
        rng.standard_normal(out=raw)
        raw *= np.sqrt(0.5)
    else:
        rng.standard_normal(out=buf, dtype=buf.dtype)
    if symmetric:
        m = int(round(np.sqrt(buf.size)))
        d = buf.reshape((m, m), order="F")
        d += d.T.copy()
        d *= 0.5


def symmetric_nearfield(seed=2, n=1_000_000, leaf_min=20, leaf_max=200, k_near=6, dtype=np.complex128,
                        permuted=False, threads=8, leaves=None, return_structure=False, scattered=False,
                        diag_shift=0.0):
    """C2: BEM near-field style SymmetricBlockMatrix. Leaves ~U{leaf_min..leaf_max} tile the n unknowns;
    every leaf has a (symmetrised) diagonal block; every leaf i >= 1 has ONE half-stored off-diagonal
    block whose rows are the leaf and whose columns are the union of min(i, k_near) lower-numbered
    near leaves drawn from the 2*k_near leaves before it (so column sets overlap between blocks, as in
    the reference's cuboid/sphere fixture). With the defaults: ~9.1k leaves, ~12.7 GB of ComplexF64.
    permuted=True applies a random renumbering of the unknowns (arbitrary index vectors, as the reference's
    fixture has); scattered=True draws the near leaves from ALL lower-numbered leaves instead of a band.
    diag_shift is added to the diagonal of every diagonal block (a shift of a few hundred makes the operator well
    conditioned — definite enough for the CG / COCG solver loop — without changing its structure).
    Block values come from per-block seeds, so `leaves=(lo, hi)` materialises exactly the blocks the
    slab owning leaves [lo, hi) needs (its diagonal blocks, its off-diagonal rows, and the blocks of
    other leaves whose column set touches the slab) with the same values as in the full matrix."""
    S = NearfieldStructure(seed, n, leaf_min, leaf_max, k_near, scattered=scattered)
    bounds, nl, sizes, near = S.bounds, S.nl, S.sizes, S.near
    perm = (S.rng.permutation(n).astype(np.int64) + 1) if permuted else None

    def idx(lo, hi):
        return perm[lo:hi] if permuted else np.arange(lo + 1, hi + 1, dtype=np.int64)

    lo_l, hi_l = (0, nl) if leaves is None else leaves
    dsel = list(range(lo_l, hi_l))
    osel = [i for i in range(1, nl)
            if (lo_l <= i < hi_l) or np.any((near[i] >= lo_l) & (near[i] < hi_l))]
    leaf_idx = {}

    def lidx(i):
        if i not in leaf_idx:
            leaf_idx[i] = idx(bounds[i], bounds[i + 1])
        return leaf_idx[i]

    dshapes = [(int(sizes[i]), int(sizes[i])) for i in dsel]
    oshapes = [(int(sizes[i]), int(sizes[near[i]].sum())) for i in osel]
    total = sum(m * k for m, k in dshapes) + sum(m * k for m, k in oshapes)
    flat = np.empty(total, np.dtype(dtype))
    nd = sum(m * k for m, k in dshapes)
    diag = _views(flat[:nd], dshapes)
    off = _views(flat[nd:], oshapes)
    jobs = [(d.reshape(-1, order="F"), [seed, 1, i], True) for d, i in zip(diag, dsel)] + \
           [(o.reshape(-1, order="F"), [seed, 2, i], False) for o, i in zip(off, osel)]

    def work(job):
        _fill_block(job[0], job[1], dtype, job[2])

    if threads > 1 and len(jobs) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, jobs, chunksize=16))
    else:
        for j in jobs:
            work(j)
    if diag_shift:
        for d in diag:
            d[np.diag_indices(d.shape[0])] += diag_shift
    A = SymmetricBlockMatrix(diag, [lidx(i) for i in dsel], off, [lidx(i) for i in osel],
                             [np.concatenate([lidx(j) for j in near[i]]) for i in osel], (n, n))
    return (A, S) if return_structure else A


# ------------------------------------------------------------------------------------------- C3
def vbcrs_variable(seed=3, n=4_000_000, tile_min=8, tile_max=64, extra_per_row=0.39, dtype=np.float64,
                   threads=8, as_blocksparse=False):
    """C3: variable-block CRS. One random tiling (tile sizes ~U{tile_min..tile_max}) is used for rows
    and columns; block row r holds the diagonal tile plus Poisson(extra_per_row) more blocks on
    distinct random column tiles (defaults: ~111k block rows, ~154k blocks, ~2e8 entries)."""
    rng = np.random.default_rng(seed)
    bounds = tiling(n, tile_min, tile_max, rng)
    nt = len(bounds) - 1
    sizes = np.diff(bounds)
    extra = rng.poisson(extra_per_row, nt)
    extra = np.minimum(extra, nt - 1)
    brow = np.repeat(np.arange(nt), 1 + extra)
    bcol = np.empty(len(brow), np.int64)
    pos = 0
    for r in range(nt):
        k = 1 + extra[r]
        bcol[pos] = r
        if k > 1:
            others = rng.choice(nt - 1, k - 1, replace=False)
            bcol[pos + 1:pos + k] = others + (others >= r)
        pos += k
    order = rng.permutation(len(brow))           # blocks are handed over unsorted
    brow, bcol = brow[order], bcol[order]
    shapes = list(zip(sizes[brow].tolist(), sizes[bcol].tolist()))
    flat = _values(int(np.sum(sizes[brow] * sizes[bcol])), dtype, seed + 1000, threads)
    blocks = _views(flat, shapes)
    rs = bounds[brow] + 1
    cs = bounds[bcol] + 1
    if as_blocksparse:
        rows = [np.arange(bounds[r] + 1, bounds[r + 1] + 1, dtype=np.int64) for r in brow]
        cols = [np.arange(bounds[c] + 1, bounds[c + 1] + 1, dtype=np.int64) for c in bcol]
        return BlockSparseMatrix(blocks, rows, cols, (n, n))
    return VariableBlockCompressedRowStorage(blocks, rs, cs, (n, n))


# ------------------------------------------------------------------------------------------- C4
def blocksparse_large(seed=4, grid=64, bs=1024, density=0.05, dtype=np.float32, threads=8):
    """C4: bs x bs dense blocks on `density` of a grid x grid block grid (defaults: 205 blocks of
    1024 x 1024 Float32, 860 MB)."""
    rng = np.random.default_rng(seed)
    nb = max(1, int(round(grid * grid * density)))
    picks = rng.permutation(rng.choice(grid * grid, nb, replace=False))
    br, bc = picks // grid, picks % grid
    flat = _values(nb * bs * bs, dtype, seed + 1000, threads)
    blocks = _views(flat, [(bs, bs)] * nb)
    base = np.arange(1, bs + 1, dtype=np.int64)
    rows = [base + r * bs for r in br]
    cols = [base + c * bs for c in bc]
    return BlockSparseMatrix(blocks, rows, cols, (grid * bs, grid * bs))
