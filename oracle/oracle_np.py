"""oracle_np.py — NumPy/SciPy restatement of BlockSparseMatrices.jl's multiply path and of the
structure functions either side of it, plus the ctypes binding of oracle/libbsm_oracle.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.

PARITY STATUS: "parity unpinned" for output vectors (no Julia here; the reference's tests hold no
golden outputs). Pinned instead by the reference's own test properties on its shipped fixture and
by C-oracle ≡ NumPy-oracle ≡ SciPy CSC product of the restated sparse(A) (tests/test_oracle.py).

Everything here works on raw containers (lists of 2-D arrays, lists of 1-based int64 index
vectors), exactly the fields the reference's structs hold:
  BlockSparseMatrix      /root/reference/src/blockmatrix.jl:26-34
  SymmetricBlockMatrix   /root/reference/src/symmetricblockmatrix.jl:33-44
  VBCRS                  /root/reference/src/vbcrs.jl:36-43
All indices stay 1-based, as in Julia. op ∈ {"N", "T", "C"} = A, transpose(A), A'.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

_HERE = Path(__file__).resolve().parent

# ----------------------------------------------------------------------------- containers


@dataclass
class OBSM:
    blocks: List[np.ndarray]
    rowindices: List[np.ndarray]
    colindices: List[np.ndarray]
    size: Tuple[int, int]


@dataclass
class OSBM:
    diagonals: List[np.ndarray]
    diagonalindices: List[np.ndarray]
    offdiagonals: List[np.ndarray]
    rowindices: List[np.ndarray]
    colindices: List[np.ndarray]
    size: Tuple[int, int]


@dataclass
class OVBCRS:
    blocks: List[np.ndarray]
    rowptr: np.ndarray      # 1-based, length nblockrows + 1, sentinel nblocks + 1
    colindices: np.ndarray  # 1-based start column per block
    rowindices: np.ndarray  # 1-based start row per block ROW
    size: Tuple[int, int]


def _opblock(b, op):
    # block(A', i) = adjoint(block), block(transpose(A), i) = transpose(block)
    # /root/reference/src/blockmatrix.jl:154-160
    if op == "N":
        return b
    if op == "T":
        return b.T
    return b.conj().T


def _scale(y, beta, beta_is_false):
    # y .*= β, `false` is a strong zero (/root/reference/src/blockmatrix.jl:231)
    if beta_is_false:
        y[...] = 0
    else:
        y *= beta


# ----------------------------------------------------------------------------- multiply (NumPy)


def mul_bsm(A: OBSM, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None):
    """/root/reference/src/blockmatrix.jl:225-247 (serial colour [1:nblocks], :91-92)."""
    nout = A.size[0] if op == "N" else A.size[1]
    dt = np.result_type(A.blocks[0].dtype if A.blocks else x.dtype, x.dtype)
    y = np.zeros(nout, dtype=dt) if y is None else y
    _scale(y, beta, beta_is_false)
    for b, r, c in zip(A.blocks, A.rowindices, A.colindices):
        ri, ci = (r, c) if op == "N" else (c, r)  # src/symmetricblockmatrix.jl:345-365
        np.add.at(y, ri - 1, alpha * (_opblock(b, op) @ x[ci - 1]))
    return y


def mul_sbm(A: OSBM, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None):
    """/root/reference/src/symmetricblockmatrix.jl:386-435, three sweeps."""
    dt = np.result_type(A.diagonals[0].dtype if A.diagonals else x.dtype, x.dtype)
    y = np.zeros(A.size[0], dtype=dt) if y is None else y
    _scale(y, beta, beta_is_false)
    for o, r, c in zip(A.offdiagonals, A.rowindices, A.colindices):   # :394-405
        ri, ci = (r, c) if op == "N" else (c, r)
        np.add.at(y, ri - 1, alpha * (_opblock(o, op) @ x[ci - 1]))
    for o, r, c in zip(A.offdiagonals, A.rowindices, A.colindices):   # :407-418
        ri, ci = (r, c) if op == "N" else (c, r)
        np.add.at(y, ci - 1, alpha * (_opblock(o, op).T @ x[ri - 1]))   # always transpose (:412)
    for d, ix in zip(A.diagonals, A.diagonalindices):                 # :420-432
        np.add.at(y, ix - 1, alpha * (_opblock(d, op) @ x[ix - 1]))
    return y


def mul_vbcrs(A: OVBCRS, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None):
    """/root/reference/src/vbcrs.jl:266-288 (forward) and :303-329 (transposed/adjoint)."""
    nout = A.size[0] if op == "N" else A.size[1]
    dt = np.result_type(A.blocks[0].dtype, x.dtype)
    y = np.zeros(nout, dtype=dt) if y is None else y
    _scale(y, beta, beta_is_false)
    for br in range(len(A.rowptr) - 1):
        for bidx in range(A.rowptr[br] - 1, A.rowptr[br + 1] - 1):
            blk = A.blocks[bidx]
            m, n = blk.shape
            c0 = A.colindices[bidx] - 1
            r0 = A.rowindices[br] - 1
            if op == "N":
                y[r0:r0 + m] += alpha * (blk @ x[c0:c0 + n])
            else:
                y[c0:c0 + n] += alpha * (_opblock(blk, op) @ x[r0:r0 + m])
    return y


# ----------------------------------------------------------------------------- structure


def nnz_bsm(A: OBSM) -> int:          # /root/reference/src/blockmatrix.jl:208-223
    return int(sum(b.size for b in A.blocks))


def nnz_sbm(A: OSBM) -> int:          # /root/reference/src/symmetricblockmatrix.jl:367-384
    return int(2 * sum(o.size for o in A.offdiagonals) + sum(d.size for d in A.diagonals))


def nnz_vbcrs(A: OVBCRS) -> int:      # /root/reference/src/vbcrs.jl:290-296
    return int(sum(b.size for b in A.blocks))


def vbcrs_from_blocks(matrices: Sequence[np.ndarray], rowstarts, colstarts, size) -> OVBCRS:
    """Sorting constructor, /root/reference/src/vbcrs.jl:78-122, literal loop restatement:
    stable sort on (rowstart, colstart) (:84); a new block row starts whenever the starting row
    changes (:89-94, :107-112); rowptr 1-based with sentinel (:103, :117)."""
    n = len(matrices)
    if n == 0:
        raise IndexError("VBCRS constructor needs at least one block (matrices[1], src/vbcrs.jl:81)")
    perm = sorted(range(n), key=lambda i: (int(rowstarts[i]), int(colstarts[i])))  # stable
    rowptr = [1]
    rowidx = [int(rowstarts[perm[0]])]
    blocks, cind = [], []
    for outidx, inidx in enumerate(perm, start=1):
        if int(rowstarts[inidx]) != rowidx[-1]:
            rowptr.append(outidx)
            rowidx.append(int(rowstarts[inidx]))
        blocks.append(matrices[inidx])
        cind.append(int(colstarts[inidx]))
    rowptr.append(n + 1)
    return OVBCRS(blocks, np.asarray(rowptr, np.int64), np.asarray(cind, np.int64),
                  np.asarray(rowidx, np.int64), tuple(size))


def vbcrs_from_bsm(A: OBSM) -> OVBCRS:
    """/root/reference/src/vbcrs.jl:150-160, functors :201-219: first(indices) only."""
    return vbcrs_from_blocks(A.blocks, [r[0] for r in A.rowindices],
                             [c[0] for c in A.colindices], A.size)


def vbcrs_from_sbm(A: OSBM) -> OVBCRS:
    """/root/reference/src/vbcrs.jl:189-199, functors :222-264: diagonals, off-diagonals, then
    transposed off-diagonals with row/col starts swapped."""
    mats = list(A.diagonals) + list(A.offdiagonals) + [o.T for o in A.offdiagonals]
    rs = ([d[0] for d in A.diagonalindices] + [r[0] for r in A.rowindices]
          + [c[0] for c in A.colindices])
    cs = ([d[0] for d in A.diagonalindices] + [c[0] for c in A.colindices]
          + [r[0] for r in A.rowindices])
    return vbcrs_from_blocks(mats, rs, cs, A.size)


def _push(rows, cols, vals, b, ri, ci):
    # _pushblocktoarrays!, /root/reference/src/sparse.jl:131-139 (row-major push order)
    m, n = b.shape
    rows.append(np.repeat(np.asarray(ri, np.int64), n))
    cols.append(np.tile(np.asarray(ci, np.int64), m))
    vals.append(np.asarray(b).reshape(-1, order="C"))


def rowcolvals_bsm(A: OBSM, op="N"):
    """/root/reference/src/sparse.jl:17-40 with the serial colouring (one colour, block order)."""
    rows, cols, vals = [], [], []
    for b, r, c in zip(A.blocks, A.rowindices, A.colindices):
        ri, ci = (r, c) if op == "N" else (c, r)
        _push(rows, cols, vals, _opblock(b, op), ri, ci)
    return _cat(rows, cols, vals, A.blocks)


def rowcolvals_sbm(A: OSBM, op="N"):
    """/root/reference/src/sparse.jl:42-91 (block order within each sweep = colour order for one
    colour; CSC canonicalisation makes the order irrelevant when no two blocks overlap)."""
    rows, cols, vals = [], [], []
    for o, r, c in zip(A.offdiagonals, A.rowindices, A.colindices):
        ri, ci = (r, c) if op == "N" else (c, r)
        _push(rows, cols, vals, _opblock(o, op), ri, ci)
    for o, r, c in zip(A.offdiagonals, A.rowindices, A.colindices):
        ri, ci = (r, c) if op == "N" else (c, r)
        _push(rows, cols, vals, _opblock(o, op).T, ci, ri)
    for d, ix in zip(A.diagonals, A.diagonalindices):
        _push(rows, cols, vals, _opblock(d, op), ix, ix)
    return _cat(rows, cols, vals, A.diagonals or A.offdiagonals)


def rowcolvals_vbcrs(A: OVBCRS):
    """/root/reference/src/sparse.jl:93-123 (column-major fill per block)."""
    rows, cols, vals = [], [], []
    for br in range(len(A.rowptr) - 1):
        for bidx in range(A.rowptr[br] - 1, A.rowptr[br + 1] - 1):
            blk = A.blocks[bidx]
            m, n = blk.shape
            r0, c0 = int(A.rowindices[br]), int(A.colindices[bidx])
            rows.append(np.tile(np.arange(r0, r0 + m, dtype=np.int64), n))
            cols.append(np.repeat(np.arange(c0, c0 + n, dtype=np.int64), m))
            vals.append(np.asarray(blk).reshape(-1, order="F"))
    return _cat(rows, cols, vals, A.blocks)


def _cat(rows, cols, vals, like):
    dt = like[0].dtype if like else np.float64
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, dt)
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals).astype(dt, copy=False)


def sparse_from_rcv(rows, cols, vals, size) -> sp.csc_matrix:
    """SparseArrays.sparse(I, J, V, m, n) (/root/reference/src/sparse.jl:127-129): CSC, row
    indices sorted inside each column, duplicates summed, explicit zeros kept."""
    A = sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=size).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    return A


def sparse_bsm(A: OBSM, op="N"):
    size = A.size if op == "N" else A.size[::-1]
    return sparse_from_rcv(*rowcolvals_bsm(A, op), size)


def sparse_sbm(A: OSBM, op="N"):
    return sparse_from_rcv(*rowcolvals_sbm(A, op), A.size)


def sparse_vbcrs(A: OVBCRS):
    return sparse_from_rcv(*rowcolvals_vbcrs(A), A.size)


# ----------------------------------------------------------------------------- C oracle binding

_DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex128): 2}
_OP = {"N": 0, "T": 1, "C": 2}
_lib = None


def c_lib():
    """Load oracle/libbsm_oracle.so (build it with `make -C oracle` if missing)."""
    global _lib
    if _lib is None:
        so = _HERE / "libbsm_oracle.so"
        if not so.exists():
            import subprocess
            subprocess.check_call(["make", "-C", str(_HERE)], stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(str(so))
        _lib.oracle_greedy_color.restype = ctypes.c_int64
    return _lib


def c_threads() -> int:
    return int(c_lib().oracle_max_threads())


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pool(vecs):
    ptr = np.zeros(len(vecs) + 1, np.int64)
    if len(vecs):
        ptr[1:] = np.cumsum([len(v) for v in vecs])
        pool = np.concatenate([np.asarray(v, np.int64) for v in vecs])
    else:
        pool = np.zeros(0, np.int64)
    return pool, ptr


def _blockptrs(blocks, dt):
    keep = [np.asfortranarray(b, dtype=dt) for b in blocks]
    arr = (ctypes.c_void_p * max(len(keep), 1))(*[b.ctypes.data for b in keep])
    return keep, arr


def greedy_colors(vecs):
    """First-fit greedy colouring → (ncolors, color_ptr, color_blocks) in CSR form (0-based block
    ids). Conflict = shared index (/root/reference/src/coloring.jl:45-61)."""
    pool, ptr = _pool(vecs)
    nb = len(vecs)
    col = np.zeros(max(nb, 1), np.int64)
    maxidx = int(pool.max()) if pool.size else 0
    nc = c_lib().oracle_greedy_color(ctypes.c_int64(nb), _p(pool), _p(ptr), ctypes.c_int64(maxidx), _p(col))
    assert nc >= 0
    col = col[:nb]
    order = np.argsort(col, kind="stable").astype(np.int64)
    cptr = np.zeros(nc + 1, np.int64)
    np.add.at(cptr, col + 1, 1)
    cptr = np.cumsum(cptr)
    return int(nc), cptr, order


def _serial_colors(nb):
    return 1, np.array([0, nb], np.int64), np.arange(nb, dtype=np.int64)


def _scalars(dt, alpha, beta):
    a = np.array([alpha], dtype=dt)
    b = np.array([beta], dtype=dt)
    return a, b


class CBsm:
    """Pre-marshalled BlockSparseMatrix for repeated (timed) C-oracle multiplies: block pointers, index pools
    and the colourings of both index families are built once."""

    def __init__(self, A: OBSM, threads=1):
        self.A = A
        self.dt = np.dtype(A.blocks[0].dtype)
        self.keep, self.bp = _blockptrs(A.blocks, self.dt)
        self.m = _i64([b.shape[0] for b in A.blocks])
        self.n = _i64([b.shape[1] for b in A.blocks])
        self.rpool, self.rptr = _pool(A.rowindices)
        self.cpool, self.cptr = _pool(A.colindices)
        self.threads = threads
        self._colors = {}

    def colors(self, op):
        key = "N" if op == "N" else "T"
        if key not in self._colors:
            if self.threads > 1:
                self._colors[key] = greedy_colors(self.A.rowindices if op == "N" else self.A.colindices)
            else:
                self._colors[key] = _serial_colors(len(self.A.blocks))
        return self._colors[key]

    def mul(self, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None):
        A, dt = self.A, self.dt
        x = np.ascontiguousarray(x, dtype=dt)
        nout = A.size[0] if op == "N" else A.size[1]
        y = np.zeros(nout, dt) if y is None else y
        nc, colptr, colblk = self.colors(op)
        a, b = _scalars(dt, alpha, beta)
        rc = c_lib().oracle_bsm_mul(_DT[dt], _OP[op], ctypes.c_int64(len(A.blocks)), self.bp, _p(self.m), _p(self.n),
                                    _p(self.rpool), _p(self.rptr), _p(self.cpool), _p(self.cptr), ctypes.c_int64(nc),
                                    _p(colptr), _p(colblk), _p(a), _p(b), int(beta_is_false), _p(x),
                                    _p(y), ctypes.c_int64(y.size), int(self.threads))
        assert rc == 0
        return y


def c_mul_bsm(A: OBSM, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None, threads=1):
    return CBsm(A, threads).mul(x, op, alpha, beta, beta_is_false, y)


class CSbm:
    """Pre-marshalled SymmetricBlockMatrix for repeated timed C-oracle multiplies."""

    def __init__(self, A: OSBM, threads=1):
        self.A = A
        self.dt = np.dtype((A.diagonals or A.offdiagonals)[0].dtype)
        self.keep_d, self.dp = _blockptrs(A.diagonals, self.dt)
        self.keep_o, self.op_ = _blockptrs(A.offdiagonals, self.dt)
        self.dsz = _i64([d.shape[0] for d in A.diagonals])
        self.om = _i64([o.shape[0] for o in A.offdiagonals])
        self.on = _i64([o.shape[1] for o in A.offdiagonals])
        self.dpool, self.dptr = _pool(A.diagonalindices)
        self.rpool, self.rptr = _pool(A.rowindices)
        self.cpool, self.cptr = _pool(A.colindices)
        self.threads = threads
        # the reference ALWAYS colours the three conflict graphs of an SBM
        # (/root/reference/src/symmetricblockmatrix.jl:104-110)
        self.crow = greedy_colors(A.rowindices)
        self.ccol = greedy_colors(A.colindices)
        self.cdiag = greedy_colors(A.diagonalindices)

    def mul(self, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None):
        A, dt = self.A, self.dt
        x = np.ascontiguousarray(x, dtype=dt)
        y = np.zeros(A.size[0], dt) if y is None else y
        a, b = _scalars(dt, alpha, beta)
        rc = c_lib().oracle_sbm_mul(
            _DT[dt], _OP[op], ctypes.c_int64(len(A.diagonals)), self.dp, _p(self.dsz), _p(self.dpool),
            _p(self.dptr), ctypes.c_int64(len(A.offdiagonals)), self.op_, _p(self.om), _p(self.on),
            _p(self.rpool), _p(self.rptr), _p(self.cpool), _p(self.cptr),
            ctypes.c_int64(self.crow[0]), _p(self.crow[1]), _p(self.crow[2]),
            ctypes.c_int64(self.ccol[0]), _p(self.ccol[1]), _p(self.ccol[2]),
            ctypes.c_int64(self.cdiag[0]), _p(self.cdiag[1]), _p(self.cdiag[2]),
            _p(a), _p(b), int(beta_is_false), _p(x), _p(y), ctypes.c_int64(y.size), int(self.threads))
        assert rc == 0
        return y


def c_mul_sbm(A: OSBM, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None, threads=1):
    return CSbm(A, threads).mul(x, op, alpha, beta, beta_is_false, y)


class CVbcrs:
    """Pre-marshalled VBCRS for repeated (timed) C-oracle multiplies."""

    def __init__(self, A: OVBCRS, threads=1):
        self.A = A
        self.dt = np.dtype(A.blocks[0].dtype)
        self.keep, self.bp = _blockptrs(A.blocks, self.dt)
        self.m = _i64([b.shape[0] for b in A.blocks])
        self.n = _i64([b.shape[1] for b in A.blocks])
        self.rowptr, self.cs, self.rs = _i64(A.rowptr), _i64(A.colindices), _i64(A.rowindices)
        self.threads = threads

    def mul(self, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None):
        A, dt = self.A, self.dt
        x = np.ascontiguousarray(x, dtype=dt)
        nout = A.size[0] if op == "N" else A.size[1]
        y = np.zeros(nout, dt) if y is None else y
        a, b = _scalars(dt, alpha, beta)
        rc = c_lib().oracle_vbcrs_mul(_DT[dt], _OP[op], ctypes.c_int64(len(self.rowptr) - 1), _p(self.rowptr),
                                      _p(self.cs), _p(self.rs), self.bp, _p(self.m), _p(self.n),
                                      _p(a), _p(b), int(beta_is_false), _p(x), _p(y),
                                      ctypes.c_int64(y.size), int(self.threads))
        assert rc == 0
        return y


def c_mul_vbcrs(A: OVBCRS, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None, threads=1):
    return CVbcrs(A, threads).mul(x, op, alpha, beta, beta_is_false, y)


# ----------------------------------------------------------------------------- golden fixture


def load_golden_sbm(name: str) -> OSBM:
    """tests/golden/symmetricblockexamples_<name>.npz → OSBM (see tests/golden/make_golden.py)."""
    z = np.load(_HERE.parent / "tests" / "golden" / f"symmetricblockexamples_{name}.npz")

    def mats(shapes, flat):
        out, p = [], 0
        for m, n in shapes:
            out.append(np.asfortranarray(flat[p:p + m * n].reshape((m, n), order="F")))
            p += m * n
        return out

    def vecs(pool, ptr):
        return [pool[ptr[i]:ptr[i + 1]].copy() for i in range(len(ptr) - 1)]

    n = int(z["n"])
    return OSBM(mats(z["diag_shapes"], z["diag_values"]), vecs(z["diag_idx"], z["diag_ptr"]),
                mats(z["off_shapes"], z["off_values"]), vecs(z["row_idx"], z["row_ptr"]),
                vecs(z["col_idx"], z["col_ptr"]), (n, n))


def sbm_to_bsm(A: OSBM) -> OBSM:
    """Expand a symmetric matrix into a general BlockSparseMatrix (diag + off + offᵀ) — used to get
    a real-structure BSM test case from the one shipped fixture."""
    blocks = list(A.diagonals) + list(A.offdiagonals) + [np.asfortranarray(o.T) for o in A.offdiagonals]
    rows = list(A.diagonalindices) + list(A.rowindices) + list(A.colindices)
    cols = list(A.diagonalindices) + list(A.colindices) + list(A.rowindices)
    return OBSM(blocks, rows, cols, A.size)
