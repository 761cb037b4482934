"""Host-side mirror of the reference's operator / storage API for the multiply path.

Same names, argument meaning and error behaviour as BlockSparseMatrices.jl v0.3.1 (file:line relative
to /root/reference):

  BlockSparseMatrix(blocks, rowindices, colindices, size)                 src/blockmatrix.jl:26-109
  SymmetricBlockMatrix(diagonals, diagonalindices, offdiagonals,
                       rowindices, colindices, size)                      src/symmetricblockmatrix.jl:33-126
  VariableBlockCompressedRowStorage(matrices, rowindices, colindices, size)   src/vbcrs.jl:78-122
  VariableBlockCompressedRowStorage(bsm | sbm)                            src/vbcrs.jl:150-264
  A * x, A @ x, mul_(y, A, x[, α, β]), adjoint(A) / A.H, transpose(A) / A.T   (LinearMaps surface,
                                                                           src/abstractblockmatrix.jl:13-34)
  nnz, size, eltype, eachblockindex, block, rowindices, colindices,
  offdiagonal, diagonal, diagonalindices, rowcolvals, sparse

Indices are 1-based, blocks are column-major — exactly the data Julia holds. These types are plain
containers: every product runs on the GPU through libbsm_b200 (device.py); there is no CPU multiply
in this package.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

_SUPPORTED = (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.complex128))


def _as_index(v) -> np.ndarray:
    a = np.asarray(v)
    if a.dtype != np.int64:
        a = a.astype(np.int64)
    return np.ascontiguousarray(a)


def _common_dtype(mats) -> np.dtype:
    if len(mats) == 0:
        return np.dtype(np.float64)
    dt = np.result_type(*[m.dtype for m in mats[:64]]) if len(mats) > 1 else np.dtype(mats[0].dtype)
    return np.dtype(dt)


class AbstractBlockMatrix:
    """abstract type AbstractBlockMatrix{T} <: LinearMap{T} (src/abstractblockmatrix.jl:13)."""

    size: Tuple[int, int]
    _device = None

    # -- LinearMap surface
    @property
    def shape(self):
        return self.size

    @property
    def dtype(self):
        return self._dtype

    def adjoint(self):
        return AdjointMap(self)

    def transpose(self):
        return TransposeMap(self)

    @property
    def H(self):
        return AdjointMap(self)

    @property
    def T(self):
        return TransposeMap(self)

    def __mul__(self, x):
        return _apply(self, "N", x)

    __matmul__ = __mul__

    def __getitem__(self, key):
        return _getindex(self, key)

    # -- device mirror (the package-extension constructor; built on first use)
    def device(self, **kw):
        from .device import DeviceMatrix
        if self._device is None or kw:
            dev = DeviceMatrix(self, **kw)
            if kw:
                return dev
            self._device = dev
        return self._device


class _Wrapped:
    """LinearMaps.AdjointMap / TransposeMap: lazy wrappers with field .lmap."""
    _op = "N"

    def __init__(self, lmap):
        self.lmap = lmap

    @property
    def size(self):
        return self.lmap.size[::-1]

    shape = size

    @property
    def dtype(self):
        return self.lmap.dtype

    def __mul__(self, x):
        return _apply(self.lmap, self._op, x)

    __matmul__ = __mul__

    def __getitem__(self, key):
        return _getindex(self, key)


class AdjointMap(_Wrapped):
    _op = "C"

    def adjoint(self):
        return self.lmap

    H = property(adjoint)


class TransposeMap(_Wrapped):
    _op = "T"

    def transpose(self):
        return self.lmap

    T = property(transpose)


def adjoint(A):
    return A.adjoint() if hasattr(A, "adjoint") else AdjointMap(A)


def transpose(A):
    return A.transpose() if hasattr(A, "transpose") else TransposeMap(A)


def _unwrap(A):
    if isinstance(A, _Wrapped):
        return A.lmap, A._op
    return A, "N"


def _apply(A, op, x):
    from .device import apply
    return apply(A, op, x)


def _getindex(A, key):
    """A[i, j], A[:, j], A[i, :], A[:, :] as LinearMaps does it (getindex through products with unit vectors,
    the way the reference's tests materialise `b[:, :]`, test/test_blockmatrix.jl:38-49): the selected columns
    are ONE multi-RHS product on the GPU. Indices are 0-based here (Python), slices and integer arrays allowed."""
    if not isinstance(key, tuple) or len(key) != 2:
        raise IndexError("block matrices are indexed with two subscripts")
    nr, nc = A.size
    rows = np.arange(nr)[key[0]]
    cols = np.arange(nc)[key[1]]
    cvec = np.atleast_1d(cols)
    X = np.zeros((nc, len(cvec)), dtype=A.dtype, order="F")
    X[cvec, np.arange(len(cvec))] = 1
    Y = A * X if len(cvec) else np.zeros((nr, 0), A.dtype)
    out = Y[rows] if np.ndim(cols) else Y[rows, 0]
    return out


def mul_(y, A, x, alpha=True, beta=False):
    """LinearAlgebra.mul!(y, A, x[, α, β]); β === False is Julia's strong zero
    (src/abstractblockmatrix.jl:27-34)."""
    from .device import mul_into
    parent, op = _unwrap(A)
    return mul_into(y, parent, op, x, alpha, beta)


# ------------------------------------------------------------------------------- BlockSparseMatrix


class BlockSparseMatrix(AbstractBlockMatrix):
    """struct BlockSparseMatrix (src/blockmatrix.jl:26-34). `scheduler` / `coloringalgorithm` are
    accepted for signature compatibility and ignored: the GPU schedule is atomic-free and needs no
    colouring."""

    def __init__(self, blocks, rowindices, colindices, size, cols=None, scheduler=None,
                 coloringalgorithm=None):
        if cols is not None:            # BlockSparseMatrix(blocks, rowindices, colindices, rows, cols)
            size = (int(size), int(cols))
        self.blocks: List[np.ndarray] = list(blocks)
        self.rowindices: List[np.ndarray] = [_as_index(v) for v in rowindices]
        self.colindices: List[np.ndarray] = [_as_index(v) for v in colindices]
        self.size = (int(size[0]), int(size[1]))
        self._dtype = _common_dtype(self.blocks)
        self.scheduler = scheduler


def eachblockindex(A):
    A, _ = _unwrap(A)
    return range(1, len(A.blocks) + 1)      # eachindex(A.blocks), 1-based (src/blockmatrix.jl:124-134)


def block(A, i):
    """block(A, i) with the lazy adjoint/transpose of the wrappers (src/blockmatrix.jl:150-160)."""
    P, op = _unwrap(A)
    b = P.blocks[i - 1]
    return b if op == "N" else (b.T if op == "T" else b.conj().T)


def rowindices(A, i):
    P, op = _unwrap(A)                      # src/symmetricblockmatrix.jl:341-352
    return (P.rowindices if op == "N" else P.colindices)[i - 1]


def colindices(A, i):
    P, op = _unwrap(A)                      # src/symmetricblockmatrix.jl:354-365
    return (P.colindices if op == "N" else P.rowindices)[i - 1]


# ------------------------------------------------------------------------------- SymmetricBlockMatrix


class SymmetricBlockMatrix(AbstractBlockMatrix):
    """struct SymmetricBlockMatrix (src/symmetricblockmatrix.jl:33-44): diagonal blocks plus
    half-stored off-diagonal blocks."""

    def __init__(self, diagonals, diagonalindices, offdiagonals, rowindices, colindices, size,
                 cols=None, scheduler=None):
        if cols is not None:
            size = (int(size), int(cols))
        self.diagonals: List[np.ndarray] = list(diagonals)
        self.diagonalindices: List[np.ndarray] = [_as_index(v) for v in diagonalindices]
        self.offdiagonals: List[np.ndarray] = list(offdiagonals)
        self.rowindices: List[np.ndarray] = [_as_index(v) for v in rowindices]
        self.colindices: List[np.ndarray] = [_as_index(v) for v in colindices]
        self.size = (int(size[0]), int(size[1]))
        self._dtype = _common_dtype(self.diagonals if self.diagonals else self.offdiagonals)
        self.scheduler = scheduler


def eachoffdiagonalindex(A):
    A, _ = _unwrap(A)
    return range(1, len(A.offdiagonals) + 1)


def eachdiagonalindex(A):
    A, _ = _unwrap(A)
    return range(1, len(A.diagonals) + 1)


def offdiagonal(A, i):
    P, op = _unwrap(A)                      # src/symmetricblockmatrix.jl:197-233
    b = P.offdiagonals[i - 1]
    return b if op == "N" else (b.T if op == "T" else b.conj().T)


def diagonal(A, i):
    P, op = _unwrap(A)
    b = P.diagonals[i - 1]
    return b if op == "N" else (b.T if op == "T" else b.conj().T)


def diagonalindices(A, i):
    P, _ = _unwrap(A)                       # src/symmetricblockmatrix.jl:327-339
    return P.diagonalindices[i - 1]


# ------------------------------------------------------------------------------- VBCRS


class VariableBlockCompressedRowStorage(AbstractBlockMatrix):
    """struct VariableBlockCompressedRowStorage (src/vbcrs.jl:36-43).

    VariableBlockCompressedRowStorage(matrices, rowindices, colindices, size): sorting constructor
    (src/vbcrs.jl:78-122) — `rowindices` / `colindices` are the 1-based START row / column of every
    block. VariableBlockCompressedRowStorage(bsm) / (sbm): conversions (src/vbcrs.jl:150-199)."""

    def __init__(self, matrices, rowindices=None, colindices=None, matrixsize=None, scheduler=None):
        if isinstance(matrices, BlockSparseMatrix):
            b = matrices
            mats = b.blocks
            rs = np.fromiter((r[0] for r in b.rowindices), np.int64, len(mats))   # first(indices), :203-204
            cs = np.fromiter((c[0] for c in b.colindices), np.int64, len(mats))
            matrixsize = b.size
            scheduler = scheduler if scheduler is not None else b.scheduler
        elif isinstance(matrices, SymmetricBlockMatrix):
            s = matrices                                                            # :222-264
            mats = list(s.diagonals) + list(s.offdiagonals) + [o.T for o in s.offdiagonals]
            d0 = [d[0] for d in s.diagonalindices]
            r0 = [r[0] for r in s.rowindices]
            c0 = [c[0] for c in s.colindices]
            rs = np.asarray(d0 + r0 + c0, np.int64)
            cs = np.asarray(d0 + c0 + r0, np.int64)
            matrixsize = s.size
            scheduler = scheduler if scheduler is not None else s.scheduler
        else:
            mats = list(matrices)
            rs = _as_index(rowindices)
            cs = _as_index(colindices)
        n = len(mats)
        if n == 0:
            # matrices[1] / perm[1] throw BoundsError in the reference (src/vbcrs.jl:81, :104)
            raise IndexError("VariableBlockCompressedRowStorage needs at least one block")
        if len(rs) != n or len(cs) != n:
            raise ValueError("rowindices / colindices must hold one start index per block")
        # sortperm(1:n; by = i -> (rowindices[i], colindices[i])), stable (src/vbcrs.jl:84)
        perm = np.lexsort((cs, rs))
        rs_sorted = rs[perm]
        newrow = np.ones(n, dtype=bool)
        newrow[1:] = rs_sorted[1:] != rs_sorted[:-1]                                # :107-112
        starts = np.flatnonzero(newrow)
        self.blocks = [mats[i] for i in perm]
        self.rowptr = np.concatenate([starts + 1, [n + 1]]).astype(np.int64)        # 1-based + sentinel, :103, :117
        self.colindices = np.ascontiguousarray(cs[perm])
        self.rowindices = np.ascontiguousarray(rs_sorted[starts])
        self.size = (int(matrixsize[0]), int(matrixsize[1]))
        self._dtype = _common_dtype(self.blocks)
        self.scheduler = scheduler


# ------------------------------------------------------------------------------- nnz / sparse


def _blk_nnz(b) -> int:
    return int(b.shape[0]) * int(b.shape[1])


def nnz(A) -> int:
    """SparseArrays.nnz (src/blockmatrix.jl:208-223, src/symmetricblockmatrix.jl:367-384,
    src/vbcrs.jl:290-296)."""
    P, _ = _unwrap(A)
    if isinstance(P, SymmetricBlockMatrix):
        return 2 * sum(_blk_nnz(o) for o in P.offdiagonals) + sum(_blk_nnz(d) for d in P.diagonals)
    return sum(_blk_nnz(b) for b in P.blocks)


def size(A):
    return A.size


def eltype(A):
    return A.dtype


def _push_blocks(blocks, rows, cols, op):
    """_pushblocktoarrays! (src/sparse.jl:131-139) for a list of blocks, vectorised: row-major push
    order inside every block."""
    R, C, V = [], [], []
    for b, r, c in zip(blocks, rows, cols):
        bb = b if op == "N" else (b.T if op == "T" else b.conj().T)
        m, n = bb.shape
        R.append(np.repeat(r, n))
        C.append(np.tile(c, m))
        V.append(np.asarray(bb).reshape(-1, order="C"))
    return R, C, V


def rowcolvals(A):
    """rowcolvals(A) → (rows, cols, vals), 1-based COO triplets (src/sparse.jl:17-123). For
    BlockSparseMatrix / SymmetricBlockMatrix the push order is block order within every sweep (the
    serial colouring); VBCRS fills column-major per block."""
    P, op = _unwrap(A)
    if isinstance(P, VariableBlockCompressedRowStorage):
        if op != "N":
            raise TypeError("rowcolvals is not defined for wrapped VBCRS matrices in the reference")
        R, C, V = [], [], []
        for br in range(len(P.rowptr) - 1):
            r0 = int(P.rowindices[br])
            for bidx in range(P.rowptr[br] - 1, P.rowptr[br + 1] - 1):
                b = P.blocks[bidx]
                m, n = b.shape
                c0 = int(P.colindices[bidx])
                R.append(np.tile(np.arange(r0, r0 + m, dtype=np.int64), n))
                C.append(np.repeat(np.arange(c0, c0 + n, dtype=np.int64), m))
                V.append(np.asarray(b).reshape(-1, order="F"))
    elif isinstance(P, SymmetricBlockMatrix):
        ri, ci = (P.rowindices, P.colindices) if op == "N" else (P.colindices, P.rowindices)
        R, C, V = _push_blocks(P.offdiagonals, ri, ci, op)
        # transposed sweep: transpose(offdiagonal(A, b)) pushed at (colindices, rowindices), :63-73
        tb = [(o if op == "N" else (o.T if op == "T" else o.conj().T)).T for o in P.offdiagonals]
        R2, C2, V2 = _push_blocks(tb, ci, ri, "N")
        R3, C3, V3 = _push_blocks(P.diagonals, P.diagonalindices, P.diagonalindices, op)
        R, C, V = R + R2 + R3, C + C2 + C3, V + V2 + V3
    else:
        ri, ci = (P.rowindices, P.colindices) if op == "N" else (P.colindices, P.rowindices)
        R, C, V = _push_blocks(P.blocks, ri, ci, op)
    if not R:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, P.dtype)
    return np.concatenate(R), np.concatenate(C), np.concatenate(V).astype(P.dtype, copy=False)


def sparse(A):
    """SparseArrays.sparse(A) (src/sparse.jl:127-129) → scipy CSC with Julia's canonical form: row
    indices sorted inside each column, duplicates summed, explicit zeros kept."""
    import scipy.sparse as sp
    r, c, v = rowcolvals(A)
    S = sp.coo_matrix((v, (r - 1, c - 1)), shape=A.size).tocsc()
    S.sum_duplicates()
    S.sort_indices()
    return S
