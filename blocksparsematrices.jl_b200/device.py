"""Device-resident operators: the Python counterpart of the Julia package extension.

DeviceMatrix(A) packs a host BlockSparseMatrix / SymmetricBlockMatrix / VBCRS into the HBM arena
through the C ABI (bsm_create_*) and exposes the multiply (bsm_mul / bsm_mul_host). x and y may be
NumPy arrays (host: copies happen inside bsm_mul_host) or torch CUDA tensors (device pointers, launched
on torch's current stream). There is no CPU fallback: without libbsm_b200.so or without a GPU every
product raises.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_int64, c_uint8, c_void_p

import numpy as np

from . import _lib as L
from .host import (AbstractBlockMatrix, BlockSparseMatrix, SymmetricBlockMatrix,
                   VariableBlockCompressedRowStorage, _Wrapped)

_DT = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64, np.dtype(np.complex128): L.C64}
_NP = {L.F32: np.dtype(np.float32), L.F64: np.dtype(np.float64), L.C64: np.dtype(np.complex128)}
_OPS = {"N": L.OP_N, "T": L.OP_T, "C": L.OP_C}


def _i64p(a):
    return a.ctypes.data_as(POINTER(c_int64))


def _pool(vecs):
    ptr = np.zeros(len(vecs) + 1, np.int64)
    if len(vecs):
        np.cumsum([len(v) for v in vecs], out=ptr[1:])
        pool = np.concatenate(vecs) if ptr[-1] > 0 else np.zeros(0, np.int64)
    else:
        pool = np.zeros(0, np.int64)
    return np.ascontiguousarray(pool, dtype=np.int64), ptr


def _device_blocks(tensors, shapes, dt, allow_transposed=False):
    """Device pointers of torch CUDA blocks (column-major: stride (1, m); or, where allowed, C-contiguous = the
    parent of a lazy transpose). Returns (pointer array, transposed flags)."""
    tdt = _torch_dtype(dt)
    ptrs = np.zeros(max(len(tensors), 1), np.uintp)
    tr = np.zeros(len(tensors), np.uint8)
    if len(tensors) != len(shapes):
        raise ValueError("device_blocks must list one tensor per block, in creation order")
    for i, (t, (m, n)) in enumerate(zip(tensors, shapes)):
        if not t.is_cuda or t.dtype != tdt or tuple(t.shape) != (m, n):
            raise TypeError(f"device block {i}: expected a CUDA tensor of shape {(m, n)} and dtype {tdt}")
        if m * n == 0 or (t.stride(0) == 1 and (n <= 1 or t.stride(1) == m)):
            pass
        elif allow_transposed and t.is_contiguous():
            tr[i] = 1
        else:
            raise ValueError(f"device block {i} must be column-major (stride (1, rows))")
        ptrs[i] = t.data_ptr()
    return ptrs, tr


def _colmajor(blocks, dt, allow_transposed=False):
    """Column-major views of the blocks (copies only where the memory is not already column-major of
    the right dtype). Returns (keepalive list, pointer array, m, n, transposed flags)."""
    keep, ptrs = [], np.zeros(max(len(blocks), 1), np.uintp)
    m = np.zeros(len(blocks), np.int64)
    n = np.zeros(len(blocks), np.int64)
    tr = np.zeros(len(blocks), np.uint8)
    for i, b in enumerate(blocks):
        b = np.asarray(b)
        if b.ndim != 2:
            raise ValueError(f"block {i} is not a matrix")
        m[i], n[i] = b.shape
        if b.dtype == dt and b.flags.f_contiguous:
            a = b
        elif allow_transposed and b.dtype == dt and b.flags.c_contiguous:
            a = b              # lazy transpose wrapper: memory holds the n x m column-major parent
            tr[i] = 1
        else:
            a = np.asfortranarray(b, dtype=dt)
        keep.append(a)
        ptrs[i] = a.__array_interface__["data"][0]
    return keep, ptrs, m, n, tr


class DeviceMatrix:
    """Opaque handle + size: what `B200(A)` of the Julia extension returns."""

    def __init__(self, A: AbstractBlockMatrix, device: int = -1, variant: int = L.VARIANT_AUTO,
                 own_rows=None, own_cols=None, plan_hints: int = 0, device_blocks=None):
        """device_blocks: optional list of torch CUDA tensors holding the block VALUES in HBM (creation order;
        symmetric: the diagonal blocks, then the off-diagonal blocks). The arena is then filled by a gather kernel on
        the device — nothing crosses PCIe; A only supplies the structure (its host blocks are not read)."""
        lib = L.lib()
        self.host = A
        self.size = A.size
        self.dtype = np.dtype(A.dtype)
        if self.dtype not in _DT:
            raise TypeError(f"unsupported element type {self.dtype}; supported: float32, float64, complex128")
        dt = _DT[self.dtype]
        opt = L.Options()
        lib.bsm_default_options(byref(opt))
        opt.device = device
        opt.variant = variant
        opt.plan_hints = plan_hints
        opt.blocks_on_device = 1 if device_blocks is not None else 0
        if own_rows is not None:
            opt.own_row_lo, opt.own_row_hi = int(own_rows[0]), int(own_rows[1])
        if own_cols is not None:
            opt.own_col_lo, opt.own_col_hi = int(own_cols[0]), int(own_cols[1])
        h = c_void_p()
        vp = lambda a: a.ctypes.data_as(POINTER(c_void_p))
        shp = lambda bs: [tuple(np.shape(b)) for b in bs]
        if isinstance(A, BlockSparseMatrix):
            keep, ptrs, m, n, _ = _colmajor(A.blocks, self.dtype)
            if device_blocks is not None:
                ptrs, _ = _device_blocks(device_blocks, shp(A.blocks), self.dtype)
            rp, rptr = _pool(A.rowindices)
            cp, cptr = _pool(A.colindices)
            L.check(lib.bsm_create_blocksparse(dt, A.size[0], A.size[1], len(A.blocks), vp(ptrs), _i64p(m),
                                               _i64p(n), _i64p(rp), _i64p(rptr), _i64p(cp), _i64p(cptr),
                                               byref(opt), byref(h)))
        elif isinstance(A, SymmetricBlockMatrix):
            keepd, dptrs, dm, dn, _ = _colmajor(A.diagonals, self.dtype)
            if np.any(dm != dn):
                raise ValueError("diagonal blocks must be square")
            keepo, optrs, om, on, _ = _colmajor(A.offdiagonals, self.dtype)
            if device_blocks is not None:
                nd = len(A.diagonals)
                dptrs, _ = _device_blocks(device_blocks[:nd], shp(A.diagonals), self.dtype)
                optrs, _ = _device_blocks(device_blocks[nd:], shp(A.offdiagonals), self.dtype)
            dp, dptr = _pool(A.diagonalindices)
            rp, rptr = _pool(A.rowindices)
            cp, cptr = _pool(A.colindices)
            L.check(lib.bsm_create_symmetric(dt, A.size[0], A.size[1], len(A.diagonals), vp(dptrs), _i64p(dm),
                                             _i64p(dp), _i64p(dptr), len(A.offdiagonals), vp(optrs),
                                             _i64p(om), _i64p(on), _i64p(rp), _i64p(rptr), _i64p(cp),
                                             _i64p(cptr), byref(opt), byref(h)))
        elif isinstance(A, VariableBlockCompressedRowStorage):
            keep, ptrs, m, n, tr = _colmajor(A.blocks, self.dtype, allow_transposed=True)
            if device_blocks is not None:
                ptrs, tr = _device_blocks(device_blocks, shp(A.blocks), self.dtype, allow_transposed=True)
            rowptr = np.ascontiguousarray(A.rowptr, np.int64)
            cs = np.ascontiguousarray(A.colindices, np.int64)
            rs = np.ascontiguousarray(A.rowindices, np.int64)
            L.check(lib.bsm_create_vbcrs(dt, A.size[0], A.size[1], len(rowptr) - 1, len(A.blocks), _i64p(rowptr),
                                         _i64p(cs), _i64p(rs), vp(ptrs), _i64p(m), _i64p(n),
                                         tr.ctypes.data_as(POINTER(c_uint8)), byref(opt), byref(h)))
        else:
            raise TypeError(f"cannot build a device matrix from {type(A).__name__}")
        self._h = h
        self.device = device

    def sparse(self, op="N"):
        """SparseArrays.sparse(op(A)) built on the device (bsm_sparse_build / _fetch) -> scipy CSC in Julia's
        canonical form (sorted rows, duplicates summed, explicit zeros kept)."""
        import scipy.sparse as sp
        nnz = c_int64(0)
        L.check(L.lib().bsm_sparse_build(self._h, _OPS[op], byref(nnz)))
        nc = self.size[1] if op == "N" else self.size[0]
        nr = self.size[0] if op == "N" else self.size[1]
        colptr = np.zeros(nc + 1, np.int64)
        rowval = np.zeros(nnz.value, np.int64)
        nzval = np.zeros(nnz.value, self.dtype)
        L.check(L.lib().bsm_sparse_fetch(self._h, _i64p(colptr), _i64p(rowval), nzval.ctypes.data_as(c_void_p)))
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(nr, nc))

    def update_values(self, A):
        """Re-uploads the block values of `A` (same structure as the matrix this handle was built from) into the
        arena, without re-planning."""
        if isinstance(A, SymmetricBlockMatrix):
            blocks = list(A.diagonals) + list(A.offdiagonals)
            allow_tr = False
        else:
            blocks = list(A.blocks)
            allow_tr = isinstance(A, VariableBlockCompressedRowStorage)
        keep, ptrs, m, n, tr = _colmajor(blocks, self.dtype, allow_transposed=allow_tr)
        L.check(L.lib().bsm_update_values(self._h, ptrs.ctypes.data_as(POINTER(c_void_p)), len(blocks)))
        self.host = A

    def update_values_dev(self, device_blocks):
        """New values from blocks that live in HBM (torch CUDA tensors, creation order): bsm_update_values_dev."""
        A = self.host
        if isinstance(A, SymmetricBlockMatrix):
            shapes = [tuple(np.shape(b)) for b in list(A.diagonals) + list(A.offdiagonals)]
            allow_tr = False
        else:
            shapes = [tuple(np.shape(b)) for b in A.blocks]
            allow_tr = isinstance(A, VariableBlockCompressedRowStorage)
        ptrs, _ = _device_blocks(device_blocks, shapes, self.dtype, allow_transposed=allow_tr)
        L.check(L.lib().bsm_update_values_dev(self._h, ptrs.ctypes.data_as(POINTER(c_void_p)), len(shapes)))

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None):
            L.lib().bsm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- LinearMap surface
    @property
    def shape(self):
        return self.size

    def adjoint(self):
        return DeviceAdjoint(self)

    def transpose(self):
        return DeviceTranspose(self)

    H = property(adjoint)
    T = property(transpose)

    def __mul__(self, x):
        return apply(self, "N", x)

    __matmul__ = __mul__

    # ---- queries

    def nnz(self) -> int:
        return int(L.lib().bsm_nnz(self._h))

    def stored_entries(self) -> int:
        return int(L.lib().bsm_stored_entries(self._h))

    def work(self, op="N", nrhs=1, beta_used=False):
        b, f, t = c_double(), c_double(), c_double()
        L.check(L.lib().bsm_work(self._h, _OPS[op], nrhs, int(beta_used), byref(b), byref(f), byref(t)))
        return {"bytes": b.value, "flops": f.value, "index_table_bytes": t.value}

    def launch_count(self, op="N") -> int:
        return int(L.lib().bsm_launch_count(self._h, _OPS[op]))

    def plan_stats(self, op="N") -> dict:
        """Work split of the plan used for `op` between the three multiply kernels."""
        out = np.zeros(12, np.int64)
        L.check(L.lib().bsm_plan_stats(self._h, _OPS[op], _i64p(out)))
        names = ("sym_fused_tma_kernel", "stream_warp_kernel", "gather_gemv_kernel")
        return {"slices": dict(zip(names, out[0:3].tolist())), "bytes": dict(zip(names, out[3:6].tolist())),
                "warp_items": int(out[6]), "warp_chunks": int(out[7]), "scratch_elems": int(out[8]),
                "finalized_rows": int(out[9]), "spmm": int(out[10]),
                "spmm_kernel": "spmm_tma_kernel" if out[11] else "spmm_dmma_kernel"}

    def set_profiling(self, on: bool):
        L.check(L.lib().bsm_set_profiling(self._h, int(on)))

    def profile(self):
        """(main kernel ms, gather/finalize kernel ms) of the most recent multiply."""
        a, b = c_double(), c_double()
        L.check(L.lib().bsm_get_profile(self._h, byref(a), byref(b)))
        return a.value, b.value

    def set_variant(self, variant: int):
        L.check(L.lib().bsm_set_variant(self._h, variant))

    def table(self, table: int, plan: int = 0) -> np.ndarray:
        """Export one packing table (bit-exact checks)."""
        lib = L.lib()
        cnt = lib.bsm_table_count(self._h, table, plan)
        if cnt < 0:
            raise L.BsmError(cnt, "unknown table")
        if table == L.TAB_ARENA:
            dt = self.dtype
        elif table in (L.TAB_BLOCK_OFF, L.TAB_SET_POOL_OFF, L.TAB_GATHER_PTR, L.TAB_GATHER_POS, L.TAB_GROUP_PTR,
                       L.TAB_CONTRIB_TOFF):
            dt = np.dtype(np.int64)
        elif table == L.TAB_CONTRIB:
            dt = CONTRIB_DTYPE
        elif table == L.TAB_SLICE:
            dt = SLICE_DTYPE
        elif table == L.TAB_WCHUNK:
            dt = WCHUNK_DTYPE
        else:
            dt = np.dtype(np.int32)
        out = np.zeros(cnt, dt)
        L.check(lib.bsm_table_copy(self._h, table, plan, out.ctypes.data_as(c_void_p), out.nbytes))
        return out

    # ---- solver loop on the device (bsm_cg)
    def cg(self, b, rtol=1e-10, maxit=200, hermitian=False, check_every=8, stream=None):
        """Solves A x = b by conjugate gradients kept on the device (hermitian=False: the unconjugated COCG form for
        complex symmetric operators; the same as CG for real ones). b: CUDA tensor of the operator's dtype.
        Returns (x, iterations, |r|/|b|)."""
        import torch
        if not _is_torch(b) or not b.is_cuda or b.dtype != _torch_dtype(self.dtype) or b.shape != (self.size[0],):
            raise TypeError("b must be a CUDA vector of the operator's dtype and length")
        b = b.contiguous()
        x = torch.empty_like(b)
        opt = L.CgOptions(rtol, maxit, int(hermitian), check_every)
        it, rr = c_int64(0), c_double(0.0)
        st = torch.cuda.current_stream(b.device).cuda_stream if stream is None else stream
        L.check(L.lib().bsm_cg(self._h, c_void_p(b.data_ptr()), c_void_p(x.data_ptr()), byref(opt), byref(it), byref(rr),
                               c_void_p(st)))
        return x, int(it.value), float(rr.value)

    # ---- multiply
    def mul(self, op, x, y=None, alpha=True, beta=False, stream=None):
        """y = alpha*op(A)*x + beta*y. beta is False (the bool) = Julia's strong zero."""
        lib = L.lib()
        nout = self.size[0] if op == "N" else self.size[1]
        nin = self.size[1] if op == "N" else self.size[0]
        beta_false = isinstance(beta, (bool, np.bool_)) and not beta
        a = np.array([alpha], dtype=self.dtype)
        b = np.array([0 if beta_false else beta], dtype=self.dtype)
        if _is_torch(x):
            import torch
            tdt = _torch_dtype(self.dtype)
            if not x.is_cuda:
                raise TypeError("torch inputs must be CUDA tensors (use NumPy arrays for host data)")
            if x.dtype != tdt:
                x = x.to(tdt)
            if x.shape[0] != nin:
                raise ValueError(f"DimensionMismatch: x has {x.shape[0]} rows, operator needs {nin}")
            nrhs = 1 if x.dim() == 1 else x.shape[1]
            if x.dim() == 1:
                xm = x if x.is_contiguous() else x.contiguous()
            elif x.stride(0) == 1 and x.stride(1) >= nin:
                xm = x                                               # column-major already (any leading dimension)
            else:
                xm = x.t().contiguous().t()                          # column-major storage
            if y is None:
                if not beta_false:
                    raise ValueError("beta needs an existing y")
                y = torch.empty((nrhs, nout), dtype=tdt, device=x.device).t() if x.dim() == 2 else \
                    torch.empty(nout, dtype=tdt, device=x.device)
            if y.dtype != tdt or y.shape[0] != nout:
                raise ValueError("DimensionMismatch: y does not match the operator")
            ldx = xm.stride(1) if xm.dim() == 2 else nin
            ldy = y.stride(1) if y.dim() == 2 else nout
            if y.dim() == 2 and (y.stride(0) != 1):
                raise ValueError("y must be column-major (stride(0) == 1)")
            st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
            L.check(lib.bsm_mul(self._h, _OPS[op], a.ctypes.data_as(c_void_p), b.ctypes.data_as(c_void_p),
                                int(beta_false), c_void_p(xm.data_ptr()), ldx, c_void_p(y.data_ptr()), ldy,
                                nrhs, c_void_p(st)))
            return y
        x = np.asarray(x)
        if x.shape[0] != nin:
            raise ValueError(f"DimensionMismatch: x has {x.shape[0]} rows, operator needs {nin}")
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        xm = np.asfortranarray(x, dtype=self.dtype) if x.ndim == 2 else np.ascontiguousarray(x, dtype=self.dtype)
        if y is None:
            if not beta_false:
                raise ValueError("beta needs an existing y")
            y = np.empty((nout, nrhs), self.dtype, order="F") if x.ndim == 2 else np.empty(nout, self.dtype)
        if y.dtype != self.dtype or y.shape[0] != nout or (y.ndim == 2 and not y.flags.f_contiguous) or \
                (y.ndim == 1 and not y.flags.c_contiguous):
            raise ValueError("DimensionMismatch: y must be a contiguous (column-major) array of the operator's dtype")
        L.check(lib.bsm_mul_host(self._h, _OPS[op], a.ctypes.data_as(c_void_p), b.ctypes.data_as(c_void_p),
                                 int(beta_false), xm.ctypes.data_as(c_void_p), nin, y.ctypes.data_as(c_void_p),
                                 nout, nrhs))
        return y


CONTRIB_DTYPE = np.dtype([("off", np.int64), ("m", np.int32), ("n", np.int32), ("in_set", np.int32),
                          ("form", np.int32), ("out_len", np.int32), ("block", np.int32)])
SLICE_DTYPE = np.dtype([("out_set", np.int32), ("r0", np.int32), ("r1", np.int32), ("c_begin", np.int32),
                        ("c_end", np.int32), ("flags", np.int32), ("scratch_off", np.int64)])
WCHUNK_DTYPE = np.dtype([("src16", np.uint32), ("bytes16", np.uint16), ("ncols", np.uint16), ("m", np.uint8),
                         ("flags", np.uint8), ("delta", np.uint8), ("seg_len", np.uint8), ("x_ref", np.int32),
                         ("out_col", np.uint8), ("src16_hi", np.uint8), ("smem16", np.uint16), ("lag", np.uint8),
                         ("reserved", np.uint8, (3,)), ("out", np.int64)])
assert CONTRIB_DTYPE.itemsize == 32 and SLICE_DTYPE.itemsize == 32 and WCHUNK_DTYPE.itemsize == 32


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _torch_dtype(dt):
    import torch
    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.complex128): torch.complex128}[np.dtype(dt)]


def _device_of(A) -> DeviceMatrix:
    if isinstance(A, DeviceMatrix):
        return A
    if isinstance(A, AbstractBlockMatrix):
        return A.device()
    raise TypeError(f"not a block matrix: {type(A).__name__}")


def apply(A, op, x):
    """y = op(A) * x with LinearMaps' promotion: complex blocks times real x promote x
    (the reference's own VBCRS tests do this, test/test_vbcrs.jl:34-35); real blocks times complex x
    are two real right-hand sides."""
    D = _device_of(A)
    if _is_torch(x):
        import torch
        if x.is_complex() and D.dtype.kind != "c":
            yr = D.mul(op, x.real.contiguous())
            yi = D.mul(op, x.imag.contiguous())
            return torch.complex(yr, yi)
        return D.mul(op, x)
    x = np.asarray(x)
    if np.iscomplexobj(x) and D.dtype.kind != "c":
        return D.mul(op, np.ascontiguousarray(x.real)) + 1j * D.mul(op, np.ascontiguousarray(x.imag))
    return D.mul(op, x)


def mul_into(y, A, op, x, alpha=True, beta=False):
    D = _device_of(A)
    if not _is_torch(x):
        x = np.asarray(x)
        if np.iscomplexobj(x) and D.dtype.kind != "c":
            raise TypeError("mul_: complex x with a real operator needs a complex y; use A * x")
    return D.mul(op, x, y, alpha, beta)


class DeviceAdjoint(_Wrapped):
    _op = "C"


class DeviceTranspose(_Wrapped):
    _op = "T"


def vbcrs_sort_device(rowstart, colstart):
    """The VBCRS sorting constructor on the device (bsm_vbcrs_sort_dev, src/vbcrs.jl:78-122): rowstart / colstart are
    1-based int64 CUDA tensors (one entry per block, unsorted). Returns (perm, rowptr, rowindices, colindices) as CUDA
    tensors: perm[k] = 0-based input index of the block that becomes block k (stable), rowptr 1-based with sentinel."""
    import torch
    nb = int(rowstart.numel())
    if rowstart.dtype != torch.int64 or colstart.dtype != torch.int64 or not rowstart.is_cuda or colstart.numel() != nb:
        raise TypeError("rowstart / colstart must be int64 CUDA tensors of equal length")
    dev = rowstart.device
    perm = torch.empty(nb, dtype=torch.int64, device=dev)
    rowptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    rowidx = torch.empty(max(nb, 1), dtype=torch.int64, device=dev)
    colidx = torch.empty(max(nb, 1), dtype=torch.int64, device=dev)
    nbrows = c_int64(0)
    st = torch.cuda.current_stream(dev).cuda_stream
    L.check(L.lib().bsm_vbcrs_sort_dev(nb, c_void_p(rowstart.contiguous().data_ptr()), c_void_p(colstart.contiguous().data_ptr()),
                                       c_void_p(perm.data_ptr()), c_void_p(rowptr.data_ptr()), c_void_p(rowidx.data_ptr()),
                                       c_void_p(colidx.data_ptr()), byref(nbrows), c_void_p(st)))
    nr = int(nbrows.value)
    return perm, rowptr[:nr + 1], rowidx[:nr], colidx[:nb]
