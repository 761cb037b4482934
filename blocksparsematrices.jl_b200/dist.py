"""Single-box multi-GPU operator: block-row slabs + NCCL all-gather of x, through the C ABI
(bsm_dist_* / bsm_mul_dist in include/bsm_b200.h). One process per GPU (torchrun); torch.distributed is
used only for the rendezvous (shipping the 128-byte NCCL id), the data path is libbsm_b200's own
communicator. No CPU fallback."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_int, c_int64, c_void_p

import numpy as np

from . import _lib as L
from .device import DeviceMatrix, _DT, _OPS, _torch_dtype
from .partition import extract_slab, slab_cuts


class Comm:
    """NCCL communicator owned by libbsm_b200 (bsm_dist_init)."""

    def __init__(self, id_bytes: bytes, nranks: int, rank: int, device: int):
        self.nranks, self.rank, self.device = nranks, rank, device
        h = c_void_p()
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(id_bytes)
        L.check(L.lib().bsm_dist_init(buf, nranks, rank, device, byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = (ctypes.c_ubyte * 128)()
        L.check(L.lib().bsm_dist_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch(cls, device: int):
        """Rendezvous over an initialised torch.distributed process group (any backend)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], world, rank, device)

    def set_overlap(self, on: bool):
        L.check(L.lib().bsm_dist_set_overlap(self._h, int(on)))

    def set_collective(self, use_broadcasts: bool):
        L.check(L.lib().bsm_dist_set_collective(self._h, int(use_broadcasts)))

    def set_debug(self, flags: int):
        L.check(L.lib().bsm_dist_set_debug(self._h, int(flags)))

    def debug_read(self):
        out = np.zeros(8, np.int64)
        L.check(L.lib().bsm_dist_debug_read(self._h, out.ctypes.data_as(POINTER(c_int64))))
        return out

    def nccl_version(self) -> int:
        v = c_int(0)
        L.check(L.lib().bsm_dist_info(self._h, None, None, byref(v)))
        return int(v.value)

    def allgather_rows(self, x, cuts, stream=None):
        """In-place all-gather of the row slabs of a column-major CUDA tensor (vector or rows x nrhs)."""
        import torch
        dt = _DT[np.dtype({torch.float32: np.float32, torch.float64: np.float64,
                           torch.complex128: np.complex128}[x.dtype])]
        nrhs = 1 if x.dim() == 1 else x.shape[1]
        ldx = x.shape[0] if x.dim() == 1 else x.stride(1)
        if x.dim() == 2 and x.stride(0) != 1:
            raise ValueError("x must be column-major (stride(0) == 1)")
        c = np.ascontiguousarray(cuts, np.int64)
        st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        L.check(L.lib().bsm_dist_allgather_rows(self._h, dt, c_void_p(x.data_ptr()), ldx, nrhs,
                                                c.ctypes.data_as(POINTER(c_int64)), c_void_p(st)))

    def alloc(self, n: int, dtype):
        """Collective: a length-n CUDA tensor of `dtype` in memory that every rank of the box maps (bsm_dist_alloc,
        CUDA IPC). Keep x in such a tensor to use SlabMatrix.mul_peer; release it with free()."""
        import torch
        dt = np.dtype(dtype)
        p = c_void_p()
        L.check(L.lib().bsm_dist_alloc(self._h, n * dt.itemsize, byref(p)))

        class _Holder:   # __cuda_array_interface__ view of the raw allocation (owned by the communicator)
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (n,), "typestr": dt.str, "data": (p.value, False), "version": 3,
                                      "strides": None}
        t = torch.as_tensor(h, device=torch.device("cuda", self.device))
        t._bsm_holder = h
        return t

    def free(self, t):
        L.check(L.lib().bsm_dist_free(self._h, c_void_p(t.data_ptr())))

    def allreduce_max(self, t, stream=None):
        import torch
        assert t.dtype == torch.float64 and t.is_cuda
        st = torch.cuda.current_stream(t.device).cuda_stream if stream is None else stream
        L.check(L.lib().bsm_dist_allreduce_max_f64(self._h, c_void_p(t.data_ptr()), t.numel(), c_void_p(st)))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().bsm_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SlabMatrix:
    """Rank-local part of a block matrix cut into nnz-balanced block-row slabs.

    SlabMatrix(A, comm)                 A: the full host matrix (every rank holds it; blocks are shared,
                                        only this rank's slab is packed into HBM). Non-square operators: rows and
                                        columns are partitioned separately (cuts / col_cuts)
    SlabMatrix(S, comm, cuts=cuts)      S: a host container that already holds only this rank's blocks
    y = op(A) x: x is a full-length CUDA tensor of which this rank owns rows in_cuts[rank]:in_cuts[rank+1];
    mul() all-gathers it in place and writes y[out_cuts[rank]:out_cuts[rank+1]].
    """

    def __init__(self, A, comm: Comm, cuts=None, ops=("N",), variant=L.VARIANT_AUTO, col_cuts=None):
        self.comm = comm
        self.size = A.size
        square = A.size[0] == A.size[1]
        if cuts is None:
            # rows of y = A x and, for a non-square operator or when only transposed products are wanted, columns
            # (= rows of y = A' x) are partitioned separately, each balanced by the bytes streamed for its outputs
            cuts = slab_cuts(A, comm.nranks, "N" if ("N" in ops or not square) else "T")
            if col_cuts is None and not square:
                col_cuts = slab_cuts(A, comm.nranks, "T")
            cc = cuts if col_cuts is None else col_cuts
            A = extract_slab(A, int(cuts[comm.rank]), int(cuts[comm.rank + 1]), ops,
                             cols=(int(cc[comm.rank]), int(cc[comm.rank + 1])))
        elif not square and col_cuts is None:
            raise ValueError("a non-square operator needs col_cuts beside the row cuts")
        self.cuts = np.ascontiguousarray(cuts, np.int64)                    # rows: y of op N, x of op T / C
        self.col_cuts = self.cuts if col_cuts is None else np.ascontiguousarray(col_cuts, np.int64)
        lo, hi = int(self.cuts[comm.rank]), int(self.cuts[comm.rank + 1])
        clo, chi = int(self.col_cuts[comm.rank]), int(self.col_cuts[comm.rank + 1])
        self.own = (lo, hi)
        self.own_cols = (clo, chi)
        self.local = DeviceMatrix(A, device=comm.device, variant=variant, own_rows=(lo, hi), own_cols=(clo, chi))
        self.dtype = self.local.dtype
        # marshalled once: a multiply on 8 GPUs lasts tens of microseconds, the call must not cost as much
        self._one = np.array([1], dtype=self.dtype)
        self._zero = np.array([0], dtype=self.dtype)
        self._cuts_p = self.cuts.ctypes.data_as(POINTER(c_int64))
        self._col_cuts_p = self.col_cuts.ctypes.data_as(POINTER(c_int64))
        self._fn_peer = L.lib().bsm_mul_dist_peer

    def in_cuts_p(self, op):
        """x of op N is partitioned like the columns, x of op T / C like the rows."""
        return self._col_cuts_p if op == "N" else self._cuts_p

    def out_range(self, op):
        return self.own if op == "N" else self.own_cols

    def mul(self, op, x, y, alpha=True, beta=False, stream=None):
        import torch
        D = self.local
        beta_false = isinstance(beta, (bool, np.bool_)) and not beta
        a = np.array([alpha], dtype=D.dtype)
        b = np.array([0 if beta_false else beta], dtype=D.dtype)
        if x.dtype != _torch_dtype(D.dtype) or y.dtype != x.dtype or not x.is_cuda:
            raise TypeError("x and y must be CUDA tensors of the operator's dtype")
        nrhs = 1 if x.dim() == 1 else x.shape[1]
        ldx = x.shape[0] if x.dim() == 1 else x.stride(1)
        ldy = y.shape[0] if y.dim() == 1 else y.stride(1)
        st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        L.check(L.lib().bsm_mul_dist(self.comm._h, D._h, _OPS[op], a.ctypes.data_as(c_void_p),
                                     b.ctypes.data_as(c_void_p), int(beta_false), c_void_p(x.data_ptr()), ldx,
                                     c_void_p(y.data_ptr()), ldy, nrhs, self.in_cuts_p(op), c_void_p(st)))
        return y

    def mul_peer(self, op, x_shared, y, alpha=True, beta=False, stream=None):
        """y[own] = alpha*op(A_slab)*x + beta*y[own] with NO collective: x_shared is a tensor of Comm.alloc of which
        this rank has written its own slab; the kernels read every other entry from its owner over NVLink
        (bsm_mul_dist_peer: flag barrier, multiply, flag barrier)."""
        import torch
        D = self.local
        beta_false = beta is False or (isinstance(beta, np.bool_) and not beta)
        a = self._one if alpha is True else np.array([alpha], dtype=D.dtype)
        b = self._zero if beta_false else np.array([beta], dtype=D.dtype)
        if x_shared.dim() != 1 or x_shared.dtype != _torch_dtype(D.dtype) or y.dtype != x_shared.dtype:
            raise TypeError("x and y must be 1-D CUDA tensors of the operator's dtype")
        st = torch.cuda.current_stream(x_shared.device).cuda_stream if stream is None else stream
        rc = self._fn_peer(self.comm._h, D._h, _OPS[op], a.ctypes.data, b.ctypes.data, int(beta_false),
                           x_shared.data_ptr(), y.data_ptr(), self.in_cuts_p(op), st)
        if rc:
            L.check(rc)
        return y

    def mul_peer_host(self, op, x_host_slab, x_shared, y_dev, y_host_slab, alpha=True, beta=False, stream=None):
        """The collective multiply with this rank's x / y slabs in HOST memory (NumPy arrays, ideally pinned):
        bsm_mul_dist_peer_host — H2D of the x slab into x_shared, peer-mode multiply, D2H of the y slab, sync."""
        import torch
        D = self.local
        beta_false = isinstance(beta, (bool, np.bool_)) and not beta
        a = np.array([alpha], dtype=D.dtype)
        b = np.array([0 if beta_false else beta], dtype=D.dtype)
        lo, hi = self.out_range(op)
        ilo, ihi = self.own_cols if op == "N" else self.own
        if x_host_slab.dtype != D.dtype or y_host_slab.dtype != D.dtype or len(x_host_slab) != ihi - ilo or \
                len(y_host_slab) != hi - lo:
            raise ValueError("DimensionMismatch: host slabs must hold this rank's rows in the operator's dtype")
        st = torch.cuda.current_stream(x_shared.device).cuda_stream if stream is None else stream
        L.check(L.lib().bsm_mul_dist_peer_host(self.comm._h, D._h, _OPS[op], a.ctypes.data_as(c_void_p),
                                               b.ctypes.data_as(c_void_p), int(beta_false),
                                               x_host_slab.ctypes.data_as(c_void_p), c_void_p(x_shared.data_ptr()),
                                               c_void_p(y_dev.data_ptr()), y_host_slab.ctypes.data_as(c_void_p),
                                               self.in_cuts_p(op), lo, hi, c_void_p(st)))
        return y_host_slab

    def cg(self, b, x, rtol=1e-10, maxit=200, hermitian=False, check_every=8, stream=None):
        """Collective: conjugate gradients on the sharded operator (bsm_cg_dist). b, x: full-length CUDA tensors of which
        this rank reads / writes its own rows. Returns (iterations, |r|/|b|)."""
        import torch
        from ctypes import c_double
        D = self.local
        if b.dtype != _torch_dtype(D.dtype) or x.dtype != b.dtype or b.dim() != 1 or x.shape != b.shape:
            raise TypeError("b and x must be full-length CUDA vectors of the operator's dtype")
        opt = L.CgOptions(rtol, maxit, int(hermitian), check_every)
        it, rr = c_int64(0), c_double(0.0)
        st = torch.cuda.current_stream(b.device).cuda_stream if stream is None else stream
        L.check(L.lib().bsm_cg_dist(self.comm._h, D._h, c_void_p(b.data_ptr()), c_void_p(x.data_ptr()),
                                    self.cuts.ctypes.data_as(POINTER(c_int64)), byref(opt), byref(it), byref(rr),
                                    c_void_p(st)))
        return int(it.value), float(rr.value)


def rhs_grid(nranks: int, nrhs: int):
    """(row groups R, column groups C), R * C == nranks, for a matrix right-hand side on `nranks` GPUs: the most square
    grid with C >= R whose column groups divide the right-hand sides evenly. More column groups = less of X on every rank
    (and over PCIe); more row groups = less of A streamed per rank."""
    best = (1, nranks)
    for r in range(1, nranks + 1):
        if nranks % r == 0 and r <= nranks // r and nrhs % (nranks // r) == 0:
            best = (r, nranks // r)
    if nrhs % best[1]:
        raise ValueError("the right-hand sides must divide evenly among the column groups")
    return best


class GridSplitMatrix:
    """A matrix right-hand side on a 2-D grid of ranks, no exchange at all: rank i*C + j holds the block rows of slab i
    (balanced by streamed bytes, partition.slab_cuts) and multiplies them with column group j of X, writing rows
    out_rows of column group j of Y — one plain multi-RHS multiply per rank (spmm_tma_kernel). R = 1 is the pure column
    split (A replicated); row slabs alone (C = 1) would need all of X on every rank (512 MB per step for C5). The R ranks
    of a column hold the same column group of X."""

    def __init__(self, A, comm: Comm, nrhs: int, op="N", grid=None, variant=L.VARIANT_AUTO):
        self.comm, self.nrhs, self.op = comm, int(nrhs), op
        self.grid = R, C = tuple(grid) if grid is not None else rhs_grid(comm.nranks, self.nrhs)
        if R * C != comm.nranks or self.nrhs % C:
            raise ValueError("grid must have nranks entries and its column groups must divide nrhs")
        i, j = divmod(comm.rank, C)
        per = self.nrhs // C
        self.cols = (j * per, (j + 1) * per)
        nout = A.size[0] if op == "N" else A.size[1]
        self.size = A.size
        if R == 1:
            self.out_rows = (0, nout)
            self.local = DeviceMatrix(A, device=comm.device, variant=variant)
        else:
            cuts = slab_cuts(A, R, op)
            lo, hi = int(cuts[i]), int(cuts[i + 1])
            self.out_rows = (lo, hi)
            S = extract_slab(A, lo, hi, (op,), cols=(lo, hi))
            own = dict(own_rows=(lo, hi)) if op == "N" else dict(own_cols=(lo, hi))
            self.local = DeviceMatrix(S, device=comm.device, variant=variant, **own)
        self.dtype = self.local.dtype

    def mul(self, x_cols, y_cols=None, alpha=True, beta=False, stream=None):
        """x_cols: this rank's column group of X (rows x (cols[1] - cols[0]), column-major CUDA tensor or host array);
        rows out_rows of y_cols are written, the others are left alone."""
        if (1 if x_cols.ndim == 1 else x_cols.shape[1]) != self.cols[1] - self.cols[0]:
            raise ValueError("DimensionMismatch: x must hold this rank's column group")
        return self.local.mul(self.op, x_cols, y_cols, alpha, beta, stream)


class ColumnSplitMatrix(GridSplitMatrix):
    """The R = 1 grid: every rank holds A and its column group of X / Y."""

    def __init__(self, A, comm: Comm, nrhs: int, variant=L.VARIANT_AUTO):
        super().__init__(A, comm, nrhs, "N", (1, comm.nranks), variant)

    def mul(self, op, x_cols, y_cols=None, alpha=True, beta=False, stream=None):
        if (1 if x_cols.ndim == 1 else x_cols.shape[1]) != self.cols[1] - self.cols[0]:
            raise ValueError("DimensionMismatch: x must hold this rank's column group")
        return self.local.mul(op, x_cols, y_cols, alpha, beta, stream)
