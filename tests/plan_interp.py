"""Test-only NumPy interpreter of the device plan exported by libbsm_b200 (bsm_table_copy).

It executes slices / contributions / gather lists exactly as the CUDA kernels are specified to
(blocksparsematrices.jl_b200/csrc/kernels.cuh), so the packer (pack.cpp) can be validated against the
oracle on machines without a GPU. It is NOT a product path: nothing in the package imports it.
"""
import numpy as np

from bsm_b200 import _lib as L
from bsm_b200 import BlockSparseMatrix, SymmetricBlockMatrix, VariableBlockCompressedRowStorage


def host_blocks(A):
    if isinstance(A, SymmetricBlockMatrix):
        return list(A.diagonals) + list(A.offdiagonals)
    return list(A.blocks)


def build_arena(A, D):
    off = D.table(L.TAB_BLOCK_OFF)
    blocks = host_blocks(A)
    total = int(off[-1] + blocks[-1].size + 64) if len(blocks) else 0
    arena = np.zeros(total, dtype=D.dtype)
    for b, o in zip(blocks, off):
        arena[o:o + b.size] = np.asarray(b).reshape(-1, order="F")
    return arena


class Sets:
    def __init__(self, D):
        self.len = D.table(L.TAB_SET_LEN)
        self.start = D.table(L.TAB_SET_START)
        self.poff = D.table(L.TAB_SET_POOL_OFF)
        self.pool = D.table(L.TAB_POOL)

    def idx(self, s, lo=0, hi=None):
        hi = self.len[s] if hi is None else hi
        if self.start[s] >= 0:
            return np.arange(self.start[s] + lo, self.start[s] + hi)
        return self.pool[self.poff[s] + lo:self.poff[s] + hi].astype(np.int64)


def run_plan(A, D, op, x, alpha=1.0, beta=0.0, beta_false=True, y=None, own=None, variant="auto"):
    plan = 0 if op == "N" else 1
    fused_exists = D.table(L.TAB_SLICE, 2).size > 0
    if variant == "fused" or (variant == "auto" and fused_exists):
        assert fused_exists
        plan += 2
    conj = op == "C"
    arena = build_arena(A, D)
    S = Sets(D)
    contrib = D.table(L.TAB_CONTRIB, plan)
    toff = D.table(L.TAB_CONTRIB_TOFF, plan)
    slices = D.table(L.TAB_SLICE, plan)
    grow = D.table(L.TAB_GATHER_ROWS, plan)
    gptr = D.table(L.TAB_GATHER_PTR, plan)
    gpos = D.table(L.TAB_GATHER_POS, plan)
    nout = A.size[0] if op == "N" else A.size[1]
    dt = np.result_type(D.dtype, x.dtype)
    y = np.zeros(nout, dt) if y is None else y
    nscratch = int(sum(int(s["r1"] - s["r0"]) for s in slices if not (s["flags"] & 1)))
    nscratch += int(sum(int(c["n"]) for c in contrib if c["form"] & 2))
    scratch = np.full(nscratch, np.nan, dt)
    seen_fused_first = True
    for k in range(1, len(slices)):      # fused slices come first
        assert not ((slices[k]["flags"] & 4) and not (slices[k - 1]["flags"] & 4))
    written = np.zeros(nout, np.int32)
    for s in slices:
        r0, r1 = int(s["r0"]), int(s["r1"])
        h = r1 - r0
        fused = bool(s["flags"] & 4)
        assert 0 < h <= (256 if fused else 128) and (r0 == 0 or not fused)
        acc = np.zeros(h, dt)
        for ci in range(s["c_begin"], s["c_end"]):
            c = contrib[ci]
            m, n = int(c["m"]), int(c["n"])
            B = arena[c["off"]:c["off"] + m * n].reshape((m, n), order="F")
            if conj:
                B = B.conj()
            if c["form"] & 2:
                assert fused and not (c["form"] & 1) and toff[ci] >= 0
                t0 = int(toff[ci])
                assert np.all(np.isnan(scratch[t0:t0 + n]))
                scratch[t0:t0 + n] = B.T @ x[S.idx(s["out_set"], 0, m)]
            else:
                assert toff[ci] == -1
            if (c["form"] & 1) == 0:
                hi = min(r1, int(c["out_len"]))
                if hi > r0:
                    acc[:hi - r0] += B[r0:hi, :] @ x[S.idx(c["in_set"], 0, n)]
            else:
                hi = min(r1, int(c["out_len"]))
                if hi > r0:
                    acc[:hi - r0] += B[:, r0:hi].T @ x[S.idx(c["in_set"], 0, m)]
            if s["flags"] & 2:   # vector loads promised: 16-byte alignment of every column
                v = 16 // D.dtype.itemsize
                assert m % v == 0 and r0 % v == 0 and (c["off"] * D.dtype.itemsize) % 128 == 0
        if s["flags"] & 1:
            rows = S.idx(s["out_set"], r0, r1)
            assert np.all(written[rows] == 0), "direct rows written twice"
            written[rows] += 1
            y[rows] = alpha * acc + (0 if beta_false else beta * y[rows])
        else:
            so = int(s["scratch_off"])
            assert np.all(np.isnan(scratch[so:so + h]))
            scratch[so:so + h] = acc
    assert not np.any(np.isnan(scratch))
    seen = np.zeros(nout, bool)
    for i, rr in enumerate(grow):
        row = int(rr) & 0x7FFFFFFF
        assert not seen[row]
        seen[row] = True
        ssum = scratch[gpos[gptr[i]:gptr[i + 1]]].sum() if gptr[i + 1] > gptr[i] else 0
        if int(rr) < 0:
            assert written[row] == 1
            y[row] = y[row] + alpha * ssum
        else:
            assert written[row] == 0
            y[row] = alpha * ssum + (0 if beta_false else beta * y[row])
    lo, hi = (0, nout) if own is None else own
    covered = (written > 0) | seen
    assert np.all(covered[lo:hi]), "an owned row is neither written directly nor finalised"
    assert not np.any(covered[:lo]) and not np.any(covered[hi:]), "a row outside the slab was written"
    return y
