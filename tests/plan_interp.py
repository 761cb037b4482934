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


def run_plan(A, D, op, x, alpha=1.0, beta=0.0, beta_false=True, y=None, own=None, variant="auto", in_own=None):
    """in_own = (lo, hi): slab handle under bsm_mul_dist — the slices NOT flagged remote (bit 4) run before the
    all-gather has delivered anything, so they are executed here with x poisoned (NaN) outside [lo, hi)."""
    plan = 0 if op == "N" else 1
    if variant == "color":
        return run_color_plan(A, D, op, x, alpha, beta, beta_false, y, plan + 4)
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
    cls = [(0 if (s["flags"] & 4) else 1 if (s["flags"] & 8) else 2, 1 if (s["flags"] & 16) else 0) for s in slices]
    assert cls == sorted(cls), "slices must be ordered by class (CTA-stream, warp-stream, gather), local before remote"
    x_full = x
    if in_own is not None:
        x_local = x.astype(np.result_type(x.dtype, np.float32), copy=True)
        x_local[:in_own[0]] = np.nan
        x_local[in_own[1]:] = np.nan
    else:
        assert own is not None or not any(s["flags"] & 16 for s in slices), "remote slices only exist in slab handles"
        x_local = x
    run_warp_stream(D, plan, arena, S, slices, x_full, x_local, y, scratch, written, conj, alpha, beta, beta_false, dt)
    for s in slices:
        x = x_full if (s["flags"] & 16) else x_local
        if s["flags"] & 8:
            continue            # executed from the chunk stream above
        r0, r1 = int(s["r0"]), int(s["r1"])
        h = r1 - r0
        fused = bool(s["flags"] & 4)
        assert 0 < h <= (256 if fused else 128)
        if fused and r0 > 0 or (fused and r1 < S.len[s["out_set"]]):
            # sub-range of a long segment: either T-form blocks of <= 1024 rows only (a run of whole block columns), or
            # N-form blocks only, cut into 256-row pieces whose columns all start 16-byte aligned
            forms = {int(contrib[ci]["form"]) & 3 for ci in range(s["c_begin"], s["c_end"])}
            assert forms in ({0}, {1}), forms
            if forms == {1}:
                assert all(contrib[ci]["m"] <= 1024 for ci in range(s["c_begin"], s["c_end"]))
            else:
                assert r0 % 256 == 0 and all((int(contrib[ci]["m"]) * D.dtype.itemsize) % 16 == 0
                                             for ci in range(s["c_begin"], s["c_end"]))
        acc = np.zeros(h, dt)
        for ci in range(s["c_begin"], s["c_end"]):
            c = contrib[ci]
            m, n = int(c["m"]), int(c["n"])
            if fused and (c["form"] & 1):
                assert m <= 1024       # x window of the CTA kernel
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


def run_warp_stream(D, plan, arena, S, slices, x_full, x_local, y, scratch, written, conj, alpha, beta, beta_false, dt):
    """Executes the bsm_wchunk stream exactly as stream_warp_kernel does: per work item, chunk by chunk,
    reading the arena BYTES the bulk copy would fetch."""
    chunks = D.table(L.TAB_WCHUNK, plan)
    iptr = D.table(L.TAB_WITEM_PTR, plan)
    wsl = [s for s in slices if s["flags"] & 8]
    if not wsl:
        assert chunks.size == 0 and iptr.size == 0
        return
    cta_mode = bool(np.any(chunks["flags"] & 64))
    assert iptr[0] == 0 and iptr[-1] == len(chunks)
    if cta_mode:
        # CTA-part mode: items 4c .. 4c+3 are the parts of segment c (part 0 never empty, later parts may be); every
        # part ends with a CtaPart record carrying the SAME output reference; the CTA sums the parts in order
        assert (len(iptr) - 1) == 4 * len(wsl) and np.all(np.diff(iptr) >= 0) and np.all(np.diff(iptr)[0::4] > 0)
        assert np.all((chunks["flags"][(chunks["flags"] & 16) != 0] & 64) != 0), "every segment end must be a part end"
    else:
        assert np.all(np.diff(iptr) > 0)
    raw = arena.view(np.uint8)
    isz = arena.dtype.itemsize
    nseg = 0
    seg = 0
    part_acc, part_info = None, None
    for it in range(len(iptr) - 1):
        acc = None
        check_ring_schedule(chunks[iptr[it]:iptr[it + 1]], isz)
        # a work item is all-local or all-remote (local items run while x is being gathered)
        if cta_mode:
            x = x_full if (wsl[it // 4]["flags"] & 16) else x_local
        else:
            nseg_item = int(np.sum((chunks[iptr[it]:iptr[it + 1]]["flags"] & 16) != 0))
            kinds = {bool(wsl[seg + k]["flags"] & 16) for k in range(nseg_item)}
            assert len(kinds) == 1, "a warp work item mixes local and remote segments"
            x = x_full if kinds.pop() else x_local
            seg += nseg_item
        for q in range(iptr[it], iptr[it + 1]):
            c = chunks[q]
            fl, m, nc, Lseg = int(c["flags"]), int(c["m"]), int(c["ncols"]), int(c["seg_len"])
            if fl & 8:
                assert acc is None
                acc = np.zeros(Lseg, dt)
            assert acc is not None, "work item must start at a segment boundary"
            assert 0 < m <= 64 and 0 < nc <= 64 and Lseg <= 64
            src = ((int(c["src16_hi"]) << 32) | int(c["src16"])) * 16
            nbytes = int(c["bytes16"]) * 16
            # a chunk and its x values take at most half the ring: the next chunk is in flight while this one is consumed
            assert nbytes + (((m if (fl & 1) else nc) * isz + 15) & ~15) + (16 if isz < 16 else 0) <= RING_BYTES // 2
            assert int(c["delta"]) + m * nc * isz <= nbytes
            buf = raw[src:src + nbytes]
            Bc = buf[int(c["delta"]):int(c["delta"]) + m * nc * isz].view(arena.dtype).reshape((m, nc), order="F")
            if conj:
                Bc = Bc.conj()
            cnt = m if (fl & 1) else nc
            pos = np.arange(c["x_ref"], c["x_ref"] + cnt)
            xi = x[S.pool[pos]] if (fl & 2) else x[pos]
            if fl & 1:
                oc = int(c["out_col"])
                acc[oc:oc + nc] += Bc.T @ xi
            else:
                acc[:m] += Bc @ xi
            if (fl & 16) and (fl & 64):
                info = (fl & (4 | 32), int(c["out"]), Lseg)
                assert q == iptr[it + 1] - 1, "a part holds exactly one (partial) segment"
                if part_acc is None:
                    assert it % 4 == 0
                    part_acc, part_info = acc.copy(), info
                else:
                    assert info == part_info, "the parts of a segment must agree on its output"
                    part_acc += acc
                acc = None
            elif fl & 16:
                o = int(c["out"])
                if fl & 32:
                    rows = S.pool[o:o + Lseg].astype(np.int64) if (fl & 4) else np.arange(o, o + Lseg)
                    assert np.all(written[rows] == 0), "direct rows written twice"
                    written[rows] += 1
                    y[rows] = alpha * acc + (0 if beta_false else beta * y[rows])
                else:
                    assert np.all(np.isnan(scratch[o:o + Lseg]))
                    scratch[o:o + Lseg] = acc
                acc = None
                nseg += 1
        assert acc is None, "work item must end at a segment boundary"
        if cta_mode and it % 4 == 3:          # the CTA's single write
            pfl, o, Lseg = part_info
            if pfl & 32:
                rows = S.pool[o:o + Lseg].astype(np.int64) if (pfl & 4) else np.arange(o, o + Lseg)
                assert np.all(written[rows] == 0), "direct rows written twice"
                written[rows] += 1
                y[rows] = alpha * part_acc + (0 if beta_false else beta * y[rows])
            else:
                assert np.all(np.isnan(scratch[o:o + Lseg]))
                scratch[o:o + Lseg] = part_acc
            part_acc, part_info = None, None
            nseg += 1
    assert nseg == len(wsl)
    for s in wsl:
        assert s["r0"] == 0 and s["r1"] == S.len[s["out_set"]] <= 64


RING_BYTES, RING_SLOTS = 11264, 16


def check_ring_schedule(ch, isz):
    """Replays the issue / consume order of one warp work item: chunk i is issued as soon as i - lag
    chunks have been consumed; its ring region must not overlap any chunk still live, and at most
    RING_SLOTS chunks (one mbarrier each) may be live."""
    n = len(ch)
    foot = ch["bytes16"].astype(np.int64) * 16
    cnt = np.where(ch["flags"] & 1, ch["m"], ch["ncols"]).astype(np.int64)
    foot += (cnt * isz + 15) // 16 * 16 + (16 if isz < 16 else 0)     # x values: the enclosing 16-byte aligned range
    off = ch["smem16"].astype(np.int64) * 16
    assert np.all(off + foot <= RING_BYTES)
    lag = ch["lag"].astype(np.int64)
    assert np.all(lag <= np.arange(n)) and np.all(lag < RING_SLOTS)
    ii, live = 0, []
    for done in range(n + 1):          # state after `done` chunks have been consumed
        live = [k for k in live if k >= done]
        while ii < n and ii - lag[ii] <= done:
            for k in live:
                assert off[ii] + foot[ii] <= off[k] or off[k] + foot[k] <= off[ii], "ring regions overlap"
            live.append(ii)
            assert len(live) <= RING_SLOTS
            ii += 1
        assert done == n or (live and live[0] == done), "the chunk to consume was never issued"
    assert ii == n


def run_color_plan(A, D, op, x, alpha, beta, beta_false, y, plan):
    """Colour-ordered variant: y <- beta*y, then launch by launch; inside a launch no row may be touched
    twice (that is what the colouring guarantees), every slice accumulates straight into y."""
    arena = build_arena(A, D)
    S = Sets(D)
    contrib = D.table(L.TAB_CONTRIB, plan)
    slices = D.table(L.TAB_SLICE, plan)
    cptr = D.table(L.TAB_COLOR_PTR, plan)
    assert cptr.size >= 2 and cptr[0] == 0 and cptr[-1] == len(slices)
    nout = A.size[0] if op == "N" else A.size[1]
    dt = np.result_type(D.dtype, x.dtype)
    y = np.zeros(nout, dt) if (y is None or beta_false) else beta * y
    conj = op == "C"
    covered = np.zeros(len(contrib), np.int64)
    for l in range(len(cptr) - 1):
        touched = np.zeros(nout, bool)
        for s in slices[cptr[l]:cptr[l + 1]]:
            assert s["c_end"] - s["c_begin"] == 1 and (s["flags"] & 1)
            c = contrib[s["c_begin"]]
            r0, r1 = int(s["r0"]), int(s["r1"])
            assert 0 < r1 - r0 <= 128 and r1 <= c["out_len"]
            covered[s["c_begin"]] += r1 - r0
            m, n = int(c["m"]), int(c["n"])
            Bm = arena[c["off"]:c["off"] + m * n].reshape((m, n), order="F")
            if conj:
                Bm = Bm.conj()
            rows = S.idx(s["out_set"], r0, r1)
            assert not np.any(touched[rows]), "two slices of one colour share a row"
            touched[rows] = True
            if c["form"] & 1:
                y[rows] += alpha * (Bm[:, r0:r1].T @ x[S.idx(c["in_set"], 0, m)])
            else:
                y[rows] += alpha * (Bm[r0:r1, :] @ x[S.idx(c["in_set"], 0, n)])
    assert np.array_equal(covered, contrib["out_len"]), "every block must be applied exactly once"
    return y
