"""Host logic on machines without a GPU: the packer's plan (host-only handle, device = BSM_DEVICE_NONE)
executed by the NumPy plan interpreter must reproduce the oracle for every type / op, including
overlapping blocks, repeated indices, uncovered rows and slab-restricted plans."""
import numpy as np
import pytest

import bsm_b200 as B
from bsm_b200 import _lib as L
from oracle import oracle_np as O
from plan_interp import run_plan
from test_oracle import leaf_contiguous

OPS = ["N", "T", "C"]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def host_only(A, **kw):
    return A.device(device=L.DEVICE_NONE, **kw)


@pytest.fixture(scope="module", params=["cuboid", "sphere"])
def fixture(request):
    return O.load_golden_sbm(request.param)


@pytest.mark.parametrize("op", OPS)
def test_sbm_plan(fixture, op):
    A = fixture
    P = B.SymmetricBlockMatrix(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    D = host_only(P)
    assert D.nnz() == O.nnz_sbm(A) == B.nnz(P)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(A.size[1]) + 1j * rng.standard_normal(A.size[1])
    y0 = rng.standard_normal(A.size[0]) + 1j * rng.standard_normal(A.size[0])
    for variant in ("fused", "gather"):
        assert rel(run_plan(P, D, op, x, variant=variant), O.mul_sbm(A, x, op)) < 1e-13
        assert rel(run_plan(P, D, op, x, 1j, 2j, False, y0.copy(), variant=variant),
                   O.mul_sbm(A, x, op, 1j, 2j, False, y0.copy())) < 1e-13
    # fused plan: every half-stored block appears exactly once, no separate transposed segments
    cf = D.table(L.TAB_CONTRIB, 2)
    assert len(cf) == len(A.diagonals) + len(A.offdiagonals)
    assert ((cf["form"] & 2) != 0).sum() == len(A.offdiagonals)
    # the leaf segments own their rows: diagonal + forward off-diagonal contributions are direct
    sl = D.table(L.TAB_SLICE, 0)
    assert (sl["flags"] & 1).sum() > 0


@pytest.mark.parametrize("op", OPS)
def test_bsm_plan(fixture, op):
    A = O.sbm_to_bsm(fixture)
    P = B.BlockSparseMatrix(A.blocks, A.rowindices, A.colindices, A.size)
    D = host_only(P)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.size[1]) + 1j * rng.standard_normal(A.size[1])
    assert rel(run_plan(P, D, op, x), O.mul_bsm(A, x, op)) < 1e-13


@pytest.mark.parametrize("op", OPS)
def test_vbcrs_plan(fixture, op):
    A = leaf_contiguous(fixture)
    Ps = B.SymmetricBlockMatrix(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    V = B.VariableBlockCompressedRowStorage(Ps)
    OV = O.vbcrs_from_sbm(A)
    # structure of the sorting constructor / conversion is bit-exact vs the oracle's literal loops
    assert np.array_equal(V.rowptr, OV.rowptr)
    assert np.array_equal(V.colindices, OV.colindices)
    assert np.array_equal(V.rowindices, OV.rowindices)
    assert all(np.array_equal(a, b) for a, b in zip(V.blocks, OV.blocks))
    D = host_only(V)
    assert D.nnz() == O.nnz_vbcrs(OV)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(A.size[1]) + 0j
    assert rel(run_plan(V, D, op, x), O.mul_vbcrs(OV, x, op)) < 1e-13
    # most segments own their rows (runs spanning two adjacent leaves overlap and go through scratch)
    for plan in (0, 1):
        assert np.mean(D.table(L.TAB_SLICE, plan)["flags"] & 1) > 0.5


def random_bsm(rng, nrows, ncols, nb, dtype, contiguous=False, maxdim=9):
    blocks, rows, cols = [], [], []
    for _ in range(nb):
        m, n = rng.integers(1, maxdim, 2)
        if contiguous:
            r0 = rng.integers(1, nrows - m + 2)
            c0 = rng.integers(1, ncols - n + 2)
            r, c = np.arange(r0, r0 + m), np.arange(c0, c0 + n)
        else:
            r = rng.integers(1, nrows + 1, m)      # repeated indices allowed
            c = rng.integers(1, ncols + 1, n)
        b = rng.standard_normal((m, n))
        if np.dtype(dtype).kind == "c":
            b = b + 1j * rng.standard_normal((m, n))
        blocks.append(np.asfortranarray(b.astype(dtype)))
        rows.append(r.astype(np.int64))
        cols.append(c.astype(np.int64))
    return blocks, rows, cols


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("contiguous", [False, True])
def test_overlapping_and_uncovered(dtype, contiguous):
    rng = np.random.default_rng(3)
    blocks, rows, cols = random_bsm(rng, 40, 33, 25, dtype, contiguous)
    A = O.OBSM(blocks, rows, cols, (40, 33))
    P = B.BlockSparseMatrix(blocks, rows, cols, (40, 33))
    D = host_only(P)
    tol = 1e-5 if dtype == np.float32 else 1e-13
    for op in OPS:
        nin = 33 if op == "N" else 40
        x = rng.standard_normal(nin).astype(dtype)
        y0 = rng.standard_normal(73 - nin).astype(dtype)
        ref = O.mul_bsm(A, x.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64), op)
        assert rel(run_plan(P, D, op, x), ref) < tol
        ref5 = O.mul_bsm(A, x, op, 0.5, -1.5, False, y0.copy())
        assert rel(run_plan(P, D, op, x, 0.5, -1.5, False, y0.copy()), ref5) < tol


def test_symmetric_tall_leaves_mix_fused_and_gather():
    # leaves taller than the fused kernel's 256 rows fall back to two separate contributions
    rng = np.random.default_rng(10)
    sizes = [300, 40, 257, 64, 256]
    bounds = np.cumsum([0] + sizes)
    idx = [np.arange(bounds[i] + 1, bounds[i + 1] + 1) for i in range(5)]
    cplx = lambda m, n: rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    diag = []
    for sz in sizes:
        d = cplx(sz, sz)
        diag.append(np.asfortranarray(d + d.T))
    off = [np.asfortranarray(cplx(sizes[i], sizes[i - 1] + sizes[0] * (i > 1))) for i in range(1, 5)]
    rows = [idx[i] for i in range(1, 5)]
    cols = [np.concatenate([idx[i - 1]] + ([idx[0]] if i > 1 else [])) for i in range(1, 5)]
    n = int(bounds[-1])
    A = O.OSBM(diag, idx, off, rows, cols, (n, n))
    P = B.SymmetricBlockMatrix(diag, idx, off, rows, cols, (n, n))
    D = host_only(P)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for op in OPS:
        for variant in ("fused", "gather"):
            assert rel(run_plan(P, D, op, x, variant=variant), O.mul_sbm(A, x, op)) < 1e-13
    sl = D.table(L.TAB_SLICE, 2)
    # leaves taller than 256 rows: since round 2 their all-N-form segments are cut into 256-row pieces of the
    # CTA-stream kernel (per-column piece copies) instead of 128-row gather slices
    assert np.any((sl["flags"] & 4).astype(bool) & (sl["r0"] > 0)), "expected row sub-range pieces"


def test_tall_and_wide_blocks_are_sliced():
    rng = np.random.default_rng(4)
    blocks = [np.asfortranarray(rng.standard_normal((300, 7))), np.asfortranarray(rng.standard_normal((5, 700))),
              np.asfortranarray(rng.standard_normal((300, 300)))]
    rows = [np.arange(1, 301), np.arange(301, 306), np.arange(1, 301)]
    cols = [np.arange(1, 8), np.arange(1, 701), np.arange(401, 701)]
    A = O.OBSM(blocks, rows, cols, (305, 700))
    P = B.BlockSparseMatrix(blocks, rows, cols, (305, 700))
    D = host_only(P)
    for op in OPS:
        x = rng.standard_normal(700 if op == "N" else 305)
        assert rel(run_plan(P, D, op, x), O.mul_bsm(A, x, op)) < 1e-13
    sl = D.table(L.TAB_SLICE, 0)
    assert np.all(sl["r1"] - sl["r0"] <= 128) and len(sl) >= 4


def tall_shared_rows_bsm(rng, dtype=np.float64):
    """Four tall blocks on ONE 700-row index set (plus one on a 520-row set and a short one): op N feeds long all-N-form
    segments, cut into 256-row pieces of the CTA-stream kernel and, being heavy, along the contribution list too."""
    mk = lambda m, n: np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
    blocks = [mk(700, 300), mk(700, 210), mk(700, 44), mk(700, 260), mk(520, 300), mk(30, 50)]
    r700, r520 = np.arange(1, 701), np.arange(701, 1221)
    rows = [r700, r700, r700, r700, r520, np.arange(1221, 1251)]
    cols = [np.arange(1, 301), np.arange(301, 511), np.arange(511, 555), np.arange(600, 860), np.arange(11, 311),
            np.arange(1, 51)]
    return blocks, rows, cols, (1250, 900)


def test_tall_n_form_pieces_are_cut_along_the_contribution_list():
    rng = np.random.default_rng(41)
    blocks, rows, cols, size = tall_shared_rows_bsm(rng)
    A = O.OBSM(blocks, rows, cols, size)
    P = B.BlockSparseMatrix(blocks, rows, cols, size)
    D = host_only(P)
    for op in OPS:
        x = rng.standard_normal(size[1] if op == "N" else size[0])
        assert rel(run_plan(P, D, op, x), O.mul_bsm(A, x, op)) < 1e-13
        y0 = rng.standard_normal(size[0] if op == "N" else size[1])
        assert rel(run_plan(P, D, op, x, 0.5, -1.5, False, y0.copy()), O.mul_bsm(A, x, op, 0.5, -1.5, False, y0.copy())) < 1e-13
    sl = D.table(L.TAB_SLICE, 2)      # stream plan of op N
    fused = (sl["flags"] & 4) != 0
    pieces = sl[fused & (sl["r1"] - sl["r0"] <= 256) & (sl["r1"] > 256)]
    assert len(pieces) >= 4, "expected 256-row pieces of the long N-form segments"
    # 256 x 300 x 8 bytes per contribution and piece is above the work-item budget: the partial items are not direct
    assert np.any((pieces["flags"] & 1) == 0) and np.any((pieces["flags"] & 1) != 0)
    assert np.any(pieces["c_end"] - pieces["c_begin"] == 1)


def test_clean_vbcrs_is_single_launch():
    rng = np.random.default_rng(8)
    tiles = np.cumsum(np.r_[1, rng.integers(3, 9, 12)])
    mats, rs, cs = [], [], []
    for i in range(12):
        for j in (i, (i + 1) % 12, (i + 5) % 12):
            mats.append(rng.standard_normal((tiles[i + 1] - tiles[i], tiles[j + 1] - tiles[j])))
            rs.append(tiles[i])
            cs.append(tiles[j])
    n = tiles[-1] - 1
    V = B.VariableBlockCompressedRowStorage(mats, rs, cs, (n, n))
    D = host_only(V)
    for plan in (0, 1):
        assert np.all(D.table(L.TAB_SLICE, plan)["flags"] & 1)
        assert D.table(L.TAB_GATHER_ROWS, plan).size == 0
    assert D.launch_count("N") == 1 and D.launch_count("T") == 1
    OV = O.vbcrs_from_blocks(mats, rs, cs, (n, n))
    for op in OPS:
        x = rng.standard_normal(n)
        assert rel(run_plan(V, D, op, x), O.mul_vbcrs(OV, x, op)) < 1e-13


def test_vbcrs_variable_heights_and_shared_starts():
    # blocks of one block row with different heights; two block rows overlapping in rows
    rng = np.random.default_rng(5)
    mats = [rng.standard_normal((4, 3)), rng.standard_normal((6, 2)), rng.standard_normal((5, 5)),
            rng.standard_normal((3, 3))]
    rs, cs = [1, 1, 4, 9], [1, 6, 2, 9]
    V = B.VariableBlockCompressedRowStorage(mats, rs, cs, (12, 12))
    OV = O.vbcrs_from_blocks(mats, rs, cs, (12, 12))
    D = host_only(V)
    for op in OPS:
        x = rng.standard_normal(12)
        assert rel(run_plan(V, D, op, x), O.mul_vbcrs(OV, x, op)) < 1e-13


@pytest.mark.parametrize("op", OPS)
def test_slab_plans_partition_the_product(fixture, op):
    """Row-slab plans (own_rows / own_cols): every rank writes only its slice, together they give y."""
    A = fixture
    P = B.SymmetricBlockMatrix(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    n = A.size[0]
    rng = np.random.default_rng(6)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ref = O.mul_sbm(A, x, op)
    cuts = [0, n // 3, 2 * n // 3 + 5, n]
    y = np.zeros(n, np.complex128)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        D = host_only(P, own_rows=(lo, hi), own_cols=(lo, hi))
        run_plan(P, D, op, x, y=y, own=(lo, hi))
    assert rel(y, ref) < 1e-13
    y = np.zeros(n, np.complex128)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        D = host_only(P, own_rows=(lo, hi), own_cols=(lo, hi))
        run_plan(P, D, op, x, y=y, own=(lo, hi), variant="gather")
    assert rel(y, ref) < 1e-13


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_warp_stream_schedule_c3_shape(dtype):
    """C3-shaped VBCRS: every block row goes to the warp-stream kernel; the chunk stream, its ring
    placement and the work items are replayed by the interpreter (several items, ring wrap-around)."""
    from bsm_b200 import generators as G
    V = G.vbcrs_variable(seed=21, n=30000, dtype=dtype)
    D = host_only(V)
    OV = O.OVBCRS(V.blocks, V.rowptr, V.colindices, V.rowindices, V.size)
    rng = np.random.default_rng(7)
    tol = 1e-5 if dtype == np.float32 else 1e-13
    for op in OPS:
        st = D.plan_stats(op)
        assert st["slices"]["stream_warp_kernel"] > 0 and st["slices"]["gather_gemv_kernel"] == 0
        # small problem: multi-block segments are cut into several work items (+ the gather pass)
        assert st["warp_items"] > 1 and D.launch_count(op) in (1, 2)
        x = rng.standard_normal(30000).astype(dtype)
        ref = O.mul_vbcrs(OV, x.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64), op)
        assert rel(run_plan(V, D, op, x), ref) < tol
    chunks = D.table(L.TAB_WCHUNK, 2)
    assert chunks["lag"].max() >= 3           # several chunks in flight per warp


def test_long_tform_segments_are_cut_into_column_ranges():
    """C4-shaped transpose product: 1024-row blocks, output segments of 1024 columns are cut into
    column sub-ranges for the CTA stream kernel (whole columns are contiguous in the arena)."""
    from bsm_b200 import generators as G
    A = G.blocksparse_large(seed=22, grid=3, bs=1024, density=0.5)
    D = host_only(A)
    st = D.plan_stats("T")
    assert st["slices"]["sym_fused_tma_kernel"] > 3 and st["slices"]["gather_gemv_kernel"] == 0
    OA = O.OBSM(A.blocks, A.rowindices, A.colindices, A.size)
    x = np.random.default_rng(8).standard_normal(A.size[0]).astype(np.float32)
    assert rel(run_plan(A, D, "T", x), O.mul_bsm(OA, x.astype(np.float64), "T")) < 1e-5
    assert rel(run_plan(A, D, "N", x), O.mul_bsm(OA, x.astype(np.float64), "N")) < 1e-5


@pytest.mark.parametrize("op", OPS)
def test_color_ordered_plans(fixture, op):
    """Colour-ordered comparison variant: greedy colouring per sweep, launches touch disjoint rows."""
    A = fixture
    P = B.SymmetricBlockMatrix(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    D = host_only(P)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(A.size[1]) + 1j * rng.standard_normal(A.size[1])
    y0 = rng.standard_normal(A.size[0]) + 1j * rng.standard_normal(A.size[0])
    assert rel(run_plan(P, D, op, x, variant="color"), O.mul_sbm(A, x, op)) < 1e-13
    assert rel(run_plan(P, D, op, x, 1j, 2j, False, y0.copy(), variant="color"),
               O.mul_sbm(A, x, op, 1j, 2j, False, y0.copy())) < 1e-13
    nl = len(D.table(L.TAB_COLOR_PTR, 4)) - 1
    assert 3 <= nl <= 64           # three sweeps; the transposed one needs many colours (column sets overlap)
    D.set_variant(L.VARIANT_COLOR)
    assert D.launch_count(op) == nl + 1
    E = O.sbm_to_bsm(A)
    Pb = B.BlockSparseMatrix(E.blocks, E.rowindices, E.colindices, E.size)
    assert rel(run_plan(Pb, host_only(Pb), op, x, variant="color"), O.mul_bsm(E, x, op)) < 1e-13


def test_color_variant_refuses_repeated_indices():
    rng = np.random.default_rng(12)
    blocks, rows, cols = random_bsm(rng, 40, 33, 25, np.float64, contiguous=False)
    D = host_only(B.BlockSparseMatrix(blocks, rows, cols, (40, 33)))
    with pytest.raises(L.BsmError):
        D.set_variant(L.VARIANT_COLOR)


@pytest.mark.parametrize("op", OPS)
def test_heavy_segments_are_split_across_work_items(op):
    """A leaf segment streaming far more than the work-item budget is cut: wide off-diagonal blocks enter the
    plan as column ranges (own input sub-set, own transposed partial), the first item stays direct, the
    others deliver partial vectors through the gather lists."""
    from bsm_b200 import generators as G
    A = G.symmetric_nearfield(seed=23, n=2400, leaf_min=150, leaf_max=250, k_near=4)
    D = host_only(A)
    OA = O.OSBM(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    plan = 2 if op == "N" else 3
    sl = D.table(L.TAB_SLICE, plan)
    cf = D.table(L.TAB_CONTRIB, plan)
    assert len(cf) > len(A.diagonals) + len(A.offdiagonals)          # some blocks were split by columns
    segs, counts = np.unique(sl["out_set"], return_counts=True)
    assert counts.max() > 1                                           # several work items for one segment
    for sset in segs[counts > 1]:
        assert ((sl["flags"][sl["out_set"] == sset] & 1) != 0).sum() <= 1   # at most one of them writes y directly
    rng = np.random.default_rng(13)
    x = rng.standard_normal(2400) + 1j * rng.standard_normal(2400)
    y0 = rng.standard_normal(2400) + 1j * rng.standard_normal(2400)
    assert rel(run_plan(A, D, op, x), O.mul_sbm(OA, x, op)) < 1e-13
    assert rel(run_plan(A, D, op, x, 1j, 2j, False, y0.copy()), O.mul_sbm(OA, x, op, 1j, 2j, False, y0.copy())) < 1e-13


def test_errors():
    b = [np.ones((2, 2))]
    with pytest.raises(L.BsmError):      # index out of range
        host_only(B.BlockSparseMatrix(b, [np.array([1, 5])], [np.array([1, 2])], (3, 3)))
    with pytest.raises(L.BsmError):      # size mismatch
        host_only(B.BlockSparseMatrix(b, [np.array([1, 2, 3])], [np.array([1, 2])], (3, 3)))
    with pytest.raises(IndexError):      # empty VBCRS throws in the reference too
        B.VariableBlockCompressedRowStorage([], [], [], (1, 1))
    D = host_only(B.BlockSparseMatrix(b, [np.array([1, 2])], [np.array([1, 2])], (3, 3)))
    with pytest.raises(L.BsmError):      # no CPU fallback
        D.mul("N", np.ones(3))


def test_abi_exports_every_declared_symbol():
    import re
    from pathlib import Path
    hdr = (Path(__file__).resolve().parent.parent / "include" / "bsm_b200.h").read_text()
    declared = set(re.findall(r"\b(bsm_[a-z0-9_]+)\s*\(", hdr))
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == {n for n, _, _ in L.SIGNATURES}


@pytest.mark.parametrize("seed", range(6))
def test_random_structures_all_plans_and_slabs(seed):
    """Seeded fuzz of the packer: random block sizes (1..90 rows/cols, so segments land in every kernel class),
    contiguous or arbitrary index vectors, overlaps, all three ops, every plan family (stream / gather / colour
    where it exists) and a 3-way slab split with the local/remote phase partition — all replayed by the plan
    interpreter against the oracle."""
    rng = np.random.default_rng(100 + seed)
    dtype = [np.float64, np.complex128, np.float32][seed % 3]
    contiguous = seed % 2 == 0
    nrows, ncols = int(rng.integers(150, 400)), int(rng.integers(150, 400))
    blocks, rows, cols = [], [], []
    for _ in range(int(rng.integers(20, 60))):
        m, n = int(rng.integers(1, 91)), int(rng.integers(1, 91))
        m, n = min(m, nrows), min(n, ncols)
        if contiguous:
            r0, c0 = int(rng.integers(1, nrows - m + 2)), int(rng.integers(1, ncols - n + 2))
            r, c = np.arange(r0, r0 + m), np.arange(c0, c0 + n)
        else:
            r, c = rng.permutation(nrows)[:m] + 1, rng.permutation(ncols)[:n] + 1    # no repeats inside a block
        b = rng.standard_normal((m, n))
        if np.dtype(dtype).kind == "c":
            b = b + 1j * rng.standard_normal((m, n))
        blocks.append(np.asfortranarray(b.astype(dtype)))
        rows.append(r.astype(np.int64))
        cols.append(c.astype(np.int64))
    A = O.OBSM(blocks, rows, cols, (nrows, ncols))
    P = B.BlockSparseMatrix(blocks, rows, cols, (nrows, ncols))
    D = host_only(P)
    tol = 1e-4 if dtype == np.float32 else 1e-12
    up = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
    has_color = D.table(L.TAB_COLOR_PTR, 4).size > 0
    for op in OPS:
        nin, nout = (ncols, nrows) if op == "N" else (nrows, ncols)
        x = rng.standard_normal(nin).astype(dtype)
        y0 = rng.standard_normal(nout).astype(dtype)
        ref = O.mul_bsm(A, x.astype(up), op)
        ref5 = O.mul_bsm(A, x.astype(up), op, 0.5, -1.5, False, y0.astype(up))
        for variant in ("auto", "gather") + (("color",) if has_color else ()):
            assert rel(run_plan(P, D, op, x, variant=variant), ref) < tol, (op, variant)
            assert rel(run_plan(P, D, op, x, 0.5, -1.5, False, y0.copy(), variant=variant), ref5) < tol, (op, variant)
    if nrows == ncols or True:
        # slabs over the OUTPUT dimension of op N (rows); x ownership follows the column cuts of the same fractions
        cuts_r = [0, nrows // 3, 2 * nrows // 3, nrows]
        cuts_c = [0, ncols // 3, 2 * ncols // 3, ncols]
        x = rng.standard_normal(ncols).astype(dtype)
        y = np.zeros(nrows, up)
        for k in range(3):
            Ds = host_only(P, own_rows=(cuts_r[k], cuts_r[k + 1]), own_cols=(cuts_c[k], cuts_c[k + 1]))
            run_plan(P, Ds, "N", x.astype(up), y=y, own=(cuts_r[k], cuts_r[k + 1]), in_own=(cuts_c[k], cuts_c[k + 1]))
        assert rel(y, O.mul_bsm(A, x.astype(up), "N")) < tol


@pytest.mark.parametrize("seed", range(4))
def test_random_symmetric_structures(seed):
    """Seeded fuzz of the symmetric path: leaf sizes from 5 to 320 rows (warp-eligible, CTA-kernel and taller-than-256
    leaves in one matrix), contiguous or renumbered unknowns, all ops, stream / gather / colour plans, 2-way slabs."""
    from bsm_b200 import generators as G
    rng = np.random.default_rng(200 + seed)
    dtype = [np.complex128, np.float64][seed % 2]
    n = int(rng.integers(1500, 3000))
    A = G.symmetric_nearfield(seed=300 + seed, n=n, leaf_min=5, leaf_max=320, k_near=int(rng.integers(1, 5)),
                              dtype=dtype, permuted=bool(seed & 1))
    OA = O.OSBM(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    D = host_only(A)
    x = rng.standard_normal(n).astype(dtype)
    y0 = rng.standard_normal(n).astype(dtype)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(n)
    for op in OPS:
        ref = O.mul_sbm(OA, x, op)
        for variant in ("fused", "gather", "color"):
            assert rel(run_plan(A, D, op, x, variant=variant), ref) < 1e-12, (op, variant)
        assert rel(run_plan(A, D, op, x, 0.7, -0.4, False, y0.copy()), O.mul_sbm(OA, x, op, 0.7, -0.4, False, y0.copy())) < 1e-12
    if not (seed & 1):
        from bsm_b200.partition import extract_slab, slab_cuts
        cuts = slab_cuts(A, 2)
        y = np.zeros(n, np.result_type(dtype, np.float64))
        for k in range(2):
            lo, hi = int(cuts[k]), int(cuts[k + 1])
            S = extract_slab(A, lo, hi, ("N",))
            run_plan(S, host_only(S, own_rows=(lo, hi), own_cols=(lo, hi)), "N", x, y=y, own=(lo, hi), in_own=(lo, hi))
        assert rel(y, O.mul_sbm(OA, x, "N")) < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_small_problem_cta_part_mode_is_single_launch(dtype):
    """C1 shape (BASELINE configs[0]): 312 block rows of ~6 blocks each. Small problems with few segments run as ONE
    launch: every segment is a CTA of the warp-stream kernel, its chunk list dealt to the four warps (CtaPart records),
    the rows no block touches are set by extra CTAs of the same launch — no partial sums through scratch."""
    from bsm_b200 import generators as G
    from helpers import oracle_mul
    A = G.blocksparse_uniform(seed=1, dtype=dtype)
    D = host_only(A)
    rng = np.random.default_rng(2)
    for op, plan in (("N", 2), ("T", 3)):
        ch = D.table(L.TAB_WCHUNK, plan)
        iptr = D.table(L.TAB_WITEM_PTR, plan)
        sl = D.table(L.TAB_SLICE, plan)
        assert np.any(ch["flags"] & 64) and len(iptr) - 1 == 4 * len(sl)
        assert np.all(sl["flags"] & 1), "every block row owns its rows"
        assert D.plan_stats(op)["scratch_elems"] == 0 and D.launch_count(op) == 1
        x = rng.standard_normal(A.size[1]).astype(dtype)
        y0 = rng.standard_normal(A.size[0]).astype(dtype)
        tol = 1e-5 if dtype == np.float32 else 1e-13
        assert rel(run_plan(A, D, op, x, variant="fused"), oracle_mul(A, x, op, f64=(dtype == np.float32))) < tol
        assert rel(run_plan(A, D, op, x, 0.7, -0.4, False, y0.copy(), variant="fused"),
                   oracle_mul(A, x, op, 0.7, -0.4, False, y0.copy(), f64=(dtype == np.float32))) < tol
    # the explicit hint restores per-block work items + gather pass (comparison)
    D2 = host_only(A, plan_hints=4)
    assert not np.any(D2.table(L.TAB_WCHUNK, 2)["flags"] & 64) and D2.launch_count("N") == 2


def test_warp_stream_chunks_are_as_large_as_the_ring_allows():
    # the per-chunk cost of stream_warp_kernel is fixed (~1 us of a warp's time), so the chunk COUNT is what a multiply
    # pays for: chunks of a VBCRS with 8..64-row blocks must average well above the round-1 limit of 4 KB per chunk
    # wherever the blocks are big enough, and two chunks (with their x values) must always fit the 11 KB ring
    from bsm_b200 import generators as G
    V = G.vbcrs_variable(seed=3, n=30000)
    D = host_only(V)
    for plan, tform in ((2, 0), (3, 1)):
        ch = D.table(L.TAB_WCHUNK, plan)
        assert len(ch) > 0 and np.all((ch["flags"] & 1) == tform)
        payload = ch["m"].astype(np.int64) * ch["ncols"] * 8
        cnt = np.where(ch["flags"] & 1, ch["m"], ch["ncols"]).astype(np.int64)
        foot = ch["bytes16"].astype(np.int64) * 16 + (cnt * 8 + 15) // 16 * 16 + 16
        assert foot.max() <= 11264 // 2
        stored = sum(int(b.size) for b in V.blocks) * 8
        assert payload.sum() == stored
        # blocks of 8..64 x 8..64 doubles average 11.7 KB: with <= ~5.3 KB per chunk that is ~2.7-2.9 chunks per block
        # (3.06 when the columns per chunk were rounded DOWN to a multiple of 4: 36 columns became 16 + 16 + 4)
        assert len(ch) <= 2.9 * len(V.blocks), (len(ch), len(V.blocks))
        assert payload.mean() >= 4000
