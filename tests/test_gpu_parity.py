"""Parity of the CUDA path against the oracle, through the C ABI (bsm_create_* / bsm_mul / bsm_mul_host).

Restates the reference's test battery (/root/reference/test/test_blockmatrix.jl:34-106,
test/test_symmetricblockmatrix.jl:45-107, test/test_vbcrs.jl:17-90) with the oracle in the role of
`sparse(A) * x`. Tolerances (north_star): rel ||dy||_2/||y||_2 <= 1e-12 for Float64 / ComplexF64,
<= 1e-5 for Float32; packing is bit-exact.
"""
import numpy as np
import pytest

import bsm_b200 as B
from bsm_b200 import _lib as L
from bsm_b200 import generators as G
from helpers import TOL, oracle_mul, randx, rel2, to_oracle
from oracle import oracle_np as O
from test_oracle import leaf_contiguous

pytestmark = pytest.mark.gpu
OPS = ["N", "T", "C"]


def wrap(A, op):
    return A if op == "N" else (B.transpose(A) if op == "T" else B.adjoint(A))


def battery(A, seed=0, ops=OPS, reps=2, variants=None):
    """A*x, A'*x, transpose(A)*x and 5-arg mul! (α=im|0.7, β=2im|-0.4) vs the oracle, for the stream
    (AUTO: TMA-staged kernels) and the GATHER (direct loads) plans."""
    if variants is None and not isinstance(A, B.SymmetricBlockMatrix):
        variants = (L.VARIANT_AUTO, L.VARIANT_GATHER)
        if A.device().table(L.TAB_COLOR_PTR, 4).size > 0:        # colour-ordered plan exists (no repeated indices)
            variants += (L.VARIANT_COLOR,)
    if variants:
        for v in variants:
            A.device().set_variant(v)
            battery(A, seed, ops, reps, variants=False)
        A.device().set_variant(L.VARIANT_AUTO)
        return
    dt = np.dtype(A.dtype)
    tol = TOL[dt]
    f64 = dt == np.float32
    rng = np.random.default_rng(seed)
    alpha, beta = (1j, 2j) if dt.kind == "c" else (0.7, -0.4)
    for op in ops:
        nin = A.size[1] if op == "N" else A.size[0]
        nout = A.size[0] if op == "N" else A.size[1]
        for _ in range(reps):
            x = randx(rng, nin, dt)
            y = wrap(A, op) * x
            assert y.dtype == dt and y.shape == (nout,)
            assert rel2(y, oracle_mul(A, x, op, f64=f64)) < tol, (op, "3-arg")
            y0 = randx(rng, nout, dt)
            y5 = B.mul_(y0.copy(), wrap(A, op), x, alpha, beta)
            assert rel2(y5, oracle_mul(A, x, op, alpha, beta, False, y0.copy(), f64=f64)) < tol, (op, "5-arg")


@pytest.fixture(scope="module", params=["cuboid", "sphere"])
def golden(request):
    return O.load_golden_sbm(request.param)


def test_symmetric_fixture(golden):
    A = B.SymmetricBlockMatrix(golden.diagonals, golden.diagonalindices, golden.offdiagonals,
                               golden.rowindices, golden.colindices, golden.size)
    for variant in (L.VARIANT_FUSED_TMA, L.VARIANT_FUSED, L.VARIANT_GATHER, L.VARIANT_COLOR):
        A.device().set_variant(variant)
        battery(A)
    assert A.device().launch_count("N") > 4                      # colour-ordered: scale + one launch per colour
    A.device().set_variant(L.VARIANT_AUTO)
    assert A.device().launch_count("N") == 2                     # fused kernel + finalize
    assert A.device().nnz() == B.nnz(A) == B.sparse(A).nnz      # test_symmetricblockmatrix.jl:99-107
    S = B.sparse(A)
    assert abs(S - S.T).nnz == 0                                 # issymmetric, :49


@pytest.mark.parametrize("name", ["cuboid", "sphere"])
def test_frozen_products_on_the_fixture(name):
    """Known answers (tests/golden/products_*.npz) on the reference's shipped fixture through the C ABI."""
    from pathlib import Path
    g = O.load_golden_sbm(name)
    z = np.load(Path(__file__).resolve().parent / "golden" / f"products_{name}.npz")
    A = B.SymmetricBlockMatrix(g.diagonals, g.diagonalindices, g.offdiagonals, g.rowindices, g.colindices, g.size)
    for op in OPS:
        assert rel2(wrap(A, op) * z["x"], z[f"y_{op}"]) < 1e-12
        assert rel2(B.mul_(z["y0"].copy(), wrap(A, op), z["x"], 1j, 2j), z[f"y5_{op}"]) < 1e-12


def test_blocksparse_fixture(golden):
    E = O.sbm_to_bsm(golden)
    A = B.BlockSparseMatrix(E.blocks, E.rowindices, E.colindices, E.size)
    battery(A)
    assert A.device().nnz() == B.nnz(A) == B.sparse(A).nnz


def test_vbcrs_fixture(golden):
    C = leaf_contiguous(golden)
    S = B.SymmetricBlockMatrix(C.diagonals, C.diagonalindices, C.offdiagonals, C.rowindices, C.colindices, C.size)
    E = O.sbm_to_bsm(C)
    Bm = B.BlockSparseMatrix(E.blocks, E.rowindices, E.colindices, E.size)
    for V in (B.VariableBlockCompressedRowStorage(S), B.VariableBlockCompressedRowStorage(Bm),
              B.VariableBlockCompressedRowStorage(Bm.blocks, [r[0] for r in Bm.rowindices],
                                                  [c[0] for c in Bm.colindices], Bm.size)):
        assert B.nnz(V) == B.nnz(S) == V.device().nnz()
        battery(V, reps=1)
        rng = np.random.default_rng(9)
        x = rng.standard_normal(V.size[1])                      # real x, complex blocks (test_vbcrs.jl:34)
        for op in OPS:
            a, b = wrap(V, op) * x, wrap(S, op) * x
            assert np.max(np.abs(a - b)) / np.max(np.abs(b)) < 1e-12


def test_materialised_matrix(golden):
    """A[:, :] — one product per unit vector (test_blockmatrix.jl:38-49), as a multi-RHS call."""
    A = B.SymmetricBlockMatrix(golden.diagonals, golden.diagonalindices, golden.offdiagonals,
                               golden.rowindices, golden.colindices, golden.size)
    n = A.size[0]
    cols = np.arange(0, n, 37)
    X = np.zeros((n, len(cols)), np.complex128, order="F")
    X[cols, np.arange(len(cols))] = 1
    S = B.sparse(A).toarray()
    assert np.max(np.abs(A * X - S[:, cols])) < 1e-13
    assert np.max(np.abs(B.adjoint(A) * X - S.conj().T[:, cols])) < 1e-13
    assert np.max(np.abs(B.transpose(A) * X - S.T[:, cols])) < 1e-13
    # getindex the LinearMaps way: b[:, :], adjoint(b)[:, :], single entries and slices
    m = min(n, 300)
    assert np.max(np.abs(A[:, :m] - S[:, :m])) < 1e-13
    assert np.max(np.abs(B.adjoint(A)[:, :m] - S.conj().T[:, :m])) < 1e-13
    assert abs(A[5, 7] - S[5, 7]) < 1e-13 and np.max(np.abs(A[3:9, 11] - S[3:9, 11])) < 1e-13
    E = O.sbm_to_bsm(golden)
    Ab = B.BlockSparseMatrix(E.blocks, E.rowindices, E.colindices, E.size)
    assert np.max(np.abs(B.transpose(Ab)[:, :m] - B.sparse(Ab).toarray().T[:, :m])) < 1e-13


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("permuted", [False, True])
def test_c1_shape(dtype, permuted):
    A = G.blocksparse_uniform(seed=11, n=2000, nblocks=300, bs=32, dtype=dtype, permuted=permuted)
    battery(A, reps=1)


@pytest.mark.parametrize("permuted", [False, True])
@pytest.mark.parametrize("dtype", [np.complex128, np.float64, np.float32])
def test_c2_shape(permuted, dtype):
    A = G.symmetric_nearfield(seed=12, n=12000, k_near=4, permuted=permuted, dtype=dtype)
    for variant in (L.VARIANT_FUSED_TMA, L.VARIANT_FUSED, L.VARIANT_GATHER):
        A.device().set_variant(variant)
        battery(A, reps=1)


def test_c2_heavy_segments_split():
    A = G.symmetric_nearfield(seed=20, n=6000, leaf_min=150, leaf_max=250, k_near=4)
    sl = A.device().table(L.TAB_SLICE, 2)
    assert np.unique(sl["out_set"], return_counts=True)[1].max() > 1
    battery(A, reps=1)


def test_c2_tall_leaves_mix_fused_and_gather_kernels():
    A = G.symmetric_nearfield(seed=19, n=20000, leaf_min=150, leaf_max=400, k_near=3)
    sl = A.device().table(L.TAB_SLICE, 2)
    # leaves taller than 256 rows: 256-row pieces of long N-form segments through the CTA-stream kernel
    assert np.any(((sl["flags"] & 4) != 0) & (sl["r0"] > 0))
    battery(A, reps=1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_tall_n_form_pieces_through_tensor_map_boxes(dtype):
    # long all-N-form segments: 256-row pieces; Float32 / Float64 fetch a piece of a chunk of columns as ONE tensor-map
    # box (blocks of two heights and several phases share maps, the last piece is shorter than the box), ComplexF64 keeps
    # the per-column copies
    from test_packing_cpu import tall_shared_rows_bsm
    blocks, rows, cols, size = tall_shared_rows_bsm(np.random.default_rng(41), np.float64)
    if np.dtype(dtype).kind == "c":
        rng = np.random.default_rng(5)
        blocks = [np.asfortranarray(b + 1j * rng.standard_normal(b.shape)) for b in blocks]
    else:
        blocks = [np.asfortranarray(b.astype(dtype)) for b in blocks]
    A = B.BlockSparseMatrix(blocks, rows, cols, size)
    sl = A.device().table(L.TAB_SLICE, 2)
    assert np.any(((sl["flags"] & 4) != 0) & (sl["r0"] > 0))
    battery(A, reps=1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_c3_shape(dtype):
    A = G.vbcrs_variable(seed=13, n=60000, dtype=dtype)
    battery(A, reps=1)
    Bm = G.vbcrs_variable(seed=13, n=60000, dtype=dtype, as_blocksparse=True)
    x = randx(np.random.default_rng(1), 60000, dtype)
    assert rel2(A * x, Bm * x) < TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c4_shape(dtype):
    A = G.blocksparse_large(seed=14, grid=6, bs=1024, density=0.2, dtype=dtype)
    battery(A, reps=1)


def test_odd_sizes_and_overlaps():
    from test_packing_cpu import random_bsm
    rng = np.random.default_rng(3)
    for dtype in (np.float32, np.float64, np.complex128):
        for contiguous in (False, True):
            blocks, rows, cols = random_bsm(rng, 400, 333, 120, dtype, contiguous, maxdim=40)
            battery(B.BlockSparseMatrix(blocks, rows, cols, (400, 333)), reps=1)


def test_beta_false_is_strong_zero_and_beta_zero_propagates():
    A = G.blocksparse_uniform(seed=15, n=640, nblocks=50, bs=32)
    x = np.ones(640)
    y = np.full(640, np.nan)
    out = B.mul_(y, A, x, True, False)
    assert np.all(np.isfinite(out))
    y = np.full(640, np.nan)
    out = B.mul_(y, A, x, 1.0, 0.0)
    assert np.all(np.isnan(out))


def test_arena_and_tables_bit_exact():
    A = G.symmetric_nearfield(seed=16, n=3000, k_near=3, permuted=True)
    D = A.device()
    arena = D.table(L.TAB_ARENA)
    off = D.table(L.TAB_BLOCK_OFF)
    blocks = list(A.diagonals) + list(A.offdiagonals)
    assert np.all((off * arena.itemsize) % 128 == 0)
    covered = np.zeros(arena.size, bool)
    for b, o in zip(blocks, off):
        assert np.array_equal(arena[o:o + b.size], b.reshape(-1, order="F"))     # verbatim, column-major
        covered[o:o + b.size] = True
    assert np.all(arena[~covered] == 0)
    assert D.stored_entries() == sum(b.size for b in blocks)
    # host-only handle produces the identical tables (same packer, no device)
    Dh = A.device(device=L.DEVICE_NONE)
    for t in (L.TAB_BLOCK_OFF, L.TAB_SET_LEN, L.TAB_SET_START, L.TAB_POOL):
        assert np.array_equal(D.table(t), Dh.table(t))
    for plan in (0, 1):
        for t in (L.TAB_CONTRIB, L.TAB_SLICE, L.TAB_GATHER_ROWS, L.TAB_GATHER_PTR, L.TAB_GATHER_POS, L.TAB_GROUP_PTR):
            assert np.array_equal(D.table(t, plan), Dh.table(t, plan))


def test_deterministic_and_torch_device_path():
    import torch
    A = G.symmetric_nearfield(seed=17, n=20000, k_near=4)
    D = A.device()
    rng = np.random.default_rng(0)
    x = randx(rng, 20000, np.complex128)
    xd = torch.from_numpy(x).cuda()
    y1 = D.mul("N", xd).cpu().numpy()
    y2 = D.mul("N", xd).cpu().numpy()
    assert np.array_equal(y1, y2)                      # atomic-free → bitwise reproducible
    assert np.array_equal(y1, D.mul("N", x))           # host path = device path
    assert rel2(y1, oracle_mul(A, x, "N")) < 1e-12
    # multi-RHS through device pointers, column-major X
    X = torch.randn(20000, 3, dtype=torch.complex128, device="cuda")
    Y = D.mul("C", X)
    for j in range(3):
        assert rel2(Y[:, j].cpu().numpy(), oracle_mul(A, X[:, j].cpu().numpy(), "C")) < 1e-12


def test_real_matrix_complex_x_promotes():
    A = G.vbcrs_variable(seed=18, n=5000)
    x = randx(np.random.default_rng(2), 5000, np.complex128)
    y = A * x
    ref = oracle_mul(A, np.ascontiguousarray(x.real), "N") + 1j * oracle_mul(A, np.ascontiguousarray(x.imag), "N")
    assert rel2(y, ref) < 1e-12


# ---- multi-RHS SpMM (C5): tensor-core kernels vs the oracle applied column by column ----------------------
def spmm_check(A, nrhs, ops=OPS, seed=3, expect_kernel=None):
    """Matrix right-hand sides of every dtype: spmm_tma_kernel (regular plans) / spmm_dmma_kernel (other Float64
    plans) / the column loop must all reproduce the oracle applied column by column (what LinearMaps does)."""
    import torch
    rng = np.random.default_rng(seed)
    D = A.device()
    dt = np.dtype(A.dtype)
    tol = TOL[dt]
    f64 = dt == np.float32
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
           np.dtype(np.complex128): torch.complex128}[dt]
    alpha, beta = (0.7 + 0.3j, -0.4 + 0.2j) if dt.kind == "c" else (0.7, -0.4)

    def rnd(shape):
        v = rng.standard_normal(shape)
        if dt.kind == "c":
            v = v + 1j * rng.standard_normal(shape)
        return np.asfortranarray(v.astype(dt))

    for op in ops:
        st = D.plan_stats(op)
        assert st["spmm"], "plan is not eligible for the SpMM kernels"
        if expect_kernel:
            assert st["spmm_kernel"] == expect_kernel, st
        nin = A.size[1] if op == "N" else A.size[0]
        nout = A.size[0] if op == "N" else A.size[1]
        X, Y0 = rnd((nin, nrhs)), rnd((nout, nrhs))
        ref = np.stack([oracle_mul(A, np.ascontiguousarray(X[:, j]), op, f64=f64) for j in range(nrhs)], axis=1)
        Y = wrap(A, op) * X                                        # host pointers (bsm_mul_host)
        assert Y.shape == (nout, nrhs) and Y.dtype == dt and rel2(Y, ref) < tol, (op, nrhs, rel2(Y, ref))
        ref5 = np.stack([oracle_mul(A, np.ascontiguousarray(X[:, j]), op, alpha, beta, False, Y0[:, j].copy(), f64=f64)
                         for j in range(nrhs)], axis=1)
        Y5 = B.mul_(Y0.copy(order="F"), wrap(A, op), X, alpha, beta)
        assert rel2(Y5, ref5) < tol, (op, nrhs, "5-arg", rel2(Y5, ref5))
        # device pointers with a padded leading dimension: even padding keeps the TMA path, odd padding breaks the
        # 16-byte stride rule of the tensor map and must fall back without changing the answer
        for pad in (4, 3):
            Xd = torch.zeros((nrhs, nin + pad), dtype=tdt, device="cuda").t()[:nin]
            Xd.copy_(torch.from_numpy(X))
            Yd = D.mul(op, Xd)
            assert rel2(Yd.cpu().numpy(), ref) < tol, (op, nrhs, "pad", pad)
        # the round-1 kernel (Float64) / the column loop over the SpMV kernels give the same answer
        for variant in (L.VARIANT_FUSED, L.VARIANT_GATHER):
            D.set_variant(variant)
            assert rel2(D.mul(op, X), ref) < tol, (op, nrhs, "variant", variant)
        D.set_variant(L.VARIANT_AUTO)


@pytest.mark.parametrize("nrhs", [8, 13, 64, 70])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_c5_shape_spmm(nrhs, dtype):
    A = G.blocksparse_uniform(seed=31, n=6400, nblocks=1500, bs=32, dtype=dtype)
    spmm_check(A, nrhs, expect_kernel="spmm_tma_kernel")


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_spmm_small_variable_blocks(dtype):
    # blocks of 8..32 rows / columns: masked (partial-slab) stages, segments shorter than 32 rows, wide blocks cut
    # into 32-wide contraction slabs
    spmm_check(G.vbcrs_variable(seed=36, n=12000, tile_min=8, tile_max=32, dtype=dtype), 24, expect_kernel="spmm_tma_kernel")
    spmm_check(G.vbcrs_variable(seed=37, n=4000, tile_min=1, tile_max=19, dtype=dtype), 9, expect_kernel="spmm_tma_kernel")


def test_spmm_wide_blocks_of_short_rows():
    # N-form blocks far wider than one contraction slab (32 x 100), T-form falls back (columns > 32)
    rng = np.random.default_rng(38)
    n = 3200
    blocks, rows, cols = [], [], []
    for r in range(0, n, 32):
        c0 = int(rng.integers(0, n - 100))
        blocks.append(np.asfortranarray(rng.standard_normal((32, 100))))
        rows.append(np.arange(r + 1, r + 33, dtype=np.int64))
        cols.append(np.arange(c0 + 1, c0 + 101, dtype=np.int64))
    A = B.BlockSparseMatrix(blocks, rows, cols, (n, n))
    spmm_check(A, 16, ops=("N",), expect_kernel="spmm_tma_kernel")


def test_spmm_variable_blocks_and_permuted_indices():
    spmm_check(G.vbcrs_variable(seed=32, n=20000), 24)                       # blocks up to 64 rows: round-1 kernel
    spmm_check(G.blocksparse_uniform(seed=33, n=3200, nblocks=400, bs=32, permuted=True), 16)   # index pool
    spmm_check(G.blocksparse_uniform(seed=34, n=2560, nblocks=300, bs=64), 32)                 # 64-row blocks


def test_spmm_deterministic():
    A = G.blocksparse_uniform(seed=35, n=6400, nblocks=1500, bs=32)
    X = np.asfortranarray(np.random.default_rng(0).standard_normal((6400, 64)))
    assert np.array_equal(A * X, A * X)


# ---- edge cases: empty and ragged inputs, rectangular operators, degenerate blocks ---------------------------
def test_empty_and_degenerate_inputs():
    rng = np.random.default_rng(21)
    # no blocks at all: y = beta*y (strong zero for beta === false)
    E = B.BlockSparseMatrix([], [], [], (7, 5))
    assert np.array_equal(E * np.ones(5), np.zeros(7))
    y0 = rng.standard_normal(7)
    assert np.allclose(B.mul_(y0.copy(), E, np.ones(5), 2.0, 3.0), 3.0 * y0)
    assert np.array_equal(B.transpose(E) * np.ones(7), np.zeros(5))
    # zero-row / zero-column blocks next to real ones, 1x1 blocks, a single wide and a single tall block
    blocks = [np.zeros((0, 3)), rng.standard_normal((1, 1)), np.zeros((2, 0)), rng.standard_normal((1, 9)),
              rng.standard_normal((6, 1)), rng.standard_normal((3, 4))]
    rows = [np.zeros(0, np.int64), [4], [1, 2], [5], [1, 2, 3, 4, 5, 6], [2, 3, 4]]
    cols = [[1, 2, 3], [9], np.zeros(0, np.int64), np.arange(1, 10), [3], [5, 6, 7, 8]]
    battery(B.BlockSparseMatrix(blocks, rows, cols, (6, 9)), reps=1)
    # rectangular operators, more rows than columns and the reverse, F32 / F64 / CF64
    for dt in (np.float32, np.float64, np.complex128):
        for shape in ((50, 13), (13, 50)):
            bl, r, c = [], [], []
            for _ in range(12):
                m, n = rng.integers(1, 8, 2)
                r0, c0 = rng.integers(1, shape[0] - m + 2), rng.integers(1, shape[1] - n + 2)
                b = rng.standard_normal((m, n)) + (1j * rng.standard_normal((m, n)) if np.dtype(dt).kind == "c" else 0)
                bl.append(np.asfortranarray(b.astype(dt)))
                r.append(np.arange(r0, r0 + m))
                c.append(np.arange(c0, c0 + n))
            battery(B.BlockSparseMatrix(bl, r, c, shape), reps=1)
    # symmetric matrix without off-diagonal blocks, and one with a single leaf
    d = [np.asfortranarray(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))) for _ in range(3)]
    d = [x + x.T for x in d]
    idx = [np.arange(1, 5), np.arange(5, 9), np.arange(9, 13)]
    battery(B.SymmetricBlockMatrix(d, idx, [], [], [], (12, 12)), reps=1)
    battery(B.SymmetricBlockMatrix(d[:1], idx[:1], [], [], [], (4, 4)), reps=1)
    # VBCRS with a single 1x1 block in a larger matrix
    battery(B.VariableBlockCompressedRowStorage([np.array([[2.5]])], [3], [2], (5, 4)), reps=1)


def test_dimension_mismatch_and_dtype_errors():
    A = G.blocksparse_uniform(seed=36, n=640, nblocks=20, bs=32)
    with pytest.raises(ValueError):
        A * np.ones(641)
    with pytest.raises(ValueError):
        B.mul_(np.zeros(639), A, np.ones(640))
    with pytest.raises(TypeError):
        B.BlockSparseMatrix([np.ones((2, 2), np.int32)], [[1, 2]], [[1, 2]], (2, 2)).device()


# ---- seeded fuzz through the real kernels (same structures as the CPU plan-interpreter fuzz) --------------------
@pytest.mark.parametrize("seed", range(6))
def test_fuzz_blocksparse(seed):
    rng = np.random.default_rng(100 + seed)
    dtype = [np.float64, np.complex128, np.float32][seed % 3]
    contiguous = seed % 2 == 0
    nrows, ncols = int(rng.integers(150, 400)), int(rng.integers(150, 400))
    blocks, rows, cols = [], [], []
    for _ in range(int(rng.integers(20, 60))):
        m, n = min(int(rng.integers(1, 91)), nrows), min(int(rng.integers(1, 91)), ncols)
        if contiguous:
            r0, c0 = int(rng.integers(1, nrows - m + 2)), int(rng.integers(1, ncols - n + 2))
            r, c = np.arange(r0, r0 + m), np.arange(c0, c0 + n)
        else:
            r, c = rng.permutation(nrows)[:m] + 1, rng.permutation(ncols)[:n] + 1
        b = rng.standard_normal((m, n))
        if np.dtype(dtype).kind == "c":
            b = b + 1j * rng.standard_normal((m, n))
        blocks.append(np.asfortranarray(b.astype(dtype)))
        rows.append(r.astype(np.int64))
        cols.append(c.astype(np.int64))
    battery(B.BlockSparseMatrix(blocks, rows, cols, (nrows, ncols)), reps=1)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_symmetric(seed):
    rng = np.random.default_rng(200 + seed)
    dtype = [np.complex128, np.float64][seed % 2]
    n = int(rng.integers(1500, 3000))
    A = G.symmetric_nearfield(seed=300 + seed, n=n, leaf_min=5, leaf_max=320, k_near=int(rng.integers(1, 5)),
                              dtype=dtype, permuted=bool(seed & 1))
    for variant in (L.VARIANT_FUSED_TMA, L.VARIANT_FUSED, L.VARIANT_GATHER, L.VARIANT_COLOR):
        A.device().set_variant(variant)
        battery(A, reps=1)
    A.device().set_variant(L.VARIANT_AUTO)
    V = B.VariableBlockCompressedRowStorage(G.symmetric_nearfield(seed=300 + seed, n=n, leaf_min=5, leaf_max=320,
                                                                  k_near=2, dtype=dtype)) if not (seed & 1) else None
    if V is not None:                                    # the same structure through the VBCRS conversion
        battery(V, reps=1)


def test_update_values_keeps_the_plan():
    """bsm_update_values: new block values in the existing arena, no re-planning (structure unchanged)."""
    A = G.symmetric_nearfield(seed=61, n=8000, k_near=3)
    A2 = G.symmetric_nearfield(seed=62, n=8000, k_near=3)          # same structure generator parameters, other values
    assert [b.shape for b in A.offdiagonals] != [] and len(A.offdiagonals) != 0
    # same structure: only valid when the shapes agree; build the second matrix on the first one's structure
    A2 = B.SymmetricBlockMatrix([2.0 * d for d in A.diagonals], A.diagonalindices,
                                [np.asfortranarray(-1.5 * o) for o in A.offdiagonals], A.rowindices, A.colindices, A.size)
    D = A.device()
    x = randx(np.random.default_rng(3), 8000, np.complex128)
    y1 = D.mul("N", x)
    assert rel2(y1, oracle_mul(A, x, "N")) < 1e-12
    D.update_values(A2)
    assert rel2(D.mul("N", x), oracle_mul(A2, x, "N")) < 1e-12
    assert rel2(D.mul("C", x), oracle_mul(A2, x, "C")) < 1e-12
    with pytest.raises(L.BsmError):
        L.check(L.lib().bsm_update_values(D._h, None, 3))


# ---- sparse(A) on the device vs the host restatement of src/sparse.jl -------------------------------------------
def sparse_check(A, exact_values=True):
    D = A.device()
    for op in OPS:
        S = D.sparse(op)
        if isinstance(A, B.VariableBlockCompressedRowStorage) and op != "N":
            # the reference defines rowcolvals only for the unwrapped VBCRS: compare with sparse(A)^T / sparse(A)'
            Hs = B.sparse(A).T.tocsc() if op == "T" else B.sparse(A).conj().T.tocsc()
            Hs.sort_indices()
        else:
            Hs = B.sparse(wrap(A, op))                  # host: rowcolvals + canonical CSC (scipy)
        assert S.shape == Hs.shape and S.nnz == Hs.nnz == (B.nnz(A) if exact_values else Hs.nnz)
        assert np.array_equal(S.indptr, Hs.indptr) and np.array_equal(S.indices, Hs.indices)   # structure bit-exact
        if exact_values:
            assert np.array_equal(S.data, Hs.data)      # no duplicates: values are copies (conj for op C)
        else:
            assert np.max(np.abs(S.data - Hs.data)) <= 1e-13 * max(1.0, np.max(np.abs(Hs.data)))


def test_sparse_on_device(golden):
    A = B.SymmetricBlockMatrix(golden.diagonals, golden.diagonalindices, golden.offdiagonals,
                               golden.rowindices, golden.colindices, golden.size)
    sparse_check(A)
    S = A.device().sparse("N")
    assert abs(S - S.T).nnz == 0                                     # issymmetric (test_symmetricblockmatrix.jl:49)
    E = O.sbm_to_bsm(golden)
    sparse_check(B.BlockSparseMatrix(E.blocks, E.rowindices, E.colindices, E.size))


def test_sparse_on_device_other_types():
    sparse_check(G.vbcrs_variable(seed=71, n=20000))
    sparse_check(G.blocksparse_uniform(seed=72, n=3200, nblocks=400, bs=32, dtype=np.float32, permuted=True))
    # overlapping blocks: duplicates are summed (values to rounding, structure exact); explicit zeros are kept
    from test_packing_cpu import random_bsm
    rng = np.random.default_rng(73)
    blocks, rows, cols = random_bsm(rng, 60, 50, 40, np.float64, contiguous=True, maxdim=12)
    blocks[0][:] = 0.0
    Ab = B.BlockSparseMatrix(blocks, rows, cols, (60, 50))
    sparse_check(Ab, exact_values=False)
    assert (Ab.device().sparse("N").data == 0).sum() > 0
    # empty matrix
    Es = B.BlockSparseMatrix([], [], [], (7, 5)).device().sparse("N")
    assert Es.shape == (7, 5) and Es.nnz == 0


# ---- solver loop on the device (bsm_cg) ------------------------------------------------------
@pytest.mark.parametrize("dtype,hermitian", [(np.float64, False), (np.complex128, False), (np.float32, False), (np.float64, True)])
def test_cg_through_the_operator(dtype, hermitian):
    """CG (real symmetric) / COCG (complex symmetric) kept on the device: the solution must satisfy A x = b to the
    requested tolerance when the residual is recomputed with the ORACLE's multiply, and two runs must agree bitwise."""
    import torch
    A = G.symmetric_nearfield(seed=61, n=20000, k_near=4, dtype=dtype, diag_shift=300.0 if dtype != np.complex128 else 300.0 + 60.0j)
    D = A.device()
    rng = np.random.default_rng(9)
    b = randx(rng, A.size[0], dtype)
    rtol = 1e-4 if dtype == np.float32 else 1e-10
    x, it, rel = D.cg(torch.from_numpy(b).cuda(), rtol=rtol, maxit=100, hermitian=hermitian)
    assert 2 <= it < 60 and rel <= rtol, (it, rel)
    xh = x.cpu().numpy()
    res = oracle_mul(A, xh, "N", f64=(dtype == np.float32)) - b
    assert np.linalg.norm(res) / np.linalg.norm(b) < (5e-4 if dtype == np.float32 else 5e-10)
    x2, it2, rel2_ = D.cg(torch.from_numpy(b).cuda(), rtol=rtol, maxit=100, hermitian=hermitian)
    assert it2 == it and torch.equal(x, x2)


# ---- device-side construction ------------------------------------------------------------------
def _to_cuda_colmajor(blocks, torch):
    out = []
    for b in blocks:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(b).T)).cuda().t()      # column-major on the device
        out.append(t)
    return out


@pytest.mark.parametrize("kind", ["bsm", "sbm", "vbcrs"])
def test_blocks_already_in_hbm_build_the_same_arena(kind):
    """bsm_options.blocks_on_device: the arena gathered in HBM from device-resident blocks is bit-identical to the one
    uploaded from host blocks, every table is the same, and the products agree bitwise; bsm_update_values_dev swaps in
    new values without re-planning."""
    import torch
    if kind == "bsm":
        A = G.blocksparse_uniform(seed=71, n=3200, nblocks=400, bs=32, dtype=np.complex128, permuted=True)
        blocks = A.blocks
    elif kind == "sbm":
        A = G.symmetric_nearfield(seed=72, n=8000, k_near=3)
        blocks = list(A.diagonals) + list(A.offdiagonals)
    else:
        A = G.vbcrs_variable(seed=73, n=9000, dtype=np.float32)
        blocks = A.blocks
    Dh = B.DeviceMatrix(A)
    dev_blocks = _to_cuda_colmajor(blocks, torch)
    Dd = B.DeviceMatrix(A, device_blocks=dev_blocks)
    for tab in (L.TAB_ARENA, L.TAB_BLOCK_OFF, L.TAB_POOL):
        assert np.array_equal(Dh.table(tab).view(np.uint8), Dd.table(tab).view(np.uint8)), tab
    for plan in (2, 3):
        for tab in (L.TAB_CONTRIB, L.TAB_SLICE, L.TAB_WCHUNK):
            assert np.array_equal(Dh.table(tab, plan), Dd.table(tab, plan))
    rng = np.random.default_rng(4)
    x = randx(rng, A.size[1], A.dtype)
    assert np.array_equal(Dh.mul("N", x), Dd.mul("N", x))
    # new values, same structure, straight from HBM
    scaled = [2 * t for t in dev_blocks]
    Dd.update_values_dev(scaled)
    y2 = Dd.mul("N", x)
    assert rel2(y2, 2 * oracle_mul(A, x, "N", f64=(np.dtype(A.dtype) == np.float32))) < TOL[np.dtype(A.dtype)]


def test_vbcrs_sorting_constructor_on_the_device():
    """bsm_vbcrs_sort_dev against the host constructor (host.py, stable lexsort = sortperm of src/vbcrs.jl:84): the
    permutation, rowptr, rowindices and colindices must be bit-identical, duplicates of (row start, column start)
    included (stability)."""
    import torch
    from bsm_b200.device import vbcrs_sort_device
    rng = np.random.default_rng(5)
    for nb, nrow, ncol in ((1, 1, 1), (50, 7, 9), (20000, 3000, 2500), (100000, 100000, 100000)):
        rs = rng.integers(1, nrow + 1, nb).astype(np.int64)
        cs = rng.integers(1, ncol + 1, nb).astype(np.int64)
        perm, rowptr, rowidx, colidx = vbcrs_sort_device(torch.from_numpy(rs).cuda(), torch.from_numpy(cs).cuda())
        href = np.lexsort((cs, rs))                       # stable, keyed by (row start, column start)
        assert np.array_equal(perm.cpu().numpy(), href)
        srs, scs = rs[href], cs[href]
        heads = np.flatnonzero(np.r_[True, srs[1:] != srs[:-1]])
        assert np.array_equal(rowptr.cpu().numpy(), np.r_[heads + 1, nb + 1])
        assert np.array_equal(rowidx.cpu().numpy(), srs[heads])
        assert np.array_equal(colidx.cpu().numpy(), scs)
    # the same through the host-side constructor of the mirror (what the reference's struct holds)
    V = G.vbcrs_variable(seed=74, n=20000)
    # the generator hands the blocks over unsorted; the mirror sorted them: recover the unsorted starts from its fields
    rs_sorted = np.repeat(V.rowindices, np.diff(V.rowptr))
    shuffle = rng.permutation(len(V.blocks))
    perm, rowptr, rowidx, colidx = vbcrs_sort_device(torch.from_numpy(rs_sorted[shuffle]).cuda(),
                                                     torch.from_numpy(np.asarray(V.colindices)[shuffle]).cuda())
    assert np.array_equal(rowptr.cpu().numpy(), V.rowptr) and np.array_equal(rowidx.cpu().numpy(), V.rowindices)
    assert np.array_equal(colidx.cpu().numpy(), V.colindices)
    assert np.array_equal(shuffle[perm.cpu().numpy()], np.arange(len(V.blocks)))


def test_persistent_cta_kernel_parity():
    """sym_persist_kernel (bsm_options.plan_hints bit 3; experimental, not the default — profiles/README.md): persistent
    CTAs with an issuer and an x-stager warp running ahead across work items must give the same products."""
    for A in (G.symmetric_nearfield(seed=81, n=30000, k_near=4), G.symmetric_nearfield(seed=82, n=9000, dtype=np.float32, permuted=True),
              G.symmetric_nearfield(seed=83, n=12000, leaf_min=150, leaf_max=250, k_near=4, dtype=np.float64)):
        D = B.DeviceMatrix(A, plan_hints=8)
        dt = np.dtype(A.dtype)
        rng = np.random.default_rng(6)
        alpha, beta = (1j, 2j) if dt.kind == "c" else (0.7, -0.4)
        for op in OPS:
            x = randx(rng, A.size[1], dt)
            y0 = randx(rng, A.size[0], dt)
            f64 = dt == np.float32
            assert rel2(D.mul(op, x), oracle_mul(A, x, op, f64=f64)) < TOL[dt], op
            assert rel2(D.mul(op, x, y0.copy(), alpha, beta), oracle_mul(A, x, op, alpha, beta, False, y0.copy(), f64=f64)) < TOL[dt], op
