"""Pins the oracle (CPU restatement) before it is allowed to judge the CUDA path.

Restates the reference's own test battery on the one fixture it ships:
  /root/reference/test/test_symmetricblockmatrix.jl:45-107  (SBM: sparse, issymmetric, A*x, A'*x,
      transpose(A)*x, 5-arg mul! with α=im, β=2im, nnz)
  /root/reference/test/test_blockmatrix.jl:34-91            (BSM battery, on the expanded fixture)
  /root/reference/test/test_vbcrs.jl:17-90                  (VBCRS vs BSM vs sparse, 1e-13 rel max-abs)
and cross-checks C oracle ≡ NumPy oracle ≡ SciPy CSC product of the restated sparse(A).
"""
import numpy as np
import pytest

from oracle import oracle_np as O

EXAMPLES = ["cuboid", "sphere"]
EXPECT = {"cuboid": (1344, 96, 92, 21264, 93842), "sphere": (1203, 106, 103, 16501, 87718)}


def relmax(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.fixture(scope="module", params=EXAMPLES)
def sbm(request):
    return request.param, O.load_golden_sbm(request.param)


def test_fixture_sanity(sbm):
    name, A = sbm
    n, nd, no, ed, eo = EXPECT[name]
    assert A.size == (n, n)
    assert (len(A.diagonals), len(A.offdiagonals)) == (nd, no)
    assert sum(d.size for d in A.diagonals) == ed
    assert sum(o.size for o in A.offdiagonals) == eo
    assert O.nnz_sbm(A) == ed + 2 * eo
    if name == "cuboid":
        np.testing.assert_allclose(A.diagonals[0][0, :2],
                                   [-0.50865884 + 0.17362733j, -0.31457214 - 0.33453876j], atol=1e-8)


def test_sbm_sparse_properties(sbm):
    _, A = sbm
    S = O.sparse_sbm(A)
    assert (abs(S - S.T)).nnz == 0                 # issymmetric(sparse(b)), test_symmetricblockmatrix.jl:49
    assert S.nnz == O.nnz_sbm(A)                   # nnz(b) == nnz(bsparse), :99-107
    assert O.sparse_sbm(A, "T").nnz == S.nnz
    assert abs(O.sparse_sbm(A, "C") - S.conj()).max() == 0   # A' = conj(A) for complex-symmetric A


@pytest.mark.parametrize("op", ["N", "T", "C"])
def test_sbm_mul_vs_csc(sbm, op):
    _, A = sbm
    rng = np.random.default_rng(1)
    S = O.sparse_sbm(A)
    Sop = {"N": S, "T": S.T, "C": S.conj().T}[op]
    Cm = O.CSbm(A, threads=1)
    Cp = O.CSbm(A, threads=4)
    for _ in range(3):
        x = rng.standard_normal(A.size[1]) + 1j * rng.standard_normal(A.size[1])
        ref = Sop @ x
        assert relmax(O.mul_sbm(A, x, op), ref) < 1e-13
        assert relmax(Cm.mul(x, op), ref) < 1e-13
        assert relmax(Cp.mul(x, op), ref) < 1e-13
        y0 = rng.standard_normal(A.size[0]) + 1j * rng.standard_normal(A.size[0])
        ref5 = 1j * (Sop @ x) + 2j * y0            # mul!(x, b, y, im, 2im), :82-97
        assert relmax(O.mul_sbm(A, x, op, 1j, 2j, False, y0.copy()), ref5) < 1e-13
        assert relmax(Cm.mul(x, op, 1j, 2j, False, y0.copy()), ref5) < 1e-13
        assert relmax(Cp.mul(x, op, 1j, 2j, False, y0.copy()), ref5) < 1e-13


def test_beta_false_is_strong_zero(sbm):
    _, A = sbm
    x = np.ones(A.size[1], np.complex128)
    y = np.full(A.size[0], np.nan + 0j)
    out = O.c_mul_sbm(A, x, "N", 1, 0, True, y)
    assert np.all(np.isfinite(out))
    y = np.full(A.size[0], np.nan + 0j)
    out = O.c_mul_sbm(A, x, "N", 1, 0, False, y)   # numeric β = 0 propagates NaN
    assert np.all(np.isnan(out))


@pytest.mark.parametrize("op", ["N", "T", "C"])
def test_bsm_mul_vs_csc(sbm, op):
    _, A = sbm
    B = O.sbm_to_bsm(A)
    rng = np.random.default_rng(2)
    S = O.sparse_bsm(B)
    assert S.nnz == O.nnz_bsm(B)
    Sop = {"N": S, "T": S.T, "C": S.conj().T}[op]
    assert abs(O.sparse_bsm(B, op) - Sop).max() == 0
    x = rng.standard_normal(B.size[1]) + 1j * rng.standard_normal(B.size[1])
    ref = Sop @ x
    assert relmax(O.mul_bsm(B, x, op), ref) < 1e-13
    assert relmax(O.c_mul_bsm(B, x, op, threads=1), ref) < 1e-13
    assert relmax(O.c_mul_bsm(B, x, op, threads=4), ref) < 1e-13
    y0 = rng.standard_normal(B.size[0]) + 1j * rng.standard_normal(B.size[0])
    ref5 = 1j * ref + 2j * y0
    assert relmax(O.c_mul_bsm(B, x, op, 1j, 2j, False, y0.copy(), threads=4), ref5) < 1e-13


def leaf_contiguous(A):
    """Renumber unknowns so every leaf (diagonal index set) is a contiguous ascending range and
    split every off-diagonal block into column pieces that are contiguous ranges — the
    preprocessing the reference's VBCRS conversion assumes (src/vbcrs.jl:146-147)."""
    n = A.size[0]
    new = np.zeros(n + 1, np.int64)
    order = np.concatenate(A.diagonalindices)
    assert len(order) == n and len(set(order.tolist())) == n
    new[order] = np.arange(1, n + 1)
    dix = [new[d] for d in A.diagonalindices]
    offs, rows, cols = [], [], []
    for o, r, c in zip(A.offdiagonals, A.rowindices, A.colindices):
        rn, cn = new[r], new[c]
        assert np.array_equal(rn, np.arange(rn[0], rn[0] + len(rn)))
        cuts = np.flatnonzero(np.diff(cn) != 1) + 1
        for lo, hi in zip(np.r_[0, cuts], np.r_[cuts, len(cn)]):
            offs.append(np.asfortranarray(o[:, lo:hi]))
            rows.append(rn)
            cols.append(cn[lo:hi])
    return O.OSBM(A.diagonals, dix, offs, rows, cols, A.size)


def test_vbcrs_battery(sbm):
    """test_vbcrs.jl:17-48 and :53-90 on the leaf-contiguous renumbering of the fixture."""
    _, A0 = sbm
    A = leaf_contiguous(A0)
    B = O.sbm_to_bsm(A)
    V1 = O.vbcrs_from_blocks(B.blocks, [r[0] for r in B.rowindices], [c[0] for c in B.colindices], B.size)
    V2 = O.vbcrs_from_bsm(B)
    V3 = O.vbcrs_from_sbm(A)
    rng = np.random.default_rng(3)
    for V in (V1, V2, V3):
        assert O.nnz_vbcrs(V) == O.nnz_bsm(B) == O.nnz_sbm(A)
        assert V.rowptr[0] == 1 and V.rowptr[-1] == len(V.blocks) + 1
        assert np.all(np.diff(V.rowindices) > 0)
        S = O.sparse_vbcrs(V)
        assert abs(S - O.sparse_sbm(A)).max() == 0
        for _ in range(2):
            x = rng.standard_normal(V.size[1])          # real x, complex blocks (test_vbcrs.jl:34)
            xc = x.astype(np.complex128)
            den = np.max(np.abs(S @ x))
            for op, Sop in (("N", S), ("T", S.T), ("C", S.conj().T)):
                ref = Sop @ x
                assert np.max(np.abs(O.mul_vbcrs(V, x, op) - ref)) / den < 1e-13
                assert np.max(np.abs(O.c_mul_vbcrs(V, xc, op) - ref)) / den < 1e-13
                assert np.max(np.abs(O.c_mul_vbcrs(V, xc, op, threads=4) - ref)) / den < 1e-13
                assert np.max(np.abs(O.mul_bsm(B, x, op) - ref)) / den < 1e-13
                assert np.max(np.abs(O.mul_sbm(A, x, op) - ref)) / den < 1e-13


def test_vbcrs_sort_is_stable_and_rowptr():
    # two blocks with the same (row, col) start keep input order; height is per block
    b = [np.full((2, 2), 1.0), np.full((3, 2), 2.0), np.full((2, 1), 3.0), np.full((1, 1), 4.0)]
    V = O.vbcrs_from_blocks(b, [5, 1, 5, 1], [1, 3, 1, 1], (8, 8))
    assert [blk[0, 0] for blk in V.blocks] == [4.0, 2.0, 1.0, 3.0]
    assert V.rowptr.tolist() == [1, 3, 5]
    assert V.rowindices.tolist() == [1, 5]
    assert V.colindices.tolist() == [1, 3, 1, 1]
    x = np.arange(1.0, 9.0)
    y = O.mul_vbcrs(V, x)
    yc = O.c_mul_vbcrs(V, x)
    S = O.sparse_vbcrs(V)          # duplicates (two blocks on (5,1)) are summed
    np.testing.assert_allclose(y, S @ x)
    np.testing.assert_allclose(yc, S @ x)
    with pytest.raises(IndexError):
        O.vbcrs_from_blocks([], [], [], (1, 1))


def test_overlapping_blocks_accumulate():
    # overlapping BSM blocks and repeated indices inside one index vector accumulate (+=)
    rng = np.random.default_rng(5)
    blocks = [rng.standard_normal((3, 2)), rng.standard_normal((2, 2)), rng.standard_normal((2, 3))]
    rows = [np.array([1, 2, 2]), np.array([2, 4]), np.array([4, 1])]
    cols = [np.array([1, 3]), np.array([3, 3]), np.array([2, 1, 4])]
    B = O.OBSM(blocks, rows, cols, (4, 4))
    dense = np.zeros((4, 4))
    for b, r, c in zip(blocks, rows, cols):
        for i, ri in enumerate(r):
            for j, cj in enumerate(c):
                dense[ri - 1, cj - 1] += b[i, j]
    x = rng.standard_normal(4)
    for op, D in (("N", dense), ("T", dense.T), ("C", dense.T)):
        np.testing.assert_allclose(O.mul_bsm(B, x, op), D @ x, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(O.c_mul_bsm(B, x, op), D @ x, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(O.c_mul_bsm(B, x, op, threads=3), D @ x, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(O.sparse_bsm(B).toarray(), dense, rtol=1e-14)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_real_dtypes(dt):
    rng = np.random.default_rng(7)
    blocks = [rng.standard_normal((5, 4)).astype(dt) for _ in range(6)]
    rows = [np.arange(1, 6) + 5 * (i % 3) for i in range(6)]
    cols = [np.arange(1, 5) + 4 * (i // 2) for i in range(6)]
    B = O.OBSM(blocks, rows, cols, (15, 12))
    x = rng.standard_normal(12).astype(dt)
    tol = 1e-5 if dt == np.float32 else 1e-13
    ref = O.sparse_bsm(B).astype(np.float64) @ x.astype(np.float64)
    assert relmax(O.c_mul_bsm(B, x, "N").astype(np.float64), ref) < tol
    xt = rng.standard_normal(15).astype(dt)
    reft = O.sparse_bsm(B).astype(np.float64).T @ xt.astype(np.float64)
    assert relmax(O.c_mul_bsm(B, xt, "T").astype(np.float64), reft) < tol


@pytest.mark.parametrize("op", ["N", "T", "C"])
def test_frozen_products(sbm, op):
    """Known answers on the shipped fixture (tests/golden/products_*.npz, made by make_products.py from the CSC
    product): the C oracle, the NumPy oracle and the expanded BlockSparseMatrix / VBCRS forms all reproduce them."""
    from pathlib import Path
    name, A = sbm
    z = np.load(Path(__file__).resolve().parent / "golden" / f"products_{name}.npz")
    x, y0 = z["x"], z["y0"]
    assert relmax(O.mul_sbm(A, x, op), z[f"y_{op}"]) < 1e-13
    assert relmax(O.c_mul_sbm(A, x, op, threads=4), z[f"y_{op}"]) < 1e-13
    assert relmax(O.c_mul_sbm(A, x, op, 1j, 2j, False, y0.copy(), 4), z[f"y5_{op}"]) < 1e-13
    E = O.sbm_to_bsm(A)
    assert relmax(O.c_mul_bsm(E, x, op, threads=4), z[f"y_{op}"]) < 1e-13


@pytest.mark.parametrize("seed", range(4))
def test_oracles_agree_on_random_structures(seed):
    """C oracle = NumPy oracle = SciPy CSC product on random overlapping structures (all three storage types)."""
    rng = np.random.default_rng(400 + seed)
    nr, nc = int(rng.integers(60, 200)), int(rng.integers(60, 200))
    blocks, rows, cols = [], [], []
    for _ in range(int(rng.integers(10, 40))):
        m, n = int(rng.integers(1, 30)), int(rng.integers(1, 30))
        r0, c0 = int(rng.integers(1, nr - min(m, nr) + 2)), int(rng.integers(1, nc - min(n, nc) + 2))
        m, n = min(m, nr - r0 + 1), min(n, nc - c0 + 1)
        blocks.append(np.asfortranarray(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))))
        rows.append(np.arange(r0, r0 + m))
        cols.append(np.arange(c0, c0 + n))
    A = O.OBSM(blocks, rows, cols, (nr, nc))
    S = O.sparse_bsm(A)
    V = O.vbcrs_from_blocks(blocks, [r[0] for r in rows], [c[0] for c in cols], (nr, nc))
    for op, M in (("N", S), ("T", S.T), ("C", S.conj().T)):
        x = rng.standard_normal(M.shape[1]) + 1j * rng.standard_normal(M.shape[1])
        ref = M @ x
        assert relmax(O.mul_bsm(A, x, op), ref) < 1e-12
        assert relmax(O.c_mul_bsm(A, x, op, threads=4), ref) < 1e-12
        assert relmax(O.mul_vbcrs(V, x, op), ref) < 1e-12
        assert relmax(O.c_mul_vbcrs(V, x, op, threads=4), ref) < 1e-12      # overlapping block rows: serial schedule
