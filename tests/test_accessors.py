"""The reference's accessor battery restated for host.py (no GPU): /root/reference/test/test_blockmatrix.jl:84-106
(nnz under wrappers, eachblockindex equalities, eltype(block(...)) under adjoint / transpose) and the wrapper swaps of
/root/reference/src/blockmatrix.jl:150-160, src/symmetricblockmatrix.jl:197-237, 307-365 (block / offdiagonal /
diagonal return the lazy adjoint or transpose, rowindices <-> colindices swap, diagonalindices do not)."""
import numpy as np
import pytest

import bsm_b200 as B
from oracle import oracle_np as O


@pytest.fixture(scope="module", params=["cuboid", "sphere"])
def fixture(request):
    return O.load_golden_sbm(request.param)


def test_blockmatrix_accessor_battery(fixture):
    A = O.sbm_to_bsm(fixture)
    b = B.BlockSparseMatrix(A.blocks, A.rowindices, A.colindices, A.size)
    bsparse = O.sparse_bsm(A)
    # test_blockmatrix.jl:84-91
    assert B.nnz(b) == B.nnz(B.adjoint(b)) == B.nnz(B.transpose(b)) == bsparse.nnz
    # test_blockmatrix.jl:93-98
    assert list(B.eachblockindex(b)) == list(B.eachblockindex(B.adjoint(b))) == list(B.eachblockindex(B.transpose(b))) \
        == list(range(1, len(A.blocks) + 1))
    # test_blockmatrix.jl:100-106 and the swaps of src/blockmatrix.jl:150-160, src/symmetricblockmatrix.jl:341-365
    for i in B.eachblockindex(b):
        blk = B.block(b, i)
        assert blk.dtype == B.block(B.adjoint(b), i).dtype == B.block(B.transpose(b), i).dtype == np.complex128
        assert np.array_equal(B.block(B.adjoint(b), i), blk.conj().T)
        assert np.array_equal(B.block(B.transpose(b), i), blk.T)
        assert np.array_equal(B.rowindices(B.adjoint(b), i), B.colindices(b, i))
        assert np.array_equal(B.colindices(B.adjoint(b), i), B.rowindices(b, i))
        assert np.array_equal(B.rowindices(B.transpose(b), i), B.colindices(b, i))
        assert np.array_equal(B.colindices(B.transpose(b), i), B.rowindices(b, i))
        assert blk.shape == (len(B.rowindices(b, i)), len(B.colindices(b, i)))
    assert B.size(b) == A.size and B.size(B.adjoint(b)) == A.size[::-1] and B.eltype(b) == np.complex128
    # double wrapping returns the parent (LinearMaps: adjoint(adjoint(A)) === A)
    assert B.adjoint(B.adjoint(b)) is b and B.transpose(B.transpose(b)) is b


def test_symmetricblockmatrix_accessor_battery(fixture):
    A = fixture
    s = B.SymmetricBlockMatrix(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    ssparse = O.sparse_sbm(A)
    # test_symmetricblockmatrix.jl:99-106: nnz counts the half-stored blocks twice, under every wrapper
    assert B.nnz(s) == B.nnz(B.adjoint(s)) == B.nnz(B.transpose(s)) == ssparse.nnz
    assert B.nnz(s) == sum(d.size for d in A.diagonals) + 2 * sum(o.size for o in A.offdiagonals)
    assert list(B.eachoffdiagonalindex(s)) == list(B.eachoffdiagonalindex(B.adjoint(s))) == list(range(1, len(A.offdiagonals) + 1))
    assert list(B.eachdiagonalindex(s)) == list(B.eachdiagonalindex(B.transpose(s))) == list(range(1, len(A.diagonals) + 1))
    for i in B.eachoffdiagonalindex(s):
        o = B.offdiagonal(s, i)
        assert np.array_equal(B.offdiagonal(B.adjoint(s), i), o.conj().T)          # src/symmetricblockmatrix.jl:219-233
        assert np.array_equal(B.offdiagonal(B.transpose(s), i), o.T)
        assert np.array_equal(B.rowindices(B.adjoint(s), i), B.colindices(s, i))    # :341-365
        assert np.array_equal(B.colindices(B.transpose(s), i), B.rowindices(s, i))
        assert o.shape == (len(B.rowindices(s, i)), len(B.colindices(s, i)))
    for i in B.eachdiagonalindex(s):
        d = B.diagonal(s, i)
        assert np.array_equal(B.diagonal(B.adjoint(s), i), d.conj().T)
        assert np.array_equal(B.diagonal(B.transpose(s), i), d.T)
        # diagonal index vectors are shared by rows and columns: no swap (src/symmetricblockmatrix.jl:327-339)
        assert np.array_equal(B.diagonalindices(B.adjoint(s), i), B.diagonalindices(s, i))
        assert d.shape == (len(B.diagonalindices(s, i)),) * 2
    # the symmetric matrix equals its transpose; its adjoint is the conjugate (src/symmetricblockmatrix.jl:386-435)
    assert abs(ssparse - ssparse.T).max() == 0
    assert abs(O.sparse_sbm(A, "C") - ssparse.conj()).max() == 0
