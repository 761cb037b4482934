"""blocksparsematrices.jl_b200 — B200-native multiply path of BlockSparseMatrices.jl.

Import it as `bsm_b200` (repo-root shim; the directory name is not a Python identifier).
Host containers mirror the reference API (host.py); products run on the GPU through the C ABI of
libbsm_b200.so (device.py, include/bsm_b200.h). No CPU fallback.
"""
from .host import (AbstractBlockMatrix, AdjointMap, BlockSparseMatrix, SymmetricBlockMatrix,
                   TransposeMap, VariableBlockCompressedRowStorage, adjoint, block, colindices,
                   diagonal, diagonalindices, eachblockindex, eachdiagonalindex,
                   eachoffdiagonalindex, eltype, mul_, nnz, offdiagonal, rowcolvals, rowindices,
                   size, sparse, transpose)
from .device import DeviceMatrix

__all__ = [
    "AbstractBlockMatrix", "AdjointMap", "TransposeMap", "BlockSparseMatrix", "SymmetricBlockMatrix",
    "VariableBlockCompressedRowStorage", "DeviceMatrix", "adjoint", "transpose", "mul_", "nnz", "size",
    "eltype", "eachblockindex", "block", "rowindices", "colindices", "offdiagonal", "diagonal",
    "diagonalindices", "eachoffdiagonalindex", "eachdiagonalindex", "rowcolvals", "sparse",
]
