# BSMB200 — carrier package of libbsm_b200.so (include/bsm_b200.h). Loading it next to BlockSparseMatrices triggers
# the package extension BlockSparseMatricesB200Ext, which gives the function stubs below their methods:
#
#     using BlockSparseMatrices, BSMB200
#     Ad = B200(A)                  # A::BlockSparseMatrix | SymmetricBlockMatrix | VariableBlockCompressedRowStorage
#     y  = Ad * x;  mul!(y, Ad', x, α, β);  nnz(Ad);  sparse(Ad)
#
# The carrier has no dependency of its own: it locates the shared library and owns the names a user types.
# NOT RUN IN THIS REPOSITORY'S CI (no Julia in the build image).
module BSMB200

export B200, update!, setvariant!, cg

"Path of libbsm_b200.so: ENV[\"BSM_B200_LIB\"] or next to this package."
const libbsm_b200 = get(ENV, "BSM_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "libbsm_b200.so"))

"""
    B200(A; device=-1, variant=0)

Device-resident copy of the block matrix `A` (arena + index tables in HBM, plans built once). The result is an
`AbstractBlockMatrix{T}`: `*`, `mul!`, `adjoint`, `transpose`, `A[:, :]`, `nnz`, `sparse` and Krylov solvers work
through LinearMaps unchanged. Methods are added by BlockSparseMatricesB200Ext.
"""
function B200 end

"`update!(Ad, blocks)`: new block values, same structure — re-upload without re-planning."
function update! end

"`setvariant!(Ad, v)`: kernel variant for comparison runs (0 auto, 1 gather, 2 direct loads, 3 colour-ordered)."
function setvariant! end

"`cg(Ad, b; rtol, maxit, hermitian)`: CG / COCG kept on the device (bsm_cg); returns `(x, iterations, relres)`."
function cg end

end # module
