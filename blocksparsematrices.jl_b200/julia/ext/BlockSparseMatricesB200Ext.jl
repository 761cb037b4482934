# BlockSparseMatricesB200Ext.jl — package extension that binds libbsm_b200.so (include/bsm_b200.h)
# behind BlockSparseMatrices.jl's own operator API. Modelled on ext/BlockUnicodePlots of the reference
# (Project.toml [weakdeps]/[extensions]); triggered by loading the carrier package `BSMB200`
# (blocksparsematrices.jl_b200/julia/BSMB200: locates the shared library and owns the user-facing names
# `B200`, `update!`, `setvariant!`, `cg`; see INTEGRATION.md).
#
# NOT RUN IN THIS REPOSITORY'S CI: the build image has no Julia. The file is kept thin on purpose —
# marshalling only; every numerical statement is tested through the same C ABI from tests/ (ctypes).
#
# No CUDA.jl kernels, no CPU fallback: a B200Matrix multiplies only through ccall.
module BlockSparseMatricesB200Ext

using BlockSparseMatrices
using BlockSparseMatrices: AbstractBlockMatrix, BlockSparseMatrix, SymmetricBlockMatrix,
                           VariableBlockCompressedRowStorage
using LinearAlgebra, LinearMaps, SparseArrays
import BSMB200
import BSMB200: libbsm_b200          # const libbsm_b200 = "/path/to/libbsm_b200.so"

const BSM_DTYPE = Dict(Float32 => Cint(0), Float64 => Cint(1), ComplexF64 => Cint(2))
const OP_N, OP_T, OP_C = Cint(0), Cint(1), Cint(2)

struct BsmOptions                     # mirrors bsm_options
    device::Int32
    variant::Int32
    own_row_lo::Int64
    own_row_hi::Int64
    own_col_lo::Int64
    own_col_hi::Int64
    plan_hints::Int64
    blocks_on_device::Int64
    reserved::NTuple{2,Int64}
end
BsmOptions(; device=-1, variant=0) = BsmOptions(device, variant, 0, -1, 0, -1, 0, 0, (0, 0))

check(rc) = rc == 0 || error("libbsm_b200: " * unsafe_string(ccall((:bsm_last_error, libbsm_b200), Cstring, ())))

"""
    B200Matrix(A)   # A::BlockSparseMatrix | SymmetricBlockMatrix | VariableBlockCompressedRowStorage

Device-resident copy of `A` (arena + index tables in HBM). `<: AbstractBlockMatrix{T}`, so `*`, `mul!`,
`adjoint`, `transpose`, `A[:, :]` and Krylov solvers work through LinearMaps unchanged.
"""
mutable struct B200Matrix{T} <: AbstractBlockMatrix{T}
    handle::Ptr{Cvoid}
    size::Tuple{Int,Int}
    function B200Matrix{T}(h, sz) where {T}
        A = new{T}(h, sz)
        finalizer(a -> ccall((:bsm_destroy, libbsm_b200), Cint, (Ptr{Cvoid},), a.handle), A)
        return A
    end
end

pool(vs) = (reduce(vcat, vs; init=Int64[]), Int64[0; cumsum(length.(vs))])
colmajor(T, b) = b isa Matrix{T} ? b : Matrix{T}(b)     # materialises lazy wrappers

function B200Matrix(A::BlockSparseMatrix{T}; kw...) where {T}
    blocks = [colmajor(T, b) for b in A.blocks]
    ptrs = Ptr{Cvoid}[pointer(b) for b in blocks]
    m, n = Int64.(size.(blocks, 1)), Int64.(size.(blocks, 2))
    ri, rp = pool(A.rowindices); ci, cp = pool(A.colindices)
    h = Ref{Ptr{Cvoid}}(C_NULL); opt = Ref(BsmOptions(; kw...))
    GC.@preserve blocks check(ccall((:bsm_create_blocksparse, libbsm_b200), Cint,
        (Cint, Int64, Int64, Int64, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
         Ptr{Int64}, Ptr{Int64}, Ref{BsmOptions}, Ref{Ptr{Cvoid}}),
        BSM_DTYPE[T], A.size[1], A.size[2], length(blocks), ptrs, m, n, ri, rp, ci, cp, opt, h))
    return B200Matrix{T}(h[], A.size)
end

function B200Matrix(A::SymmetricBlockMatrix{T}; kw...) where {T}
    D = [colmajor(T, b) for b in A.diagonals]; O = [colmajor(T, b) for b in A.offdiagonals]
    dp = Ptr{Cvoid}[pointer(b) for b in D]; op = Ptr{Cvoid}[pointer(b) for b in O]
    di, dptr = pool(A.diagonalindices); ri, rp = pool(A.rowindices); ci, cp = pool(A.colindices)
    h = Ref{Ptr{Cvoid}}(C_NULL); opt = Ref(BsmOptions(; kw...))
    GC.@preserve D O check(ccall((:bsm_create_symmetric, libbsm_b200), Cint,
        (Cint, Int64, Int64, Int64, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64,
         Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
         Ref{BsmOptions}, Ref{Ptr{Cvoid}}),
        BSM_DTYPE[T], A.size[1], A.size[2], length(D), dp, Int64.(size.(D, 1)), di, dptr, length(O), op,
        Int64.(size.(O, 1)), Int64.(size.(O, 2)), ri, rp, ci, cp, opt, h))
    return B200Matrix{T}(h[], A.size)
end

function B200Matrix(A::VariableBlockCompressedRowStorage{T}; kw...) where {T}
    # blocks may be lazy `transpose(parent)` wrappers (SymmetricBlockMatrix → VBCRS conversion,
    # src/vbcrs.jl:222-264): hand the parent over and let the packer materialise the transpose
    istr = UInt8[b isa Transpose ? 1 : 0 for b in A.blocks]
    keep = [b isa Transpose ? colmajor(T, parent(b)) : colmajor(T, b) for b in A.blocks]
    ptrs = Ptr{Cvoid}[pointer(b) for b in keep]
    m, n = Int64.(size.(A.blocks, 1)), Int64.(size.(A.blocks, 2))
    h = Ref{Ptr{Cvoid}}(C_NULL); opt = Ref(BsmOptions(; kw...))
    GC.@preserve keep check(ccall((:bsm_create_vbcrs, libbsm_b200), Cint,
        (Cint, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Int64},
         Ptr{Int64}, Ptr{UInt8}, Ref{BsmOptions}, Ref{Ptr{Cvoid}}),
        BSM_DTYPE[T], A.size[1], A.size[2], length(A.rowptr) - 1, length(keep), Int64.(A.rowptr),
        Int64.(A.colindices), Int64.(A.rowindices), ptrs, m, n, istr, opt, h))
    return B200Matrix{T}(h[], A.size)
end

SparseArrays.nnz(A::B200Matrix) = Int(ccall((:bsm_nnz, libbsm_b200), Int64, (Ptr{Cvoid},), A.handle))
SparseArrays.nnz(A::Union{LinearMaps.AdjointMap{<:Any,<:B200Matrix},LinearMaps.TransposeMap{<:Any,<:B200Matrix}}) = nnz(A.lmap)

opcode(::B200Matrix) = OP_N
opcode(::LinearMaps.TransposeMap{<:Any,<:B200Matrix}) = OP_T
opcode(::LinearMaps.AdjointMap{<:Any,<:B200Matrix}) = OP_C
parentmap(A::B200Matrix) = A
parentmap(A) = A.lmap

const B200Map{T} = Union{B200Matrix{T},LinearMaps.AdjointMap{T,<:B200Matrix{T}},LinearMaps.TransposeMap{T,<:B200Matrix{T}}}

# the user-facing constructor lives in the carrier package
BSMB200.B200(A::Union{BlockSparseMatrix,SymmetricBlockMatrix,VariableBlockCompressedRowStorage}; kw...) = B200Matrix(A; kw...)

# host Arrays: bsm_mul_host copies x in and y out. β === false is Julia's strong zero
# (src/abstractblockmatrix.jl:33) and is passed as beta_is_false = 1.
function mulhost!(y::StridedVecOrMat{T}, A::B200Map{T}, x::StridedVecOrMat{T}, α::Number, β::Number) where {T}
    size(x, 2) == size(y, 2) || throw(DimensionMismatch("x and y must have the same number of columns"))
    P = parentmap(A)
    a, b = Ref(T(α)), Ref(T(β))
    GC.@preserve x y check(ccall((:bsm_mul_host, libbsm_b200), Cint,
        (Ptr{Cvoid}, Cint, Ref{T}, Ref{T}, Cint, Ptr{T}, Int64, Ptr{T}, Int64, Int64),
        P.handle, opcode(A), a, b, β === false, x, max(stride(x, 2), 1), y, max(stride(y, 2), 1), size(x, 2)))
    return y
end
# anything that is not a dense column-major array of T (views with gaps, mixed element types — ComplexF64 blocks
# times Float64 x as in test/test_vbcrs.jl:34-35) goes through dense copies
dense(::Type{T}, v::StridedVecOrMat{T}) where {T} = stride(v, 1) == 1 ? v : copy(v)
dense(::Type{T}, v::AbstractVecOrMat) where {T} = convert(Array{T}, v)
function mulany!(y, A::B200Map{T}, x, α, β) where {T}
    yd = dense(T, y)
    mulhost!(yd, A, dense(T, x), α, β)
    yd === y || copyto!(y, yd)
    return y
end

# Dispatch. The reference's 3-arg method is
#     _unsafe_mul!(y::AbstractVector, A::M, x::AbstractVector) where {Z<:AbstractBlockMatrix, M<:Union{Z,AdjointMap{<:Any,Z},TransposeMap{<:Any,Z}}}
# (src/abstractblockmatrix.jl:27-34). Every method below has EXACTLY the same y / x types and a strictly more
# specific A, so it is strictly more specific than the reference's (and than LinearMaps' generic matrix methods):
# no ambiguity, and `A * x`, `mul!(y, A, x)`, `mul!(Y, A, X, α, β)` all land here.
LinearMaps._unsafe_mul!(y::AbstractVector, A::B200Map, x::AbstractVector) = mulany!(y, A, x, true, false)
LinearMaps._unsafe_mul!(y::AbstractVector, A::B200Map, x::AbstractVector, α::Number, β::Number) = mulany!(y, A, x, α, β)
LinearMaps._unsafe_mul!(y::AbstractMatrix, A::B200Map, x::AbstractMatrix) = mulany!(y, A, x, true, false)
LinearMaps._unsafe_mul!(y::AbstractMatrix, A::B200Map, x::AbstractMatrix, α::Number, β::Number) = mulany!(y, A, x, α, β)

# SparseArrays.sparse(A) built on the device from the resident arena (src/sparse.jl:127-129): canonical CSC,
# 1-based Int64 colptr / rowval exactly as SparseMatrixCSC holds them.
function SparseArrays.sparse(A::B200Map{T}) where {T}
    P = parentmap(A)
    nnzref = Ref{Int64}(0)
    check(ccall((:bsm_sparse_build, libbsm_b200), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), P.handle, opcode(A), nnzref))
    m, n = size(A)
    colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnzref[]); nzval = Vector{T}(undef, nnzref[])
    check(ccall((:bsm_sparse_fetch, libbsm_b200), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{T}),
                P.handle, colptr, rowval, nzval))
    return SparseMatrixCSC(m, n, colptr, rowval, nzval)
end

# new values, same structure: re-upload without re-planning (blocks in creation order)
function BSMB200.update!(B::B200Matrix{T}, blocks::Vector{<:AbstractMatrix}) where {T}
    keep = [colmajor(T, b) for b in blocks]
    ptrs = Ptr{Cvoid}[pointer(b) for b in keep]
    GC.@preserve keep check(ccall((:bsm_update_values, libbsm_b200), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64),
                                  B.handle, ptrs, length(ptrs)))
    return B
end

# kernel variant for comparisons: 0 auto (stream), 1 gather, 2 direct loads, 3 colour-ordered (the reference's schedule)
BSMB200.setvariant!(B::B200Matrix, v::Integer) = (check(ccall((:bsm_set_variant, libbsm_b200), Cint, (Ptr{Cvoid}, Cint), B.handle, v)); B)

# solver loop kept on the device (bsm_cg): b is copied in, x copied out; the iteration itself never leaves the GPU
struct BsmCgOptions
    rtol::Float64
    maxit::Int64
    hermitian::Int32
    check_every::Int32
end
function BSMB200.cg(A::B200Matrix{T}, b::AbstractVector; rtol=1e-10, maxit=200, hermitian=false) where {T}
    n = size(A, 1)
    length(b) == n || throw(DimensionMismatch("b has length $(length(b)), operator needs $n"))
    bh = convert(Vector{T}, b); xh = Vector{T}(undef, n)
    dev = Ref{Ptr{Cvoid}}(C_NULL); xdev = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bsm_malloc, libbsm_b200), Cint, (Cint, Csize_t, Ref{Ptr{Cvoid}}), -1, sizeof(T) * n, dev))
    check(ccall((:bsm_malloc, libbsm_b200), Cint, (Cint, Csize_t, Ref{Ptr{Cvoid}}), -1, sizeof(T) * n, xdev))
    iters = Ref{Int64}(0); relres = Ref{Float64}(0.0)
    try
        check(ccall((:bsm_memcpy_h2d, libbsm_b200), Cint, (Ptr{Cvoid}, Ptr{T}, Csize_t), dev[], bh, sizeof(T) * n))
        check(ccall((:bsm_cg, libbsm_b200), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{BsmCgOptions}, Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
            A.handle, dev[], xdev[], Ref(BsmCgOptions(rtol, maxit, hermitian, 8)), iters, relres, C_NULL))
        check(ccall((:bsm_memcpy_d2h, libbsm_b200), Cint, (Ptr{T}, Ptr{Cvoid}, Csize_t), xh, xdev[], sizeof(T) * n))
    finally
        ccall((:bsm_free, libbsm_b200), Cint, (Cint, Ptr{Cvoid}), -1, dev[])
        ccall((:bsm_free, libbsm_b200), Cint, (Cint, Ptr{Cvoid}), -1, xdev[])
    end
    return xh, iters[], relres[]
end

end # module
