// spmm_layout.h — shared-memory index arithmetic of spmm_tma_kernel (spmm_tma.cuh). Kept free of CUDA types so that
// tests/spmm_layout_emulation.cpp can compile it with g++ and replay the TMA placement + DMMA fragment addressing
// on the CPU. BSM_HD is defined by the includer (__host__ __device__ __forceinline__ under nvcc, inline under g++).
#pragma once

namespace bsm {

// Physical element index inside a 128-byte-swizzled shared-memory tile from the LINEAR element index (tile base
// 1024-byte aligned): the 16-byte chunk index (bits 4..6 of the byte offset) is XORed with the 128-byte row index
// modulo 8 (bits 7..9). S = element size in bytes.
template <int S>
BSM_HD int swz128(int lin) {
    constexpr int SH = S == 16 ? 0 : S == 8 ? 1 : 2;     // log2(elements per 16-byte chunk)
    return lin ^ (((lin >> (SH + 3)) & 7) << SH);
}
// Column of an N-tile (8 right-hand sides) that lane group g (= lane / 4) feeds: the order that makes the B-fragment
// loads conflict free under the swizzle (Float64: half-warps of 64-bit loads; ComplexF64: quarter-warps of 128-bit loads).
template <int S>
BSM_HD int ntile_col(int g) {
    if (S == 8) return ((g & 3) << 1) | (g >> 2);                       // 0,2,4,6,1,3,5,7
    if (S == 16) return ((g & 1) << 2) | (g & 2) | (g >> 2);            // 0,4,2,6,1,5,3,7
    return g;
}
// Element index of X[k, j] (k = position in the contraction slab, j = column of the pass) in the stage's X area:
// boxes of KB = 128/S contraction entries x NB columns, each box swizzled on its own.
template <int S, int NB>
BSM_HD int xtile_index(int k, int j) {
    constexpr int KB = 128 / S;
    return (k / KB) * (NB * KB) + swz128<S>(j * KB + (k % KB));
}

}  // namespace bsm
