// spmm_tma.cuh — multi-RHS block-sparse x dense product (SpMM) for Float32 / Float64 / ComplexF64 on the FP64
// tensor cores (DMMA), operands staged by the TMA engine.
//
//   Y[R_b, :] (+)= op(B_b) X[C_b, :]          nrhs >= 8 right-hand sides, blocks of <= 32 rows
//
// replaces the column loop LinearMaps applies for a matrix right-hand side
// (/root/reference/src/abstractblockmatrix.jl:27-34 per column) by ONE pass over A.
//
// One CTA = 4 consumer warps + 1 producer warp; persistent over a work item (a run of block rows = output segments
// of <= 32 rows); NB right-hand sides per pass (blockIdx.y). Per stage (one block, or one 32-wide contraction slab
// of a wide block) the elected producer lane issues
//   * ONE cp.async.bulk.tensor.2d (TMA, SASS UTMALDG) for the block slab: the arena is described as a 2-D tensor
//     of 128-byte rows, so the slab — contiguous bytes, 128-byte aligned — lands in shared memory with the
//     128-byte swizzle applied to its LINEAR byte offset;
//   * 32*s/128 TMA boxes (16 doubles x NB columns, 128-byte swizzle) for the tile of X the block multiplies
//     (rows of X beyond the matrix are zero-filled by the hardware, columns beyond nrhs too);
// both complete on the stage's "full" mbarrier (complete_tx); a 64-byte header written by the producer carries
// sizes, segment boundaries and output rows. The swizzle is what makes the DMMA fragment loads of the consumers
// bank-conflict free WITHOUT the padded layouts (and the 1536 cp.async per stage) of the round-1 kernel: for
// 32-row Float64 blocks every A-fragment load (lanes = 8 rows x 4 contraction steps) and, with the N-tile columns
// dealt to the lanes in the order 0,2,4,6,1,3,5,7, every B-fragment load hits 16 distinct banks per half-warp.
// Consumers: warp (wm, wn) owns M-tiles {wm, wm+WM, ..} x 16 (or 8) right-hand sides; accumulators stay in
// registers across all blocks of a segment and are written once (alpha/beta fused). ComplexF64 runs 4 real DMMAs
// per complex tile product on interleaved (re, im) fragments; Float32 operands are widened to Float64 in the
// fragment loads (the result is rounded to Float32 once). Deterministic: fixed order, no atomics.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

#define BSM_HD __host__ __device__ __forceinline__
#include "spmm_layout.h"

namespace bsm {

constexpr int kTConsWarps = 4;
constexpr int kTThreads = (kTConsWarps + 1) * 32;
constexpr int kTKc = 32;          // contraction slab per stage
constexpr int kTMaxM = 32;        // tallest block / longest segment this kernel takes
constexpr int kTMaxStages = 4;

struct alignas(16) TmaHdr {
    int32_t kcv;        // valid contraction entries in this stage (<= 32)
    int32_t mo;         // outputs the block covers (rows for N-form, columns for T-form)
    int32_t ld;         // rows of the stored block (leading dimension of the slab)
    int32_t flags;      // bit0 T-form, bit1 first stage of a segment, bit2 last stage of a segment, bit3 last stage of the item;
                        // bits 8..15: offset of the first contraction entry inside the X tile (alignment shift);
                        // bits 16..23: T-form: first contraction entry (block row) of this stage inside the staged block
    int32_t L;          // rows of the segment
    int32_t out_start;  // first output row if the segment is a contiguous range, else -1
    int64_t out_pool;   // pool offset of the segment's row indices when out_start < 0
};

template <class T>
struct TmaSpmmArgs {
    const bsm_contrib *contrib;
    const bsm_slice *slices;
    const int32_t *item_ptr;
    const int32_t *set_start;
    const int64_t *set_pool_off;
    const int32_t *pool;
    T *y;
    int64_t ldy;
    T alpha, beta;
    int32_t nrhs;
    int32_t beta_false;
    int32_t conj;
    int32_t nstages;
    int32_t elem_shift;     // log2(sizeof(T))
};

__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int32_t c0, int32_t c1,
                                            uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

template <class T>
struct TmaGeom {
    static constexpr int S = (int)sizeof(T);
    static constexpr int ABytes = kTMaxM * kTKc * S;          // block slab area (whole 32 x 32 slab)
    static constexpr int ARows = ABytes / 128;                // 128-byte rows of the largest arena box
    // X area: boxes of 128/S contraction entries x NB columns. The TMA engine needs the global start of a box 16-byte
    // aligned, so a tile whose first row is not (Float32: xs0 % 4, Float64: xs0 % 2) is fetched from the aligned row
    // below it and indexed with an offset: Float32 keeps a second box for that, Float64 halves the contraction slab
    static constexpr int XBoxes = S == 4 ? 2 : kTKc / (128 / S);
    template <int NB>
    __host__ __device__ static constexpr int XBytes() { return XBoxes * NB * 128; }
    template <int NB>
    __host__ __device__ static constexpr int StageBytes() { return ABytes + XBytes<NB>(); }   // multiples of 1 KB: stages stay 1 KB aligned
};

// Tensor maps of one launch (a single __grid_constant__ parameter, so that the kernel can address them):
// a[q]: the arena as rows of 128 bytes, box = 128 bytes x (ARows >> q) rows — the smallest box that covers a slab is
// used; x: the right-hand sides, box = 128 bytes of one column's contraction entries x NB columns.
struct alignas(64) TmaMaps {
    CUtensorMap a[4];
    CUtensorMap x;
};

// ---- producer --------------------------------------------------------------------------------------------------
template <class T, int NB>
__device__ __forceinline__ void spmm_tma_producer(const TmaSpmmArgs<T> &a, const CUtensorMap *amap, const CUtensorMap *xmap,
                                                  unsigned char *smem, uint64_t *full, uint64_t *empty, TmaHdr *hdrs,
                                                  int32_t s0, int32_t s1, int32_t j0) {
    using G = TmaGeom<T>;
    constexpr int S = G::S;
    constexpr int KB = 128 / S;                      // contraction entries per X box
    constexpr int XE = S == 16 ? 2 : 1;              // tensor-map elements per T (ComplexF64 = 2 doubles)
    const uint64_t pol_a = l2_evict_first_policy();  // A is streamed once
    const uint64_t pol_x = l2_evict_last_policy();   // X tiles are what L2 should keep
    const uint32_t nst = (uint32_t)a.nstages;
    uint32_t stage = 0, round = 0;
    for (int32_t si = s0; si < s1; ++si) {
        const bsm_slice sl = a.slices[si];
        const int32_t L = sl.r1;
        const int32_t ostart = __ldg(a.set_start + sl.out_set);
        const int64_t opool = __ldg(a.set_pool_off + sl.out_set);
        int32_t clast = sl.c_begin;       // last contribution that carries data (the segment ends with its last stage)
        for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
            const bsm_contrib cb = a.contrib[ci];
            if (cb.m > 0 && cb.n > 0) clast = ci;
        }
        bool first = true;
        for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
            const bsm_contrib cb = a.contrib[ci];
            if (cb.m == 0 || cb.n == 0) continue;
            const bool tform = (cb.form & 1) != 0;
            const int32_t m = cb.m, n = cb.n;
            const int32_t K = tform ? m : n;          // T-form: m <= 32, one stage holds the whole block
            const int32_t xs0 = __ldg(a.set_start + cb.in_set);
            const int32_t xd = S == 16 ? 0 : (S == 8 ? (xs0 & 1) : (xs0 & 3));      // alignment shift of the X tile
            const int32_t kstep = (S == 8 && xd) ? kTKc / 2 : kTKc;
            for (int32_t k0 = 0; k0 < K; k0 += kstep) {
                const int32_t kcv = min(kstep, K - k0);
                if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1);
                unsigned char *sb = smem + stage * G::template StageBytes<NB>();
                // slab: N-form columns k0 .. k0+kcv (contiguous; 16 or 32 columns of m elements start 128-byte
                // aligned); T-form the whole block (m <= 32), of which this stage contracts rows k0 .. k0+kcv
                const int64_t boff = (cb.off + (tform ? 0 : (int64_t)k0 * m)) << a.elem_shift;      // bytes, multiple of 128
                const int32_t slab = (tform ? m * n : m * kcv) << a.elem_shift;
                const int32_t rows = (slab + 127) >> 7;
                int q = 0;
                while (q < 3 && (G::ARows >> (q + 1)) >= rows) ++q;
                const uint32_t abytes = (uint32_t)(G::ARows >> q) * 128u;
                TmaHdr h;
                h.kcv = kcv;
                h.mo = cb.out_len;
                h.ld = m;
                const bool seg_end = (ci == clast) && (k0 + kstep >= K);
                h.flags = (tform ? 1 : 0) | (first ? 2 : 0) | (seg_end ? 4 : 0) | ((seg_end && si == s1 - 1) ? 8 : 0) |
                          (xd << 8) | ((tform ? k0 : 0) << 16);
                h.L = L;
                h.out_start = ostart;
                h.out_pool = opool;
                hdrs[stage] = h;
                // only the X boxes that hold valid contraction entries are fetched (the consumers mask the rest)
                const int32_t nbox = (kcv + xd + KB - 1) / KB;
                mbar_arrive_expect_tx(&full[stage], abytes + (uint32_t)nbox * (uint32_t)(NB * 128));
                tma_load_2d(sb, amap + q, 0, (int32_t)(boff >> 7), &full[stage], pol_a);
#pragma unroll
                for (int b = 0; b < G::XBoxes; ++b)
                    if (b < nbox)
                        tma_load_2d(sb + G::ABytes + b * (NB * 128), xmap, (xs0 + k0 - xd + b * KB) * XE, j0, &full[stage], pol_x);
                first = false;
                if (++stage == nst) {
                    stage = 0;
                    ++round;
                }
            }
        }
    }
}

// ---- consumers -------------------------------------------------------------------------------------------------
template <class T>
struct FragLoad;
template <>
struct FragLoad<double> {
    using V = double;
    static __device__ __forceinline__ V ld(const unsigned char *base, int idx) { return reinterpret_cast<const double *>(base)[idx]; }
    static __device__ __forceinline__ V zero() { return 0.0; }
};
template <>
struct FragLoad<float> {
    using V = double;
    static __device__ __forceinline__ V ld(const unsigned char *base, int idx) { return (double)reinterpret_cast<const float *>(base)[idx]; }
    static __device__ __forceinline__ V zero() { return 0.0; }
};
template <>
struct FragLoad<cplx> {
    using V = cplx;
    static __device__ __forceinline__ V ld(const unsigned char *base, int idx) { return reinterpret_cast<const cplx *>(base)[idx]; }
    static __device__ __forceinline__ V zero() { return cplx{0.0, 0.0}; }
};

// accumulator of one m8n8 tile: real: {c0, c1}; complex: {re0, re1, im0, im1}
template <class T>
struct TileAcc {
    double v[sizeof(T) == 16 ? 4 : 2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < (int)(sizeof(T) == 16 ? 4 : 2); ++i) v[i] = 0.0;
    }
};
__device__ __forceinline__ void tile_mma(TileAcc<double> &c, double a, double b) { dmma_m8n8k4(c.v[0], c.v[1], a, b); }
__device__ __forceinline__ void tile_mma(TileAcc<float> &c, double a, double b) { dmma_m8n8k4(c.v[0], c.v[1], a, b); }
__device__ __forceinline__ void tile_mma(TileAcc<cplx> &c, cplx a, cplx b) {
    dmma_m8n8k4(c.v[0], c.v[1], a.re, b.re);
    dmma_m8n8k4(c.v[0], c.v[1], -a.im, b.im);
    dmma_m8n8k4(c.v[2], c.v[3], a.re, b.im);
    dmma_m8n8k4(c.v[2], c.v[3], a.im, b.re);
}

template <class T, int NB>
struct TmaLayout {
    static constexpr int NT = NB >= 16 ? 2 : 1;                 // N-tiles (8 columns) per warp
    static constexpr int WN = NB >= 16 ? NB / 16 : 1;           // warps along the right-hand sides
    static constexpr int WM = kTConsWarps / WN;                 // warps along the segment rows
    static constexpr int MT = (kTMaxM / 8 + WM - 1) / WM;       // M-tiles per warp (tiles wm, wm+WM, ...)
    static_assert(WN * WM == kTConsWarps, "warp layout");
};

// One stage for one consumer warp. FULL: a whole 32 x 32 slab of a 32-row segment (kcv = 32, mo = L = 32) — no
// masks, the contraction loop fully unrolled so that the swizzled offsets fold into immediates. Otherwise every
// fragment load is masked (rows past the block, contraction entries past the slab: stale shared memory must not
// reach the accumulators) with selects, never with branches (the DMMAs need the warp converged).
template <class T, int NB, bool TF, bool FULL>
__device__ __forceinline__ void spmm_tma_stage(const unsigned char *As, const unsigned char *Xs, const TmaHdr &h, int conj,
                                               int wm, int wn, int g, int tg, TileAcc<T> (&acc)[TmaLayout<T, NB>::MT][TmaLayout<T, NB>::NT]) {
    using LY = TmaLayout<T, NB>;
    using FL = FragLoad<T>;
    constexpr int S = (int)sizeof(T);
    const int32_t Mt = FULL ? kTMaxM / 8 : (h.L + 7) >> 3;
    const int32_t m = FULL ? kTMaxM : h.ld;
    const int32_t xd = FULL ? 0 : (h.flags >> 8) & 0xff, kbase = FULL ? 0 : (h.flags >> 16) & 0xff;
    int jcol[LY::NT];
#pragma unroll
    for (int u = 0; u < LY::NT; ++u) jcol[u] = (LY::NT * 8) * wn + 8 * u + ntile_col<S>(g);
    if constexpr (FULL) {
#pragma unroll
        for (int32_t k4 = 0; k4 < kTKc / 4; ++k4) {
            const int32_t k = 4 * k4 + tg;
            typename FL::V b[LY::NT];
#pragma unroll
            for (int u = 0; u < LY::NT; ++u) b[u] = FL::ld(Xs, xtile_index<S, NB>(k, jcol[u]));
#pragma unroll
            for (int i = 0; i < LY::MT; ++i) {
                const int32_t o = 8 * (wm + LY::WM * i) + g;
                typename FL::V av = FL::ld(As, swz128<S>(TF ? o * m + k : k * m + o));
                if constexpr (S == 16) {
                    if (conj) av.im = -av.im;
                }
#pragma unroll
                for (int u = 0; u < LY::NT; ++u) tile_mma(acc[i][u], av, b[u]);
            }
        }
    } else {
        const int32_t nk4 = (h.kcv + 3) >> 2;
#pragma unroll 2
        for (int32_t k4 = 0; k4 < nk4; ++k4) {
            const int32_t k = 4 * k4 + tg;
            const bool kv = k < h.kcv;
            typename FL::V b[LY::NT];
#pragma unroll
            for (int u = 0; u < LY::NT; ++u) {
                b[u] = FL::ld(Xs, xtile_index<S, NB>(k + xd, jcol[u]));     // k + xd < entries of the X area
                if (!kv) b[u] = FL::zero();
            }
#pragma unroll
            for (int i = 0; i < LY::MT; ++i) {
                const int32_t t = wm + LY::WM * i;
                if (t < Mt) {      // warp-uniform
                    const int32_t o = 8 * t + g;
                    const bool ok = kv && o < h.mo;
                    typename FL::V av = FL::ld(As, ok ? swz128<S>(TF ? o * m + kbase + k : k * m + o) : 0);
                    if (!ok) av = FL::zero();
                    if constexpr (S == 16) {
                        if (conj) av.im = -av.im;
                    }
#pragma unroll
                    for (int u = 0; u < LY::NT; ++u) tile_mma(acc[i][u], av, b[u]);
                }
            }
        }
    }
}

template <class T>
__device__ __forceinline__ void spmm_store(T *yp, const double *v, int e, T alpha, T beta, int beta_false);
template <>
__device__ __forceinline__ void spmm_store<double>(double *yp, const double *v, int e, double alpha, double beta, int beta_false) {
    double r = alpha * v[e];
    if (!beta_false) r += beta * (*yp);
    __stcs(yp, r);
}
template <>
__device__ __forceinline__ void spmm_store<float>(float *yp, const double *v, int e, float alpha, float beta, int beta_false) {
    double r = (double)alpha * v[e];
    if (!beta_false) r += (double)beta * (double)(*yp);
    __stcs(yp, (float)r);
}
template <>
__device__ __forceinline__ void spmm_store<cplx>(cplx *yp, const double *v, int e, cplx alpha, cplx beta, int beta_false) {
    const cplx acc{v[e], v[2 + e]};
    cplx r = El<cplx>::mul(alpha, acc);
    if (!beta_false) r = El<cplx>::add(r, El<cplx>::mul(beta, *yp));
    __stcs(reinterpret_cast<double2 *>(yp), make_double2(r.re, r.im));
}

template <class T, int NB>
__device__ __forceinline__ void spmm_tma_consumer(const TmaSpmmArgs<T> &a, unsigned char *smem, uint64_t *full, uint64_t *empty,
                                                  const TmaHdr *hdrs, int32_t j0, int32_t ncol) {
    using G = TmaGeom<T>;
    using LY = TmaLayout<T, NB>;
    constexpr int S = (int)sizeof(T);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wn = warp % LY::WN, wm = warp / LY::WN;
    const int g = lane >> 2, tg = lane & 3;
    TileAcc<T> acc[LY::MT][LY::NT];
#pragma unroll
    for (int i = 0; i < LY::MT; ++i)
#pragma unroll
        for (int u = 0; u < LY::NT; ++u) acc[i][u].clear();
    const uint32_t nst = (uint32_t)a.nstages;
    for (uint32_t stage = 0, round = 0;;) {
        mbar_wait(&full[stage], round & 1);
        const unsigned char *sb = smem + stage * G::template StageBytes<NB>();
        const unsigned char *As = sb;
        const unsigned char *Xs = sb + G::ABytes;
        const TmaHdr h = hdrs[stage];
        if (h.flags & 2) {
#pragma unroll
            for (int i = 0; i < LY::MT; ++i)
#pragma unroll
                for (int u = 0; u < LY::NT; ++u) acc[i][u].clear();
        }
        const bool full = h.kcv == kTKc && h.mo == kTMaxM && h.L == kTMaxM && h.ld == kTMaxM && (h.flags >> 8) == 0;
        if (h.flags & 1) {
            if (full)
                spmm_tma_stage<T, NB, true, true>(As, Xs, h, a.conj, wm, wn, g, tg, acc);
            else
                spmm_tma_stage<T, NB, true, false>(As, Xs, h, a.conj, wm, wn, g, tg, acc);
        } else {
            if (full)
                spmm_tma_stage<T, NB, false, true>(As, Xs, h, a.conj, wm, wn, g, tg, acc);
            else
                spmm_tma_stage<T, NB, false, false>(As, Xs, h, a.conj, wm, wn, g, tg, acc);
        }
        if (h.flags & 4) {
            // segment complete: y = alpha*acc + beta*y, every element written exactly once
            const int32_t Mt = (h.L + 7) >> 3;
#pragma unroll
            for (int i = 0; i < LY::MT; ++i) {
                const int32_t t = wm + LY::WM * i;
                const int32_t o = 8 * t + g;
                if (t < Mt && o < h.L) {
                    const int64_t row = h.out_start >= 0 ? (int64_t)h.out_start + o : (int64_t)__ldg(a.pool + h.out_pool + o);
#pragma unroll
                    for (int u = 0; u < LY::NT; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int32_t j = (LY::NT * 8) * wn + 8 * u + ntile_col<S>(2 * tg + e);
                            if (j < ncol)
                                spmm_store<T>(a.y + (int64_t)(j0 + j) * a.ldy + row, acc[i][u].v, e, a.alpha, a.beta, a.beta_false);
                        }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (h.flags & 8) break;
        if (++stage == nst) {
            stage = 0;
            ++round;
        }
    }
}

template <class T, int NB>
__global__ void __launch_bounds__(kTThreads) spmm_tma_kernel(const TmaSpmmArgs<T> a, const __grid_constant__ TmaMaps maps) {
    // 1 KB alignment: the swizzle pattern is a function of shared-memory address bits 7..9
    extern __shared__ __align__(1024) unsigned char tsm[];
    if (smem_u32(tsm) & 1023u) __trap();
    using G = TmaGeom<T>;
    uint64_t *full = reinterpret_cast<uint64_t *>(tsm + a.nstages * G::template StageBytes<NB>());
    uint64_t *empty = full + kTMaxStages;
    TmaHdr *hdrs = reinterpret_cast<TmaHdr *>(empty + kTMaxStages);
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nstages; ++i) {
            mbar_init(&full[i], 1);                 // the producer's arrive.expect_tx; the TMA bytes complete the phase
            mbar_init(&empty[i], kTConsWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int32_t s0 = __ldg(a.item_ptr + blockIdx.x), s1 = __ldg(a.item_ptr + blockIdx.x + 1);
    if (s0 >= s1) return;
    const int32_t j0 = blockIdx.y * NB;
    const int32_t ncol = min(NB, a.nrhs - j0);
    if ((threadIdx.x >> 5) >= kTConsWarps) {
        if (threadIdx.x == kTConsWarps * 32)
            spmm_tma_producer<T, NB>(a, maps.a, &maps.x, tsm, full, empty, hdrs, s0, s1, j0);
    } else {
        spmm_tma_consumer<T, NB>(a, tsm, full, empty, hdrs, j0, ncol);
    }
}

template <class T>
__global__ void __launch_bounds__(256) spmm_uncovered_kernel_t(const int32_t *rows, int64_t n, T *y, int64_t ldy, T beta,
                                                               int beta_false) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T *p = y + (int64_t)blockIdx.y * ldy + rows[i];
    *p = beta_false ? El<T>::zero() : El<T>::mul(beta, *p);
}

template <class T, int NB>
constexpr int spmm_tma_stages() {
    constexpr int st = TmaGeom<T>::template StageBytes<NB>();
    constexpr int n = (73 * 1024) / st;
    return n > kTMaxStages ? kTMaxStages : (n < 2 ? 2 : n);
}
template <class T, int NB>
constexpr size_t spmm_tma_smem_bytes() {
    return (size_t)spmm_tma_stages<T, NB>() * TmaGeom<T>::template StageBytes<NB>() + 2 * kTMaxStages * 8 +
           kTMaxStages * sizeof(TmaHdr) + 1024;   // + slack for the 1 KB alignment of the stage ring
}

}  // namespace bsm
