// spmm.cuh — multi-RHS block-sparse x dense product (SpMM) on the FP64 tensor cores (DMMA).
//
//   Y[R_b, :] (+)= op(B_b) X[C_b, :]      for nrhs >= 8 right-hand sides, Float64
//
// With many right-hand sides the product is a real dense contraction (2*nrhs flop per 8 bytes of A),
// so the inner product runs on `mma.sync.aligned.m8n8k4.f64` (SASS DMMA) instead of the column loop
// LinearMaps uses for the reference (/root/reference/src/abstractblockmatrix.jl:27-34 applied per column).
//
// One persistent CTA per work item (a run of short output segments = block rows), 64 right-hand sides per
// pass (blockIdx.y). Warps 8..11 are producers: for every 32-wide slab of the contraction they copy, with
// cp.async (LDGSTS, register-free, completes on the stage's mbarrier), the block slab and the matching
// 32 x 64 tile of X into a 3-stage shared-memory ring — both in PADDED layouts (leading dimension = 4 mod 8
// doubles) so that every DMMA fragment load of the 8 consumer warps is bank-conflict free, which a TMA
// bulk copy of the column-major block cannot give for 32-row blocks. A 64-byte header per stage carries
// what the consumers need (sizes, form, segment boundaries, output rows), so they never touch the tables.
// Consumers: warp (wm, wn) owns the M-tiles of half of the segment's rows x 16 right-hand sides; the
// accumulators (<= 4 x 2 m8n8 tiles) stay in registers across all blocks of a segment and are written
// once (alpha/beta fused). Deterministic: fixed order, no atomics.
#pragma once
#include "kernels.cuh"

namespace bsm {

constexpr int kMConsWarps = 8;
constexpr int kMProdWarps = 4;             // one producer warp cannot issue 24 KB of cp.async per stage fast enough
constexpr int kMThreads = (kMConsWarps + kMProdWarps) * 32;
constexpr int kMMaxStages = 4;
constexpr int kMSmemBudget = 112 * 1024;    // per CTA, two CTAs per SM
constexpr int kMKc = 32;                    // contraction slab per stage
constexpr int kMRhs = 64;                   // right-hand sides per pass
constexpr int kMLd = 36;                    // leading dimension of the X tile and of T-form block slabs
constexpr int kMABytesBig = 64 * kMLd * 8;    // block slab area: max(68 x 32, 36 x 64) doubles
constexpr int kMABytesSmall = 32 * kMLd * 8;  // when no block has more than 32 rows / columns
constexpr int kMXBytes = kMRhs * kMLd * 8;    // X tile area
constexpr int kMHdrBytes = 64;
constexpr int kSpmmMinRhs = 8;

struct SpmmArgs {
    const double *arena;
    const bsm_contrib *contrib;
    const bsm_slice *slices;
    const int32_t *item_ptr;
    const int32_t *set_start;
    const int64_t *set_pool_off;
    const int32_t *pool;
    const double *x;
    double *y;
    int64_t ldx, ldy;
    double alpha, beta;
    int32_t nrhs;
    int32_t beta_false;
    int32_t abytes;     // block slab area of one stage (kMABytesSmall or kMABytesBig)
    int32_t nstages;    // stages of the ring (3 or 4)
};

struct alignas(16) SpmmHdr {
    int32_t kcv;        // valid contraction entries in this stage (<= 32)
    int32_t mo;         // outputs the block covers (rows for N-form, columns for T-form)
    int32_t ldA;        // N-form: leading dimension of the slab (element (o,k) at k*ldA+o); T-form: (o,k) at o*36+k
    int32_t flags;      // bit0 T-form, bit1 first chunk of a segment, bit2 last chunk of a segment, bit3 last chunk of the item
    int32_t L;          // rows of the segment
    int32_t out_start;  // first output row if the segment is a contiguous range, else -1
    int64_t out_pool;   // pool offset of the segment's row indices when out_start < 0
};

__device__ __forceinline__ int spmm_ldpad(int m) { return m + ((12 - (m & 7)) & 7); }  // smallest >= m with ld % 8 == 4


// Copies `nrun` runs of `len` consecutive doubles each, global -> shared, with the 32 lanes of the
// producer warp: run r goes from src + r*sstride to dst + r*dstride. wide: 16-byte pieces (len even,
// both sides 16-byte aligned), else 8-byte pieces. No integer division: short runs are dealt to
// power-of-two lane groups.
__device__ __forceinline__ void spmm_copy_runs(double *dst, int32_t dstride, const double *src, int64_t sstride,
                                               int32_t nrun, int32_t len, bool wide, int lane, int pw) {
    const int32_t per = wide ? (len >> 1) : len;   // pieces per run
    if (per <= 0) return;
    if (per <= 16) {
        const int sh = per <= 1 ? 0 : 32 - __clz(per - 1);   // log2 of the lane-group width
        const int32_t p = lane & ((1 << sh) - 1), rpi = (32 >> sh) * kMProdWarps;
        int32_t r = (lane >> sh) + pw * (32 >> sh);
        if (p < per && r < nrun) {
            const int32_t e = wide ? 2 * p : p;
            double *d = dst + r * dstride + e;
            const double *g = src + r * sstride + e;
            const int32_t dinc = rpi * dstride;
            const int64_t ginc = rpi * sstride;
            if (wide) {
#pragma unroll 4
                for (; r < nrun; r += rpi, d += dinc, g += ginc) cp_async_elem<16>(d, g);
            } else {
#pragma unroll 4
                for (; r < nrun; r += rpi, d += dinc, g += ginc) cp_async_elem<8>(d, g);
            }
        }
    } else {
        for (int32_t r = pw; r < nrun; r += kMProdWarps)
            for (int32_t p = lane; p < per; p += 32) {
                if (wide)
                    cp_async_elem<16>(dst + r * dstride + 2 * p, src + r * sstride + 2 * p);
                else
                    cp_async_elem<8>(dst + r * dstride + p, src + r * sstride + p);
            }
    }
}

__device__ __forceinline__ void spmm_producer(const SpmmArgs &a, unsigned char *smem, uint64_t *full, uint64_t *empty,
                                              int32_t s0, int32_t s1, int32_t j0, int32_t ncol) {
    const int lane = threadIdx.x & 31;
    const int pw = (threadIdx.x >> 5) - kMConsWarps;   // producer warp index
    const uint32_t nst = (uint32_t)a.nstages;
    const int32_t stage_bytes = a.abytes + kMXBytes + kMHdrBytes;
    uint32_t q = 0, stage = 0, round = 0;
    for (int32_t si = s0; si < s1; ++si) {
        const bsm_slice sl = a.slices[si];
        const int32_t L = sl.r1;
        const int32_t ostart = __ldg(a.set_start + sl.out_set);
        const int64_t opool = __ldg(a.set_pool_off + sl.out_set);
        // last contribution that carries data (the segment ends with its last chunk)
        int32_t clast = sl.c_begin;
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
            const int32_t K = tform ? m : n;
            const int32_t xs0 = __ldg(a.set_start + cb.in_set);
            const int64_t xpool = __ldg(a.set_pool_off + cb.in_set);
            const double *blk = a.arena + cb.off;
            const int32_t ldA = spmm_ldpad(m);
            for (int32_t k0 = 0; k0 < K; k0 += kMKc, ++q) {
                const int32_t kcv = min(kMKc, K - k0);
                if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1);
                unsigned char *sb = smem + stage * stage_bytes;
                double *As = reinterpret_cast<double *>(sb);
                double *Xs = reinterpret_cast<double *>(sb + a.abytes);
                // ---- block slab
                if (!tform)   // columns k0 .. k0+kcv, m rows each, to As[kk*ldA + r]
                    spmm_copy_runs(As, ldA, blk + (int64_t)k0 * m, m, kcv, m, (m & 1) == 0, lane, pw);
                else          // rows k0 .. k0+kcv of every column i, to As[i*36 + kk]
                    spmm_copy_runs(As, kMLd, blk + k0, m, n, kcv, ((m | kcv) & 1) == 0, lane, pw);
                // ---- X tile: rows (in-set positions k0 .. k0+kcv) x ncol right-hand sides, to Xs[j*36 + kk]
                if (xs0 >= 0) {
                    const int64_t xrow = (int64_t)xs0 + k0;
                    const bool wide = (((xrow | a.ldx) & 1) == 0) && ((kcv & 1) == 0) &&
                                      ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
                    spmm_copy_runs(Xs, kMLd, a.x + (int64_t)j0 * a.ldx + xrow, a.ldx, ncol, kcv, wide, lane, pw);
                } else {
                    for (int32_t kk = lane; kk < kcv; kk += 32) {
                        const int64_t row = __ldg(a.pool + xpool + k0 + kk);
                        for (int32_t j = pw; j < ncol; j += kMProdWarps)
                            cp_async_elem<8>(Xs + j * kMLd + kk, a.x + (int64_t)(j0 + j) * a.ldx + row);
                    }
                }
                cp_async_mbar_arrive_noinc(&full[stage]);
                if (lane == 0 && pw == 0) {
                    SpmmHdr h;
                    h.kcv = kcv;
                    h.mo = cb.out_len;
                    h.ldA = ldA;
                    const bool seg_end = (ci == clast) && (k0 + kMKc >= K);
                    h.flags = (tform ? 1 : 0) | (first ? 2 : 0) | (seg_end ? 4 : 0) | ((seg_end && si == s1 - 1) ? 8 : 0);
                    h.L = L;
                    h.out_start = ostart;
                    h.out_pool = opool;
                    *reinterpret_cast<SpmmHdr *>(sb + a.abytes + kMXBytes) = h;
                    mbar_arrive(&full[stage]);   // release: the header is visible to whoever observes the phase
                }
                first = false;
                if (++stage == nst) {
                    stage = 0;
                    ++round;
                }
            }
        }
    }
}

// One stage for one consumer warp: NT M-tiles (8 rows each, starting at tile t0) x 2 N-tiles (16 right-hand
// sides). Everything that does not depend on k is hoisted; the k loop is LDS + DMMA only.
template <int NT, bool TF>
__device__ __forceinline__ void spmm_stage_compute(const double *__restrict__ As, const double *__restrict__ Xs,
                                                   int32_t kcv, int32_t mo, int32_t ldA, int32_t t0, int g, int tg,
                                                   int wn, double (&acc)[4][2][2]) {
    const double *xb0 = Xs + (16 * wn + g) * kMLd + tg;
    const double *xb1 = xb0 + 8 * kMLd;
    const double *ap[NT];
    bool ok[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int32_t o = 8 * (t0 + t) + g;
        ok[t] = o < mo;
        ap[t] = TF ? As + o * kMLd + tg : As + tg * ldA + o;
    }
    const int32_t kstep = TF ? 4 : 4 * ldA;
    const int32_t nfull = kcv >> 2;
#pragma unroll 2
    for (int32_t k4 = 0; k4 < nfull; ++k4) {
        const double b0 = xb0[4 * k4], b1 = xb1[4 * k4];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            double av = 0.0;
            if (ok[t]) av = ap[t][k4 * kstep];
            dmma_m8n8k4(acc[t][0][0], acc[t][0][1], av, b0);
            dmma_m8n8k4(acc[t][1][0], acc[t][1][1], av, b1);
        }
    }
    if (kcv & 3) {   // masked tail: garbage past the slab must not reach the accumulators
        const bool kv = (4 * nfull + tg) < kcv;
        const double b0 = kv ? xb0[4 * nfull] : 0.0, b1 = kv ? xb1[4 * nfull] : 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            double av = 0.0;
            if (kv && ok[t]) av = ap[t][nfull * kstep];
            dmma_m8n8k4(acc[t][0][0], acc[t][0][1], av, b0);
            dmma_m8n8k4(acc[t][1][0], acc[t][1][1], av, b1);
        }
    }
}

__device__ __forceinline__ void spmm_consumer(const SpmmArgs &a, unsigned char *smem, uint64_t *full, uint64_t *empty,
                                              int32_t j0, int32_t ncol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, tg = lane & 3;
    double acc[4][2][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 2; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
    const uint32_t nst = (uint32_t)a.nstages;
    const int32_t stage_bytes = a.abytes + kMXBytes + kMHdrBytes;
    const unsigned char *sb = smem;
    for (uint32_t stage = 0, round = 0;;) {
        mbar_wait(&full[stage], round & 1);
        const double *As = reinterpret_cast<const double *>(sb);
        const double *Xs = reinterpret_cast<const double *>(sb + a.abytes);
        const SpmmHdr h = *reinterpret_cast<const SpmmHdr *>(sb + a.abytes + kMXBytes);
        const int32_t Mt = (h.L + 7) >> 3, MH = (Mt + 1) >> 1;
        const int32_t t0 = wm * MH;                    // first M-tile of this warp
        const int32_t nt = min(MH, Mt - t0);           // its M-tiles (<= 4, may be <= 0)
        if (h.flags & 2) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 2; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
        }
        if (h.flags & 1) {
            switch (nt) {
            case 1: spmm_stage_compute<1, true>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            case 2: spmm_stage_compute<2, true>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            case 3: spmm_stage_compute<3, true>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            case 4: spmm_stage_compute<4, true>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            default: break;
            }
        } else {
            switch (nt) {
            case 1: spmm_stage_compute<1, false>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            case 2: spmm_stage_compute<2, false>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            case 3: spmm_stage_compute<3, false>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            case 4: spmm_stage_compute<4, false>(As, Xs, h.kcv, h.mo, h.ldA, t0, g, tg, wn, acc); break;
            default: break;
            }
        }
        if (h.flags & 4) {
            // segment complete: y = alpha*acc + beta*y, every element written exactly once
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t < nt) {
                    const int32_t o = 8 * (t0 + t) + g;
                    if (o < h.L) {
                        const int64_t row = h.out_start >= 0 ? (int64_t)h.out_start + o
                                                             : (int64_t)__ldg(a.pool + h.out_pool + o);
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int32_t j = 16 * wn + 8 * u + 2 * tg + e;
                                if (j < ncol) {
                                    double *yp = a.y + (int64_t)(j0 + j) * a.ldy + row;
                                    double v = a.alpha * acc[t][u][e];
                                    if (!a.beta_false) v += a.beta * (*yp);
                                    *yp = v;
                                }
                            }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (h.flags & 8) break;
        sb += stage_bytes;
        if (++stage == nst) {
            stage = 0;
            sb = smem;
            ++round;
        }
    }
}

__global__ void __launch_bounds__(kMThreads, 2) spmm_dmma_kernel(const SpmmArgs a) {
    extern __shared__ __align__(128) unsigned char msm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(msm + a.nstages * (a.abytes + kMXBytes + kMHdrBytes));
    uint64_t *empty = full + kMMaxStages;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nstages; ++i) {
            mbar_init(&full[i], kMProdWarps * 32 + 1);   // cp.async completions of every producer lane + the header release
            mbar_init(&empty[i], kMConsWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int32_t s0 = __ldg(a.item_ptr + blockIdx.x), s1 = __ldg(a.item_ptr + blockIdx.x + 1);
    if (s0 >= s1) return;
    const int32_t j0 = blockIdx.y * kMRhs;
    const int32_t ncol = min(kMRhs, a.nrhs - j0);
    if ((threadIdx.x >> 5) >= kMConsWarps)
        spmm_producer(a, msm, full, empty, s0, s1, j0, ncol);
    else
        spmm_consumer(a, msm, full, empty, j0, ncol);
}

// rows no block touches: y <- beta*y for every right-hand side (blockIdx.y)
__global__ void __launch_bounds__(256) spmm_uncovered_kernel(const int32_t *rows, int64_t n, double *y, int64_t ldy,
                                                             double beta, int beta_false) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *p = y + (int64_t)blockIdx.y * ldy + rows[i];
    *p = beta_false ? 0.0 : beta * (*p);
}

inline int spmm_stage_bytes(bool small) { return (small ? kMABytesSmall : kMABytesBig) + kMXBytes + kMHdrBytes; }
inline int spmm_stages(bool small) {
    const int n = (kMSmemBudget - 2 * kMMaxStages * 8) / spmm_stage_bytes(small);
    return n > kMMaxStages ? kMMaxStages : n;
}
inline size_t spmm_smem_bytes(bool small) {
    return (size_t)spmm_stages(small) * spmm_stage_bytes(small) + 2 * kMMaxStages * 8;
}

}  // namespace bsm
