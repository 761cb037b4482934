// kernels.cuh — sm_100a kernels of the block-sparse multiply.
//
//   gather_gemv_kernel<T>  one CTA (128 threads) per work item ("slice"): owner-computes gather
//                          GEMV over all contributions of an output segment, direct coalesced
//                          global loads of the column-major blocks (128-bit when the leading
//                          dimension allows it), x segments staged in shared memory, warp-shuffle
//                          reductions for the T-form (dot-product) contributions, alpha/beta fused
//                          into the single write of every owned y row.
//   gather_finalize_kernel<T>  reduces scratch partial vectors through the per-row gather lists in a
//                          fixed order (deterministic, no atomics).
//
// Replaces the per-colour fork/join + per-block LinearAlgebra.mul! of
// /root/reference/src/blockmatrix.jl:231-244, src/symmetricblockmatrix.jl:392-432 and
// src/vbcrs.jl:273-286, 313-326.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/bsm_b200.h"

namespace bsm {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kXsCap = 1024;  // x-segment staging capacity (elements)

struct alignas(16) cplx {
    double re, im;
};

// ---- element arithmetic ----------------------------------------------------------------------
template <class T>
struct El;
template <>
struct El<float> {
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float conj(float a) { return a; }
    static __device__ __forceinline__ void fma(float &acc, float a, float b) { acc = fmaf(a, b, acc); }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float shfl_xor(float a, int m) { return __shfl_xor_sync(0xffffffffu, a, m); }
};
template <>
struct El<double> {
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double conj(double a) { return a; }
    static __device__ __forceinline__ void fma(double &acc, double a, double b) { acc = ::fma(a, b, acc); }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double shfl_xor(double a, int m) { return __shfl_xor_sync(0xffffffffu, a, m); }
};
template <>
struct El<cplx> {
    static __device__ __forceinline__ cplx zero() { return cplx{0.0, 0.0}; }
    static __device__ __forceinline__ cplx conj(cplx a) { return cplx{a.re, -a.im}; }
    static __device__ __forceinline__ void fma(cplx &acc, cplx a, cplx b) {
        acc.re = ::fma(a.re, b.re, acc.re);
        acc.re = ::fma(-a.im, b.im, acc.re);
        acc.im = ::fma(a.re, b.im, acc.im);
        acc.im = ::fma(a.im, b.re, acc.im);
    }
    static __device__ __forceinline__ cplx mul(cplx a, cplx b) {
        return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
    }
    static __device__ __forceinline__ cplx add(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
    static __device__ __forceinline__ cplx shfl_xor(cplx a, int m) {
        return cplx{__shfl_xor_sync(0xffffffffu, a.re, m), __shfl_xor_sync(0xffffffffu, a.im, m)};
    }
};

// ---- streaming (evict-first) loads of V consecutive elements ---------------------------------
template <class T, int V>
struct Pack {
    T v[V];
};

template <class T, int V>
__device__ __forceinline__ Pack<T, V> load_stream(const T *p) {
    Pack<T, V> r;
    constexpr int bytes = (int)sizeof(T) * V;
    if constexpr (bytes == 16) {
        const int4 q = __ldcs(reinterpret_cast<const int4 *>(p));
        *reinterpret_cast<int4 *>(&r) = q;
    } else if constexpr (bytes == 8) {
        const int2 q = __ldcs(reinterpret_cast<const int2 *>(p));
        *reinterpret_cast<int2 *>(&r) = q;
    } else {
        static_assert(bytes == 4, "unsupported pack");
        const int q = __ldcs(reinterpret_cast<const int *>(p));
        *reinterpret_cast<int *>(&r) = q;
    }
    return r;
}

// ---- where x lives ---------------------------------------------------------------------------------
// Single GPU: one array. Slab handles in peer mode (bsm_mul_dist_peer): every rank keeps a full-length x in
// peer-mapped memory but only its own slab [cuts[r], cuts[r+1]) is current; element i is read straight from
// its owner's array over NVLink (npeer > 0), so the all-gather disappears into the kernels' own x fetches.
constexpr int kMaxPeers = 8;

// Peer-mode barriers folded into the multiply kernels (no extra launches). Every rank owns a flag array in
// peer-mapped memory: ready[nranks] then done[nranks], written by the peers with the epoch (= index of the
// multiply) they have reached. `state` are four words in LOCAL device memory: [0] epochs completed, [1] epoch whose
// "ready" signal has been sent, [2] arrival counter of the exit barrier, [3] epoch whose entry barrier has been passed.
//   entry  (every kernel of the multiply that reads x; one thread per CTA / warp, before the first x read):
//          the first arrival tells every peer "my x slab of this epoch is written" (the kernel runs after the
//          stream work that wrote it), then everybody waits until every peer has said so;
//   exit   (the last x-reading kernel of the multiply; one thread per CTA / warp, after its last x read):
//          the last arrival tells every peer "I have finished reading", waits until every peer has said so and
//          publishes the epoch — the kernel cannot complete earlier, so whatever follows on the stream (the
//          solver's update of the slab) is ordered after the peers' reads.
struct PeerSync {
    int32_t *const *peer_flags;
    const int32_t *my_flags;
    int32_t *state;
    int32_t nranks, rank;
    int32_t do_exit;      // this kernel runs the exit barrier
    int32_t arrivals;     // arrivals the exit barrier expects
    int32_t debug;        // benchmarking only (bsm_dist_set_debug): bit0 skip the entry wait, bit1 skip the exit wait,
                          // bit2 every arrival does the system-scope wait itself, bit3 record %globaltimer stamps in dbg
                          // (sums over the multiplies: [0] entry wait ns, [1] kernel start -> last arrival ns,
                          // [2] exit wait ns, [3] count), bit4 signal with release instead of relaxed stores
    long long *dbg;
};

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t *p) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int32_t *p, int32_t v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// waits until *flag >= epoch (acquire: what the writer published before its release store is visible afterwards)
__device__ __forceinline__ void peer_spin(const int32_t *flag, int32_t epoch) {
    if (ld_acquire_sys(flag) >= epoch) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(64);
        if (clock64() - t0 > 120000000000ll) __trap();   // ~60 s: a peer died or never entered the collective call
    }
}
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t *p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t *p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Thousands of CTAs / warps pass here; a system-scope acquire by each of them costs ~50 us per multiply (measured on
// C3, 2 GPUs). So the system-scope wait on the peers' flags is done by whoever arrives while state[3] ("entry passed")
// is still behind; it then publishes the epoch in state[3] with a device-scope release, and everybody else gets by
// with ONE device-scope acquire of that local word (the synchronisation chain peer -> first arrival -> this thread
// is transitive).
// Signals are RELAXED system-scope stores by default. A release store (fence.acq_rel.sys + st) costs ~10 us per
// barrier on 8 GPUs (measured: C3 0.114 -> 0.094 ms per multiply) and buys nothing here: "my x slab is written" is
// published by a kernel that runs AFTER the kernels that wrote the slab completed (stream order: their stores have
// reached this GPU's L2, the point of coherence every peer reads through), and "I have finished reading" follows
// loads whose data has already been consumed (every arrival passed a __threadfence and the arrival counter).
// bsm_dist_set_debug bit4 switches the release stores back on for comparison.
__device__ __forceinline__ void peer_signal(const PeerSync &s, int32_t *flag, int32_t e) {
    if (s.debug & 16)
        st_release_sys(flag, e);
    else
        asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(e) : "memory");
}
__device__ __forceinline__ void peer_entry(const PeerSync &s) {
    const int32_t e = *reinterpret_cast<volatile int32_t *>(s.state) + 1;
    if (*reinterpret_cast<volatile int32_t *>(s.state + 1) < e && atomicMax(s.state + 1, e) < e) {
        if (s.debug & 8) s.dbg[4] = global_ns();
        for (int p = 0; p < s.nranks; ++p) peer_signal(s, s.peer_flags[p] + s.rank, e);
    }
    if (s.debug & 1) return;
    if (!(s.debug & 4) && ld_acquire_gpu(s.state + 3) >= e) return;
    for (int p = 0; p < s.nranks; ++p) peer_spin(s.my_flags + p, e);
    if (!(s.debug & 4)) {
        if ((s.debug & 8) && atomicMax(s.state + 4, e) < e) s.dbg[5] = global_ns();
        st_release_gpu(s.state + 3, e);
    }
}
__device__ __forceinline__ void peer_exit(const PeerSync &s) {
    if (!s.do_exit) return;
    __threadfence();
    if (atomicAdd(s.state + 2, 1) != s.arrivals - 1) return;
    const int32_t e = *reinterpret_cast<volatile int32_t *>(s.state) + 1;
    const long long t2 = (s.debug & 8) ? global_ns() : 0;
    for (int p = 0; p < s.nranks; ++p) peer_signal(s, s.peer_flags[p] + s.nranks + s.rank, e);
    if (!(s.debug & 2))
        for (int p = 0; p < s.nranks; ++p) peer_spin(s.my_flags + s.nranks + p, e);
    if (s.debug & 8) {
        const long long t3 = global_ns();
        s.dbg[0] += s.dbg[5] - s.dbg[4];
        s.dbg[1] += t2 - s.dbg[4];
        s.dbg[2] += t3 - t2;
        s.dbg[3] += 1;
    }
    *reinterpret_cast<volatile int32_t *>(s.state + 2) = 0;
    *reinterpret_cast<volatile int32_t *>(s.state) = e;
    __threadfence();
}

template <class T>
struct XSrc {
    const T *x;
    const T *peer[kMaxPeers];
    int32_t cuts[kMaxPeers + 1];
    int32_t npeer;   // 0: plain array
    PeerSync sync;   // npeer > 0 only
    __device__ __forceinline__ const T *ptr(int32_t i) const {
        if (npeer == 0) return x + i;
        int r = 0;
#pragma unroll
        for (int k = 1; k < kMaxPeers; ++k) r += (k < npeer && i >= cuts[k]) ? 1 : 0;
        return peer[r] + i;
    }
    __device__ __forceinline__ T at(int32_t i) const { return *ptr(i); }
};

// the multiply of a rank whose slab holds no work: the barriers alone
static __global__ void peer_sync_kernel(const PeerSync s) {
    if (threadIdx.x == 0) {
        peer_entry(s);
        peer_exit(s);
    }
}

// ---- kernel arguments ------------------------------------------------------------------------
template <class T>
struct MulArgs {
    const T *arena;
    const bsm_contrib *contrib;
    const int64_t *contrib_toff;
    const bsm_slice *slices;
    const int32_t *set_len;
    const int32_t *set_start;
    const int64_t *set_pool_off;
    const int32_t *pool;
    XSrc<T> x;
    T *y;
    T *scratch;
    T alpha, beta;
    int32_t nslices;
    int32_t beta_false;
    int32_t conj;
    // GEN plans with row pieces of tall N-form blocks (real element types): contribution -> tensor map (-1: none), 128-byte
    // CUtensorMap records in global memory (abi.cu: ensure_piece_maps)
    const int32_t *contrib_map = nullptr;
    const unsigned char *piece_maps = nullptr;
};

struct SetRef {
    int32_t start;
    const int32_t *pool;
    __device__ __forceinline__ int32_t at(int32_t k) const { return start >= 0 ? start + k : __ldg(pool + k); }
};

template <class T>
__device__ __forceinline__ SetRef set_ref(const MulArgs<T> &a, int32_t s) {
    SetRef r;
    r.start = __ldg(a.set_start + s);
    r.pool = a.pool + __ldg(a.set_pool_off + s);
    return r;
}

// ---- N-form: outputs along the rows of the block ---------------------------------------------
// thread -> (row vector iv, column phase c); columns j = c, c+P, ...; acc[V] lives in registers
template <class T, int V, bool CONJ>
__device__ __forceinline__ void nform_chunk(const T *__restrict__ base, int32_t m, int32_t cn, int32_t c,
                                            int32_t P, const T *__restrict__ xs, T (&acc)[V]) {
    int32_t j = c;
    for (; j + 3 * P < cn; j += 4 * P) {
        Pack<T, V> v0 = load_stream<T, V>(base + (int64_t)j * m);
        Pack<T, V> v1 = load_stream<T, V>(base + (int64_t)(j + P) * m);
        Pack<T, V> v2 = load_stream<T, V>(base + (int64_t)(j + 2 * P) * m);
        Pack<T, V> v3 = load_stream<T, V>(base + (int64_t)(j + 3 * P) * m);
        const T x0 = xs[j], x1 = xs[j + P], x2 = xs[j + 2 * P], x3 = xs[j + 3 * P];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            El<T>::fma(acc[q], CONJ ? El<T>::conj(v0.v[q]) : v0.v[q], x0);
            El<T>::fma(acc[q], CONJ ? El<T>::conj(v1.v[q]) : v1.v[q], x1);
            El<T>::fma(acc[q], CONJ ? El<T>::conj(v2.v[q]) : v2.v[q], x2);
            El<T>::fma(acc[q], CONJ ? El<T>::conj(v3.v[q]) : v3.v[q], x3);
        }
    }
    for (; j < cn; j += P) {
        Pack<T, V> v0 = load_stream<T, V>(base + (int64_t)j * m);
        const T x0 = xs[j];
#pragma unroll
        for (int q = 0; q < V; ++q) El<T>::fma(acc[q], CONJ ? El<T>::conj(v0.v[q]) : v0.v[q], x0);
    }
}

// ---- T-form: outputs along the columns of the block ------------------------------------------
// one warp per column, lanes stride over the rows, shuffle reduction
template <class T, int V, bool CONJ>
__device__ __forceinline__ T tform_column(const T *__restrict__ col, int32_t cm, int lane,
                                          const T *__restrict__ xs) {
    T s0 = El<T>::zero(), s1 = El<T>::zero();
    int32_t i = lane * V;
    for (; i + 32 * V < cm; i += 64 * V) {
        Pack<T, V> v0 = load_stream<T, V>(col + i);
        Pack<T, V> v1 = load_stream<T, V>(col + i + 32 * V);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            El<T>::fma(s0, CONJ ? El<T>::conj(v0.v[q]) : v0.v[q], xs[i + q]);
            El<T>::fma(s1, CONJ ? El<T>::conj(v1.v[q]) : v1.v[q], xs[i + 32 * V + q]);
        }
    }
    if (i < cm) {
        Pack<T, V> v0 = load_stream<T, V>(col + i);
#pragma unroll
        for (int q = 0; q < V; ++q) El<T>::fma(s0, CONJ ? El<T>::conj(v0.v[q]) : v0.v[q], xs[i + q]);
    }
    T s = El<T>::add(s0, s1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = El<T>::add(s, El<T>::shfl_xor(s, o));
    return s;
}

template <class T, int V, bool CONJ>
__device__ __forceinline__ void slice_body(const MulArgs<T> &a, const bsm_slice &sl, T *xs, T *red,
                                           T *accT) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int32_t r0 = sl.r0, h = sl.r1 - sl.r0;
    const int32_t hv = (h + V - 1) / V;          // row vectors (h % V == 0 whenever V > 1)
    const int32_t P = kThreads / hv;             // column phases, >= 1
    const bool active = t < P * hv;
    const int32_t iv = t % hv, c = t / hv;
    T acc[V];
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = El<T>::zero();
    accT[t] = El<T>::zero();

    for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
        const bsm_contrib cb = a.contrib[ci];
        const SetRef in = set_ref(a, cb.in_set);
        const T *blk = a.arena + cb.off;
        if ((cb.form & 1) == 0) {
            // rows [r0, r0+h) ∩ [0, out_len) of every column
            const bool rows_ok = active && (r0 + iv * V) < cb.out_len;
            for (int32_t j0 = 0; j0 < cb.n; j0 += kXsCap) {
                const int32_t cn = min(kXsCap, cb.n - j0);
                __syncthreads();
                for (int32_t k = t; k < cn; k += kThreads) xs[k] = a.x.at(in.at(j0 + k));
                __syncthreads();
                if (rows_ok)
                    nform_chunk<T, V, CONJ>(blk + (int64_t)j0 * cb.m + r0 + iv * V, cb.m, cn, c, P, xs, acc);
            }
        } else {
            // columns [r0, r0+h) ∩ [0, out_len), dot products over all rows
            const int32_t hc = min(h, cb.out_len - r0);
            for (int32_t i0 = 0; i0 < cb.m; i0 += kXsCap) {
                const int32_t cm = min(kXsCap, cb.m - i0);
                __syncthreads();
                for (int32_t k = t; k < cm; k += kThreads) xs[k] = a.x.at(in.at(i0 + k));
                __syncthreads();
                for (int32_t jj = warp; jj < hc; jj += kWarps) {
                    const T s = tform_column<T, V, CONJ>(blk + (int64_t)(r0 + jj) * cb.m + i0, cm, lane, xs);
                    if (lane == 0) accT[jj] = El<T>::add(accT[jj], s);
                }
            }
        }
    }
    // reduce the column phases, add the T-form part, write
    __syncthreads();
    if (active) {
#pragma unroll
        for (int q = 0; q < V; ++q) red[c * h + iv * V + q] = acc[q];
    }
    __syncthreads();
    if (t < h) {
        T tot = accT[t];
        for (int32_t p = 0; p < P; ++p) tot = El<T>::add(tot, red[p * h + t]);
        if (sl.flags & 1) {
            const SetRef out = set_ref(a, sl.out_set);
            const int32_t row = out.at(r0 + t);
            T v = El<T>::mul(a.alpha, tot);
            if (!a.beta_false) v = El<T>::add(v, El<T>::mul(a.beta, a.y[row]));
            a.y[row] = v;
        } else {
            a.scratch[sl.scratch_off + t] = tot;
        }
    }
}

template <class T, int VMAX>
__global__ void __launch_bounds__(kThreads) gather_gemv_kernel(const MulArgs<T> a) {
    __shared__ __align__(16) unsigned char xs_raw[kXsCap * sizeof(T)];
    __shared__ __align__(16) unsigned char red_raw[kThreads * VMAX * sizeof(T)];
    __shared__ __align__(16) unsigned char acc_raw[kThreads * sizeof(T)];
    T *xs = reinterpret_cast<T *>(xs_raw);
    T *red = reinterpret_cast<T *>(red_raw);
    T *accT = reinterpret_cast<T *>(acc_raw);
    const bsm_slice sl = a.slices[blockIdx.x];
    const bool vec = (VMAX > 1) && (sl.flags & 2);
    if (a.x.npeer) {
        if (threadIdx.x == 0) peer_entry(a.x.sync);
        __syncthreads();
    }
    if (a.conj) {
        if (vec)
            slice_body<T, VMAX, true>(a, sl, xs, red, accT);
        else
            slice_body<T, 1, true>(a, sl, xs, red, accT);
    } else {
        if (vec)
            slice_body<T, VMAX, false>(a, sl, xs, red, accT);
        else
            slice_body<T, 1, false>(a, sl, xs, red, accT);
    }
    if (a.x.npeer && threadIdx.x == 0) peer_exit(a.x.sync);   // every x read precedes the last barrier of slice_body
}

// ---- fused symmetric kernel ---------------------------------------------------------------------
// One CTA (8 warps) per output segment of <= 256 rows, all of its contributions. Lanes own rows
// (RPL rows per lane, kept in registers), warps stride over column PAIRS. For a half-stored symmetric
// off-diagonal block O (rows R = this segment, columns C) ONE pass over the block yields
//     y[R]  += op(O) x[C]          accumulated in registers, written by this CTA
//     t[C]   = op(O)^T x[R]        complete per column inside the CTA (two-column shuffle butterfly),
//                                  stored to scratch and summed into y[C] by gather_finalize_kernel
// which replaces the two sweeps /root/reference/src/symmetricblockmatrix.jl:394-418 (each streaming all
// off-diagonal storage) and their colour barriers.
constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFMaxRows = 256;

template <class T, int RPL, bool CONJ>
__device__ __forceinline__ void fused_slice(const MulArgs<T> &a, const bsm_slice &sl, T *xs, T *xrs,
                                            T *red, T *accT) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int32_t L = sl.r1;  // r0 == 0: the whole segment
    const SetRef out = set_ref(a, sl.out_set);
    xrs[t] = (t < L) ? a.x.at(out.at(t)) : El<T>::zero();   // x at the segment's own rows
    accT[t] = El<T>::zero();
    T accN[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) accN[k] = El<T>::zero();

    for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
        const bsm_contrib cb = a.contrib[ci];
        const SetRef in = set_ref(a, cb.in_set);
        const T *blk = a.arena + cb.off;
        if ((cb.form & 1) == 0) {
            const bool fusedT = (cb.form & 2) != 0;
            const int64_t toff = fusedT ? a.contrib_toff[ci] : 0;
            const int32_t m = cb.m;
            for (int32_t j0 = 0; j0 < cb.n; j0 += kXsCap) {
                const int32_t cn = min(kXsCap, cb.n - j0);
                __syncthreads();
                for (int32_t k = t; k < cn; k += kFThreads) xs[k] = a.x.at(in.at(j0 + k));
                __syncthreads();
                for (int32_t j = 2 * warp; j < cn; j += 2 * kFWarps) {
                    const bool hasB = (j + 1) < cn;
                    const T *cA = blk + (int64_t)(j0 + j) * m;
                    const T *cB = cA + m;
                    T vA[RPL], vB[RPL];
#pragma unroll
                    for (int k = 0; k < RPL; ++k) {
                        const int32_t i = lane + 32 * k;
                        vA[k] = (i < m) ? load_stream<T, 1>(cA + i).v[0] : El<T>::zero();
                        vB[k] = (hasB && i < m) ? load_stream<T, 1>(cB + i).v[0] : El<T>::zero();
                    }
                    const T xA = xs[j];
                    const T xB = hasB ? xs[j + 1] : El<T>::zero();
                    T tA = El<T>::zero(), tB = El<T>::zero();
#pragma unroll
                    for (int k = 0; k < RPL; ++k) {
                        const T eA = CONJ ? El<T>::conj(vA[k]) : vA[k];
                        const T eB = CONJ ? El<T>::conj(vB[k]) : vB[k];
                        El<T>::fma(accN[k], eA, xA);
                        El<T>::fma(accN[k], eB, xB);
                        if (fusedT) {
                            const T xr = xrs[lane + 32 * k];
                            El<T>::fma(tA, eA, xr);
                            El<T>::fma(tB, eB, xr);
                        }
                    }
                    if (fusedT) {
                        // two-column butterfly: lanes < 16 end up with column A, lanes >= 16 with B
                        const bool hi = (lane & 16) != 0;
                        const T send = hi ? tA : tB;
                        T v = El<T>::add(hi ? tB : tA, El<T>::shfl_xor(send, 16));
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) v = El<T>::add(v, El<T>::shfl_xor(v, o));
                        if (lane == 0) a.scratch[toff + j0 + j] = v;
                        if (lane == 16 && hasB) a.scratch[toff + j0 + j + 1] = v;
                    }
                }
            }
        } else {
            // T-form contribution owned by this segment (transpose(D) / adjoint(D) of a diagonal block)
            const int32_t hc = min(L, cb.out_len);
            for (int32_t i0 = 0; i0 < cb.m; i0 += kXsCap) {
                const int32_t cm = min(kXsCap, cb.m - i0);
                __syncthreads();
                for (int32_t k = t; k < cm; k += kFThreads) xs[k] = a.x.at(in.at(i0 + k));
                __syncthreads();
                for (int32_t jj = warp; jj < hc; jj += kFWarps) {
                    const T s = tform_column<T, 1, CONJ>(blk + (int64_t)jj * cb.m + i0, cm, lane, xs);
                    if (lane == 0) accT[jj] = El<T>::add(accT[jj], s);
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPL; ++k) red[warp * kFMaxRows + k * 32 + lane] = accN[k];
    __syncthreads();
    if (t < L) {
        T tot = accT[t];
#pragma unroll
        for (int w = 0; w < kFWarps; ++w) tot = El<T>::add(tot, red[w * kFMaxRows + t]);
        if (sl.flags & 1) {
            const int32_t row = out.at(t);
            T v = El<T>::mul(a.alpha, tot);
            if (!a.beta_false) v = El<T>::add(v, El<T>::mul(a.beta, a.y[row]));
            a.y[row] = v;
        } else {
            a.scratch[sl.scratch_off + t] = tot;
        }
    }
}

template <class T>
__global__ void __launch_bounds__(kFThreads, 2) sym_fused_kernel(const MulArgs<T> a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    T *xs = reinterpret_cast<T *>(fsm);
    T *xrs = xs + kXsCap;
    T *accT = xrs + kFMaxRows;
    T *red = accT + kFMaxRows;  // [kFWarps][kFMaxRows]
    const bsm_slice sl = a.slices[blockIdx.x];
    const int32_t L = sl.r1;
    if (a.x.npeer) {
        if (threadIdx.x == 0) peer_entry(a.x.sync);
        __syncthreads();
    }
    if (a.conj) {
        if (L <= 64)
            fused_slice<T, 2, true>(a, sl, xs, xrs, red, accT);
        else if (L <= 128)
            fused_slice<T, 4, true>(a, sl, xs, xrs, red, accT);
        else
            fused_slice<T, 8, true>(a, sl, xs, xrs, red, accT);
    } else {
        if (L <= 64)
            fused_slice<T, 2, false>(a, sl, xs, xrs, red, accT);
        else if (L <= 128)
            fused_slice<T, 4, false>(a, sl, xs, xrs, red, accT);
        else
            fused_slice<T, 8, false>(a, sl, xs, xrs, red, accT);
    }
    if (a.x.npeer && threadIdx.x == 0) peer_exit(a.x.sync);
}

template <class T>
constexpr size_t fused_smem_bytes() {
    return sizeof(T) * (size_t)(kXsCap + 2 * kFMaxRows + kFWarps * kFMaxRows);
}

// ---- fused symmetric kernel, TMA-staged -------------------------------------------------------------
// Same work decomposition as sym_fused_kernel, but the block data never passes through registers on
// its way from HBM: one producer thread streams whole-column chunks of every block with
// cp.async.bulk (the TMA bulk-copy engine, SASS UBLKCP) into a ring of kPStages shared-memory stages;
// completion is signalled on "full" mbarriers (complete_tx), 8 consumer warps compute from shared
// memory and hand stages back through "empty" mbarriers. Loads in flight per SM = 2 CTAs x 4 stages x
// 20 KB, independent of register pressure and occupancy.
constexpr int kPStages = 4;
constexpr int kPChunk = 20480;                 // payload bytes per stage
constexpr int kPStageBytes = kPChunk + 128;    // + slack for the 16-byte alignment shift
constexpr int kPThreads = kFThreads + 32;      // 8 consumer warps + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (16-byte aligned addresses, size multiple of 16), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// global -> shared 2-D box through a tensor map (TMA tiled mode, SASS UTMALDG): the whole box counts towards the
// barrier's transaction bytes, elements outside the tensor arrive as zeros
__device__ __forceinline__ void tma_box_2d(void *dst_smem, const void *map, int32_t c0, int32_t c1, uint64_t *bar,
                                           uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <class T>
__device__ __forceinline__ int32_t chunk_cols(int32_t m) {
    const int32_t cc = kPChunk / (m * (int32_t)sizeof(T));
    return cc < 1 ? 1 : cc;
}

// One chunk (ncols whole columns of an m-row block) resident in shared memory.
//   doN : accN[k] += op(B)[i_k, j] * xcol[j]
//   doT : t[j] = sum_i op(B)[i, j] * xrow[i]  ->  tglobal[j] = t (fused partial, global scratch)
//                                              or tsmem[j] += t (T-form contribution owned by the segment)
template <class T, int RPL, bool CONJ>
__device__ __forceinline__ void consume_chunk(const T *__restrict__ sm, int32_t m, int32_t ncols, int lane,
                                              int warp, bool doN, bool doT, const T *__restrict__ xcol,
                                              const T *__restrict__ xrow, T (&accN)[RPL], T *tglobal, T *tsmem) {
    // `warp` arrives already rotated by the number of column pairs consumed so far (see tma_consumer): with
    // tall blocks a chunk holds fewer pairs than there are warps, and without the rotation warps 0..k would do
    // all the work of every chunk while the others idle
    for (int32_t jc = 2 * warp; jc < ncols; jc += 2 * kFWarps) {
        const bool hasB = (jc + 1) < ncols;
        const T *cA = sm + jc * m;
        const T *cB = cA + m;
        const T xA = doN ? xcol[jc] : El<T>::zero();
        const T xB = (doN && hasB) ? xcol[jc + 1] : El<T>::zero();
        T tA = El<T>::zero(), tB = El<T>::zero();
        // rows in groups of KG per lane: bounds the live registers of the 8-rows-per-lane case
        constexpr int KG = RPL < 4 ? RPL : 4;
#pragma unroll
        for (int kg = 0; kg < RPL; kg += KG) {
            T vA[KG], vB[KG];
#pragma unroll
            for (int k = 0; k < KG; ++k) {
                const int32_t i = lane + 32 * (kg + k);
                vA[k] = (i < m) ? cA[i] : El<T>::zero();
                vB[k] = (hasB && i < m) ? cB[i] : El<T>::zero();
                if (CONJ) {
                    vA[k] = El<T>::conj(vA[k]);
                    vB[k] = El<T>::conj(vB[k]);
                }
            }
            if (doN) {
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    El<T>::fma(accN[kg + k], vA[k], xA);
                    El<T>::fma(accN[kg + k], vB[k], xB);
                }
            }
            if (doT) {
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    const T xr = xrow[lane + 32 * (kg + k)];
                    El<T>::fma(tA, vA[k], xr);
                    El<T>::fma(tB, vB[k], xr);
                }
            }
        }
        if (doT) {
            const bool hi = (lane & 16) != 0;
            const T send = hi ? tA : tB;
            T v = El<T>::add(hi ? tB : tA, El<T>::shfl_xor(send, 16));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v = El<T>::add(v, El<T>::shfl_xor(v, o));
            if (tglobal) {
                if (lane == 0) tglobal[jc] = v;
                if (lane == 16 && hasB) tglobal[jc + 1] = v;
            } else {
                if (lane == 0) tsmem[jc] = El<T>::add(tsmem[jc], v);
                if (lane == 16 && hasB) tsmem[jc + 1] = El<T>::add(tsmem[jc + 1], v);
            }
        }
    }
}

// T-form chunk of a block of any height (m <= kXsCap): one warp per column, lanes stride over the rows,
// shuffle reduction. Columns are dealt to the warps by their ABSOLUTE index (j % 8), so that with only a
// few tall columns per stage the eight warps still work concurrently, each on a different stage.
template <class T, bool CONJ>
__device__ __forceinline__ void consume_chunk_tall(const T *__restrict__ sm, int32_t m, int32_t ncols,
                                                   int32_t jabs0, int lane, int warp,
                                                   const T *__restrict__ xrow, T *tsmem) {
    for (int32_t jc = (warp - jabs0) & (kFWarps - 1); jc < ncols; jc += kFWarps) {
        const T *col = sm + jc * m;
        T s0 = El<T>::zero(), s1 = El<T>::zero(), s2 = El<T>::zero(), s3 = El<T>::zero();
        int32_t i = lane;
        for (; i + 96 < m; i += 128) {
            const T v0 = col[i], v1 = col[i + 32], v2 = col[i + 64], v3 = col[i + 96];
            El<T>::fma(s0, CONJ ? El<T>::conj(v0) : v0, xrow[i]);
            El<T>::fma(s1, CONJ ? El<T>::conj(v1) : v1, xrow[i + 32]);
            El<T>::fma(s2, CONJ ? El<T>::conj(v2) : v2, xrow[i + 64]);
            El<T>::fma(s3, CONJ ? El<T>::conj(v3) : v3, xrow[i + 96]);
        }
        for (; i < m; i += 32) {
            const T v0 = col[i];
            El<T>::fma(s0, CONJ ? El<T>::conj(v0) : v0, xrow[i]);
        }
        T v = El<T>::add(El<T>::add(s0, s1), El<T>::add(s2, s3));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = El<T>::add(v, El<T>::shfl_xor(v, o));
        if (lane == 0) tsmem[jc] = El<T>::add(tsmem[jc], v);
    }
}

// GEN = false: whole segments whose T-form blocks are no taller than the rows the lanes hold (every
// SymmetricBlockMatrix plan); GEN = true adds column sub-ranges of long all-T-form segments and T-form
// blocks of up to kXsCap rows (consume_chunk_tall). Two instantiations, because the extra paths cost the
// lean one registers (and 7 % of its bandwidth on C2).
template <class T, int RPL, bool CONJ, bool GEN>
__device__ __forceinline__ void tma_consumer(const MulArgs<T> &a, const bsm_slice &sl, unsigned char *stages,
                                             T *xs, T *xrs, T *accT, uint64_t *full, uint64_t *empty) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;   // t < 256
    const int32_t r0 = GEN ? sl.r0 : 0;   // > 0 only for column sub-ranges of all-T-form segments
    const int32_t L = sl.r1 - r0;
    const SetRef out = set_ref(a, sl.out_set);
    // xrs is zero-padded to 256 rows and xs to the row count read below, so lanes past the block's
    // height multiply by zero
    xrs[t] = (r0 == 0 && t < L) ? a.x.at(out.at(t)) : El<T>::zero();
    accT[t] = El<T>::zero();
    T accN[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) accN[k] = El<T>::zero();
    uint32_t q = 0;
    uint32_t pbase = 0;   // column pairs consumed so far: rotates the pair -> warp assignment from chunk to chunk
    for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
        const bsm_contrib cb = a.contrib[ci];
        const SetRef in = set_ref(a, cb.in_set);
        const bool tform = (cb.form & 1) != 0;
        const bool fusedT = (cb.form & 2) != 0;
        // GEN: a 256-row piece of a LONG N-form segment — the producer copies the rows [r0, r1) of every column as one
        // piece each (16-byte aligned by construction), the pieces sit side by side with leading dimension `m`
        const bool piece = GEN && !tform && (r0 > 0 || min(sl.r1, cb.out_len) < cb.m);
        const int32_t prow = min(sl.r1, cb.out_len) - r0;
        if (piece && prow <= 0) continue;
        // real element types: the piece of a whole chunk of columns arrives as ONE tensor-map box of kFMaxRows rows (rows
        // past the block are zero-filled, rows past the piece are never stored)
        const bool boxed = piece && a.contrib_map != nullptr && __ldg(a.contrib_map + ci) >= 0;
        const int32_t m = boxed ? kFMaxRows
                          : piece ? (prow + (int32_t)(16 / sizeof(T)) - 1) / (int32_t)(16 / sizeof(T)) * (int32_t)(16 / sizeof(T)) : cb.m;
        const int32_t cc = chunk_cols<T>(m);
        const int32_t jlo = tform ? r0 : 0;
        const int32_t jhi = tform ? min(sl.r1, cb.out_len) : cb.n;
        if (jhi <= jlo || m == 0) continue;
        const bool tall = GEN && tform && m > 32 * RPL;
        T *tg = fusedT ? a.scratch + a.contrib_toff[ci] : nullptr;
        if (tform) {
            // x at the block's rows, once per contribution (m <= kXsCap)
            const int32_t mpad = GEN ? max(m, kFMaxRows) : kFMaxRows;
            consumer_bar();
            for (int32_t k = t; k < mpad; k += kFThreads) xs[k] = (k < m) ? a.x.at(in.at(k)) : El<T>::zero();
            consumer_bar();
        }
        for (int32_t jw = jlo; jw < jhi; jw += kXsCap) {
            const int32_t wend = min(jhi, jw + kXsCap);
            if (!tform) {
                consumer_bar();
                for (int32_t k = t; k < wend - jw; k += kFThreads) xs[k] = a.x.at(in.at(jw + k));
                consumer_bar();
            }
            for (int32_t j0 = jw; j0 < wend; j0 += cc, ++q) {
                const int32_t ncols = min(cc, wend - j0);
                const uint32_t stage = q % kPStages;
                const uint32_t delta = piece ? 0u : (uint32_t)(((int64_t)j0 * m * (int64_t)sizeof(T)) & 15);
                mbar_wait(&full[stage], (q / kPStages) & 1);
                const T *sm = reinterpret_cast<const T *>(stages + stage * kPStageBytes + delta);
                if (GEN && tall)
                    consume_chunk_tall<T, CONJ>(sm, m, ncols, j0, lane, warp, xs, accT + (j0 - r0));
                else if (tform)
                    consume_chunk<T, RPL, CONJ>(sm, m, ncols, lane, (warp - pbase) & (kFWarps - 1), false, true, nullptr, xs,
                                                accN, nullptr, accT + (j0 - r0));
                else
                    consume_chunk<T, RPL, CONJ>(sm, m, ncols, lane, (warp - pbase) & (kFWarps - 1), true, fusedT,
                                                xs + (j0 - jw), xrs, accN, fusedT ? tg + j0 : nullptr, nullptr);
                pbase += (uint32_t)((ncols + 1) >> 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
        }
    }
    // every stage has been consumed: reuse the ring for the cross-warp reduction of the row sums
    consumer_bar();
    T *red = reinterpret_cast<T *>(stages);
#pragma unroll
    for (int k = 0; k < RPL; ++k) red[warp * kFMaxRows + k * 32 + lane] = accN[k];
    consumer_bar();
    if (t < L) {
        T tot = accT[t];
#pragma unroll
        for (int w = 0; w < kFWarps; ++w) tot = El<T>::add(tot, red[w * kFMaxRows + t]);
        if (sl.flags & 1) {
            const int32_t row = out.at(r0 + t);
            T v = El<T>::mul(a.alpha, tot);
            if (!a.beta_false) v = El<T>::add(v, El<T>::mul(a.beta, a.y[row]));
            a.y[row] = v;
        } else {
            a.scratch[sl.scratch_off + t] = tot;
        }
    }
}

template <class T>
__device__ __forceinline__ void tma_producer(const MulArgs<T> &a, const bsm_slice &sl, unsigned char *stages,
                                             uint64_t *full, uint64_t *empty) {
    const uint64_t policy = l2_evict_first_policy();
    uint32_t q = 0;
    for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
        const bsm_contrib cb = a.contrib[ci];
        const bool tform = (cb.form & 1) != 0;
        const bool piece = !tform && (sl.r0 > 0 || min(sl.r1, cb.out_len) < cb.m);     // see tma_consumer (GEN plans only)
        const int32_t prow = min(sl.r1, cb.out_len) - sl.r0;
        if (piece && prow <= 0) continue;
        constexpr int32_t kVec = (int32_t)(16 / sizeof(T));
        const int32_t mapi = (piece && a.contrib_map != nullptr) ? __ldg(a.contrib_map + ci) : -1;
        const bool boxed = mapi >= 0;
        const int32_t m = boxed ? kFMaxRows : piece ? (prow + kVec - 1) / kVec * kVec : cb.m;
        const int32_t cc = chunk_cols<T>(m);
        const int32_t jlo = tform ? sl.r0 : 0;
        const int32_t jhi = tform ? min(sl.r1, cb.out_len) : cb.n;
        if (jhi <= jlo || m == 0) continue;
        const unsigned char *blk = reinterpret_cast<const unsigned char *>(a.arena + cb.off);
        for (int32_t jw = jlo; jw < jhi; jw += kXsCap) {
            const int32_t wend = min(jhi, jw + kXsCap);
            for (int32_t j0 = jw; j0 < wend; j0 += cc, ++q) {
                const int32_t ncols = min(cc, wend - j0);
                const uint32_t stage = q % kPStages;
                if (q >= kPStages) mbar_wait(&empty[stage], ((q / kPStages) - 1) & 1);
                if (boxed) {
                    // the tensor map describes the arena as columns of cb.m entries starting at this block's phase: column
                    // coordinate = (block offset) / cb.m + j0; the box is kFMaxRows x chunk_cols and always counts in full
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)(kFMaxRows * cc) * (uint32_t)sizeof(T));
                    tma_box_2d(stages + stage * kPStageBytes, a.piece_maps + (size_t)mapi * 128, sl.r0,
                               (int32_t)(cb.off / cb.m) + j0, &full[stage], policy);
                    continue;
                }
                if (piece) {
                    // one copy per column: rows [r0, r0 + m) of column j0 + j (the packer only cuts such pieces when every
                    // column of the block starts 16-byte aligned and r0 is a multiple of 256)
                    const uint32_t pbytes = (uint32_t)m * (uint32_t)sizeof(T);
                    mbar_arrive_expect_tx(&full[stage], pbytes * (uint32_t)ncols);
                    for (int32_t j = 0; j < ncols; ++j)
                        bulk_g2s(stages + stage * kPStageBytes + (size_t)j * pbytes,
                                 blk + ((int64_t)(j0 + j) * cb.m + sl.r0) * (int64_t)sizeof(T), pbytes, &full[stage], policy);
                    continue;
                }
                const int64_t boff = (int64_t)j0 * m * (int64_t)sizeof(T);
                const uint32_t delta = (uint32_t)(boff & 15);
                const uint32_t bytes = (delta + (uint32_t)ncols * (uint32_t)m * (uint32_t)sizeof(T) + 15u) & ~15u;
                mbar_arrive_expect_tx(&full[stage], bytes);
                bulk_g2s(stages + stage * kPStageBytes, blk + (boff - delta), bytes, &full[stage], policy);
            }
        }
    }
}

template <class T, bool GEN>
__global__ void __launch_bounds__(kPThreads, 2) sym_fused_tma_kernel(const MulArgs<T> a) {
    extern __shared__ __align__(128) unsigned char psm[];
    unsigned char *stages = psm;                                        // kPStages * kPStageBytes
    T *xs = reinterpret_cast<T *>(psm + kPStages * kPStageBytes);       // kXsCap
    T *xrs = xs + kXsCap;                                               // kFMaxRows
    T *accT = xrs + kFMaxRows;                                          // kFMaxRows
    uint64_t *full = reinterpret_cast<uint64_t *>(accT + kFMaxRows);    // kPStages
    uint64_t *empty = full + kPStages;                                  // kPStages
    const bsm_slice sl = a.slices[blockIdx.x];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kFWarps);
        }
        mbar_fence_init();
        if (a.x.npeer) peer_entry(a.x.sync);
    }
    __syncthreads();
    if (threadIdx.x >= kFThreads) {
        if (threadIdx.x == kFThreads) tma_producer<T>(a, sl, stages, full, empty);
        return;
    }
    const int32_t L = sl.r1 - sl.r0;
    if (a.conj) {
        if (L <= 64)
            tma_consumer<T, 2, true, GEN>(a, sl, stages, xs, xrs, accT, full, empty);
        else if (L <= 128)
            tma_consumer<T, 4, true, GEN>(a, sl, stages, xs, xrs, accT, full, empty);
        else
            tma_consumer<T, 8, true, GEN>(a, sl, stages, xs, xrs, accT, full, empty);
    } else {
        if (L <= 64)
            tma_consumer<T, 2, false, GEN>(a, sl, stages, xs, xrs, accT, full, empty);
        else if (L <= 128)
            tma_consumer<T, 4, false, GEN>(a, sl, stages, xs, xrs, accT, full, empty);
        else
            tma_consumer<T, 8, false, GEN>(a, sl, stages, xs, xrs, accT, full, empty);
    }
    // every x read of the consumers precedes the consumer barriers of the final reduction
    if (a.x.npeer && threadIdx.x == 0) peer_exit(a.x.sync);
}

template <class T>
constexpr size_t fused_tma_smem_bytes() {
    return (size_t)kPStages * kPStageBytes + sizeof(T) * (size_t)(kXsCap + 2 * kFMaxRows) + 2 * kPStages * 8;
}
static_assert(kPStages * kPStageBytes >= kFWarps * kFMaxRows * 16, "reduction buffer must fit in the ring");

// ---- warp-stream kernel ----------------------------------------------------------------------------
// Small segments (<= 64 rows, blocks of <= 64 rows: VBCRS block rows, 32x32 BlockSparseMatrix blocks):
// a CTA-wide pipeline would drain at every segment, so here every WARP owns a private shared-memory
// byte ring and streams a long run of segments (a "work item") through it without ever synchronising
// with another warp. Nothing on the critical path is a global load:
//   * block data     cp.async.bulk (TMA, SASS UBLKCP) issued by lane 0 into the ring
//   * x values       per-lane cp.async (LDGSTS) gathers into the ring right behind the block data,
//                    completing on the SAME mbarrier (cp.async.mbarrier.arrive.noinc)
//   * descriptors    bsm_wchunk records, bulk-copied in batches of 8 into a 32-entry descriptor ring
// The ring placement (smem16) and the issue condition (lag) of every chunk are precomputed by the
// packer — the chunk sequence of a work item is static — so ~7 KB per warp x 16 warps per SM are in flight
// regardless of the block sizes (measured on C3: 8 warps x 24 KB rings 0.85 of the copy peak, 12 x 16 KB 0.91,
// 16 x 11 KB 0.96, 20 x 8.5 KB 0.79: per-warp instruction latency, not bytes in flight, is the limit).
// Lanes own rows (lane, lane+32):
//   N-form  acc[r] += B[r, j] * x[j]           x[j] broadcast from shared memory
//   T-form  t[j]   += sum_r B[r, j] * x[r]     four columns at a time, reduced by a halving butterfly
//                                              (6 shuffles per 4 columns instead of 20)
// At the last chunk of a segment the warp writes y (alpha/beta fused) or its partial vector.
constexpr int kWWarps = 4;
constexpr int kWRing = 11264;        // == plan.h kWRingBytes
constexpr int kWNB = 16;             // == plan.h kWSlots
constexpr int kWDBatch = 8;          // descriptors per bulk copy
constexpr int kWDSlots = 4;          // descriptor ring = kWDSlots * kWDBatch records
constexpr int kWSegMax = 64;

template <class T>
struct WarpArgs {
    const unsigned char *arena;
    const bsm_wchunk *chunks;
    const int32_t *item_ptr;
    const int32_t *pool;
    XSrc<T> x;
    T *y;
    T *scratch;
    T alpha, beta;
    int32_t nitems;
    int32_t beta_false;
    int32_t conj;
    // CTA-part mode (small problems, plan.h HostPlan::wcta): the 4 items of a CTA are parts of ONE segment; their
    // partial vectors meet in shared memory and warp 0 writes the outputs. CTAs past the item CTAs set the rows no
    // block touches (zrows) to beta*y, so that the whole multiply is a single launch.
    int32_t cta_mode;
    int32_t nz;
    const int32_t *zrows;
    // x_bulk_len > 0: x (peer mode: every owner's array) is 16-byte aligned — the x values of a chunk (a contiguous range)
    // are fetched by ONE bulk copy of the enclosing 16-byte aligned range, as long as that range ends at or before entry
    // x_bulk_len (the length of x rounded down to whole 16-byte groups); 0: per-element copies (unaligned x)
    int32_t x_bulk_len;
};

struct WDesc {
    int4 lo, hi;
    __device__ __forceinline__ uint64_t src_bytes() const {
        return ((((uint64_t)(((uint32_t)hi.x >> 8) & 0xffu)) << 32) | (uint64_t)(uint32_t)lo.x) << 4;
    }
    __device__ __forceinline__ uint32_t bytes() const { return ((uint32_t)lo.y & 0xffffu) << 4; }
    __device__ __forceinline__ int32_t ncols() const { return (int32_t)((uint32_t)lo.y >> 16); }
    __device__ __forceinline__ int32_t m() const { return (int32_t)((uint32_t)lo.z & 0xffu); }
    __device__ __forceinline__ uint32_t flags() const { return ((uint32_t)lo.z >> 8) & 0xffu; }
    __device__ __forceinline__ uint32_t delta() const { return ((uint32_t)lo.z >> 16) & 0xffu; }
    __device__ __forceinline__ int32_t seg_len() const { return (int32_t)((uint32_t)lo.z >> 24); }
    __device__ __forceinline__ int32_t x_ref() const { return lo.w; }
    __device__ __forceinline__ int32_t out_col() const { return (int32_t)((uint32_t)hi.x & 0xffu); }
    __device__ __forceinline__ uint32_t smem_off() const { return ((uint32_t)hi.x >> 16) << 4; }
    __device__ __forceinline__ int32_t lag() const { return (int32_t)((uint32_t)hi.y & 0xffu); }
    __device__ __forceinline__ int64_t out() const {
        return (int64_t)(((uint64_t)(uint32_t)hi.w << 32) | (uint64_t)(uint32_t)hi.z);
    }
};

template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void *dst_smem, const void *src) {
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES)
                     : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// FP64 tensor-core instruction: D(8x8) += A(8x4, row) * B(4x8, col). Fragments: a = A[lane/4][lane%4],
// b = B[lane%4][lane/4], c0/c1 = C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// T-form chunk in Float64 on the tensor cores: t[j] = sum_r B[r, j] x[r] as (B^T tile, 8 columns x 4 rows) times
// a B-operand whose 8 columns all hold x[r] — 7/8 of the DMMA is wasted, but one DMMA replaces 32 FMAs AND the
// whole shuffle butterfly of the scalar path (which makes T-form twice as instruction-heavy as N-form), and the
// DMMA pipe is idle in an SpMV anyway. NJT = 8-column tiles of the chunk.
template <int NJT>
__device__ __forceinline__ void wchunk_tform_dmma(const double *__restrict__ sm, const double *__restrict__ xin,
                                                  int32_t m, int32_t nc, int32_t oc, int lane, double *ts) {
    const int g = lane >> 2, tg = lane & 3;
    double c[NJT][2];
    const double *ap[NJT];
    bool jok[NJT];
#pragma unroll
    for (int t = 0; t < NJT; ++t) {
        c[t][0] = c[t][1] = 0.0;
        const int32_t j = 8 * t + g;
        jok[t] = j < nc;
        // a column past the chunk reads column 0 instead: row g of the product is garbage and never stored, the other
        // rows do not depend on it — no predicate and no zero fill in the loop
        ap[t] = sm + (jok[t] ? j : 0) * m + tg;
    }
    const double *xp = xin + tg;
    int32_t kt = m >> 2;
    // 16 rows per trip: the fragment loads use immediate offsets, one pointer bump per column tile
    for (; kt >= 4; kt -= 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double b = xp[4 * u];
#pragma unroll
            for (int t = 0; t < NJT; ++t) dmma_m8n8k4(c[t][0], c[t][1], ap[t][4 * u], b);
        }
        xp += 16;
#pragma unroll
        for (int t = 0; t < NJT; ++t) ap[t] += 16;
    }
    for (; kt > 0; --kt) {
        const double b = *xp;
#pragma unroll
        for (int t = 0; t < NJT; ++t) dmma_m8n8k4(c[t][0], c[t][1], *ap[t], b);
        xp += 4;
#pragma unroll
        for (int t = 0; t < NJT; ++t) ap[t] += 4;
    }
    if (m & 3) {
        // rows past the block: both operands zero (what lies behind the chunk in the ring need not be finite)
        const bool rok = (m & ~3) + tg < m;
        const double b = rok ? *xp : 0.0;
#pragma unroll
        for (int t = 0; t < NJT; ++t) {
            const double a = rok ? *ap[t] : 0.0;
            dmma_m8n8k4(c[t][0], c[t][1], a, b);
        }
    }
    if (tg == 0) {
#pragma unroll
        for (int t = 0; t < NJT; ++t)
            if (jok[t]) ts[oc + 8 * t + g] += c[t][0];
    }
}

// T-form chunk in Float64 with lanes owning COLUMNS: lane j walks down column j (x[r] is a broadcast), no cross-lane
// reduction and no tensor-core tile to fill. The 16 lanes of a half-warp read addresses m doubles apart: all 16 banks
// when m is odd, 2-way conflicts when m = 2 mod 4 — against 4- to 8-way conflicts of the DMMA fragment loads for the
// same heights. Heights that are multiples of 4 stay on the DMMA path (m = 4 mod 8 is conflict-free there).
template <bool TWO>
__device__ __forceinline__ void wchunk_tform_cols32(const double *__restrict__ sm, const double *__restrict__ xin,
                                                    int32_t m, int32_t nc, int32_t oc, int lane, double *ts) {
    const bool c0 = lane < nc, c1 = TWO && (lane + 32) < nc;
    const double *p0 = sm + (c0 ? lane : 0) * m;
    const double *p1 = sm + (c1 ? lane + 32 : 0) * m;
    double s0 = 0.0, u0 = 0.0, s1 = 0.0, u1 = 0.0;
    int32_t r = 0;
    for (; r + 4 <= m; r += 4) {
        const double x0 = xin[r], x1 = xin[r + 1], x2 = xin[r + 2], x3 = xin[r + 3];
        s0 = fma(p0[r], x0, s0);
        u0 = fma(p0[r + 1], x1, u0);
        s0 = fma(p0[r + 2], x2, s0);
        u0 = fma(p0[r + 3], x3, u0);
        if (TWO) {
            s1 = fma(p1[r], x0, s1);
            u1 = fma(p1[r + 1], x1, u1);
            s1 = fma(p1[r + 2], x2, s1);
            u1 = fma(p1[r + 3], x3, u1);
        }
    }
    for (; r < m; ++r) {
        const double x0 = xin[r];
        s0 = fma(p0[r], x0, s0);
        if (TWO) s1 = fma(p1[r], x0, s1);
    }
    if (c0) ts[oc + lane] += s0 + u0;
    if (c1) ts[oc + lane + 32] += s1 + u1;
}
// at most 16 columns: the two half-warps take the even and the odd rows
__device__ __forceinline__ void wchunk_tform_cols16(const double *__restrict__ sm, const double *__restrict__ xin,
                                                    int32_t m, int32_t nc, int32_t oc, int lane, double *ts) {
    const int j = lane & 15, grp = lane >> 4;
    const bool cok = j < nc;
    const double *p = sm + (cok ? j : 0) * m;
    double s = 0.0, u = 0.0;
    int32_t r = grp;
    for (; r + 2 < m; r += 4) {
        s = fma(p[r], xin[r], s);
        u = fma(p[r + 2], xin[r + 2], u);
    }
    if (r < m) s = fma(p[r], xin[r], s);
    s += u;
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (grp == 0 && cok) ts[oc + j] += s;
}

__device__ __forceinline__ void wchunk_tform_dmma_dispatch(const double *sm, const double *xin, int32_t m, int32_t nc,
                                                           int32_t oc, int lane, double *ts) {
    if ((m & 3) != 0 && nc > 8) {   // column stride spreads over the banks: lanes own columns (see above)
        if (nc > 32)
            wchunk_tform_cols32<true>(sm, xin, m, nc, oc, lane, ts);
        else if (nc > 16)
            wchunk_tform_cols32<false>(sm, xin, m, nc, oc, lane, ts);
        else
            wchunk_tform_cols16(sm, xin, m, nc, oc, lane, ts);
        return;
    }
    // at most 4 column tiles (32 columns) per pass: 8 accumulator registers, which keeps the kernel within the 64
    // registers of 4 CTAs x 256 threads per SM
    for (int32_t c0 = 0; c0 < nc; c0 += 32) {
        const int32_t ncp = min(32, nc - c0);
        const double *smp = sm + c0 * m;
        switch ((ncp + 7) >> 3) {
        case 1: wchunk_tform_dmma<1>(smp, xin, m, ncp, oc + c0, lane, ts); break;
        case 2: wchunk_tform_dmma<2>(smp, xin, m, ncp, oc + c0, lane, ts); break;
        case 3: wchunk_tform_dmma<3>(smp, xin, m, ncp, oc + c0, lane, ts); break;
        default: wchunk_tform_dmma<4>(smp, xin, m, ncp, oc + c0, lane, ts); break;
        }
    }
}

// One chunk (nc whole columns of an m-row block, column-major in shared memory) of the warp-stream
// kernel. TWO: the block has more than 32 rows, lanes also own row lane+32.
template <class T, bool CONJ, bool TWO>
__device__ __forceinline__ void wchunk_compute(const T *__restrict__ sm, const T *__restrict__ xin, int32_t m,
                                               int32_t nc, bool tform, int32_t oc, int lane, T &acc0, T &acc1,
                                               T *ts) {
    const bool r0ok = lane < m, r1ok = TWO && (lane + 32) < m;
    const T *p0 = sm + lane;
    if (!tform) {
        int32_t j = 0;
        for (; j + 4 <= nc; j += 4) {
            T v[4], w[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                v[c] = r0ok ? p0[(j + c) * m] : El<T>::zero();
                if (TWO) w[c] = r1ok ? p0[(j + c) * m + 32] : El<T>::zero();
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T xj = xin[j + c];
                El<T>::fma(acc0, CONJ ? El<T>::conj(v[c]) : v[c], xj);
                if (TWO) El<T>::fma(acc1, CONJ ? El<T>::conj(w[c]) : w[c], xj);
            }
        }
        for (; j < nc; ++j) {
            const T v = r0ok ? p0[j * m] : El<T>::zero();
            const T xj = xin[j];
            El<T>::fma(acc0, CONJ ? El<T>::conj(v) : v, xj);
            if (TWO) {
                const T w = r1ok ? p0[j * m + 32] : El<T>::zero();
                El<T>::fma(acc1, CONJ ? El<T>::conj(w) : w, xj);
            }
        }
    } else {
        const T xa = r0ok ? xin[lane] : El<T>::zero();
        const T xb = r1ok ? xin[lane + 32] : El<T>::zero();
        for (int32_t j = 0; j < nc; j += 4) {
            T p[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool cok = (j + c) < nc;
                const T v = (cok && r0ok) ? p0[(j + c) * m] : El<T>::zero();
                p[c] = El<T>::mul(CONJ ? El<T>::conj(v) : v, xa);
                if (TWO) {
                    const T w = (cok && r1ok) ? p0[(j + c) * m + 32] : El<T>::zero();
                    El<T>::fma(p[c], CONJ ? El<T>::conj(w) : w, xb);
                }
            }
            // halving butterfly: 4 columns -> lane bits (4,3) select the column
            const bool h16 = (lane & 16) != 0;
            T k0 = El<T>::add(h16 ? p[2] : p[0], El<T>::shfl_xor(h16 ? p[0] : p[2], 16));
            T k1 = El<T>::add(h16 ? p[3] : p[1], El<T>::shfl_xor(h16 ? p[1] : p[3], 16));
            const bool h8 = (lane & 8) != 0;
            T k = El<T>::add(h8 ? k1 : k0, El<T>::shfl_xor(h8 ? k0 : k1, 8));
            k = El<T>::add(k, El<T>::shfl_xor(k, 4));
            k = El<T>::add(k, El<T>::shfl_xor(k, 2));
            k = El<T>::add(k, El<T>::shfl_xor(k, 1));
            const int32_t col = j + (lane >> 3);
            if ((lane & 7) == 0 && col < nc) ts[oc + col] = El<T>::add(ts[oc + col], k);
        }
    }
}

struct WarpEnd {      // CTA-part mode: what a warp's part ended with (shared memory, one record per warp)
    int64_t out;
    uint32_t fl;
    int32_t L;
};

template <class T>
__host__ __device__ constexpr size_t stream_warp_smem_per_warp() {
    return (size_t)kWRing + (size_t)kWDSlots * kWDBatch * 32 + 2 * kWSegMax * sizeof(T) + 8 * (2 * kWNB + kWDSlots);
}
// Every work item is served by TWO warps (round 2): a producer warp walks the chunk descriptors and issues the bulk
// copies as soon as the static ring schedule allows (chunk i needs i - lag chunks consumed: `freed` mbarriers, one
// arrival per consumed chunk), a consumer warp waits on the `full` barriers and computes. The per-chunk bookkeeping of
// the two roles (~130 + ~250 instructions, almost all dependent scalar code: ncu shows warps waiting on fixed-latency
// dependencies and branches, 3 % on memory) now overlaps instead of adding up.
template <class T>
__device__ __forceinline__ void wdesc_load_batch(const WarpArgs<T> &a, const int4 *dring, uint64_t *dbar, int32_t q0,
                                                 int32_t n, int32_t b) {   // one lane
    const uint32_t cnt = (uint32_t)min(kWDBatch, n - b * kWDBatch);
    uint64_t *bar = &dbar[b & (kWDSlots - 1)];
    mbar_arrive_expect_tx(bar, cnt * 32u);
    bulk_g2s_plain(const_cast<int4 *>(dring) + (b & (kWDSlots - 1)) * kWDBatch * 2, a.chunks + q0 + b * kWDBatch,
                   cnt * 32u, bar);
}
__device__ __forceinline__ WDesc wdesc_at(const int4 *dring, int32_t i) {
    WDesc d;
    const int4 *p = dring + (i & (kWDSlots * kWDBatch - 1)) * 2;
    d.lo = p[0];
    d.hi = p[1];
    return d;
}

template <class T>
__device__ __forceinline__ void stream_warp_producer(const WarpArgs<T> &a, unsigned char *ring, const int4 *dring,
                                                     uint64_t *full, uint64_t *dbar, uint64_t *freed, int32_t q0,
                                                     int32_t n) {
    const int lane = threadIdx.x & 31;
    constexpr int32_t kXVec = (int32_t)(16 / sizeof(T));
    const uint64_t policy = l2_evict_first_policy();
    const int32_t nbatch = (n + kWDBatch - 1) / kWDBatch;
    if (lane == 0)
        for (int32_t b = 0; b < kWDSlots && b < nbatch; ++b) wdesc_load_batch(a, dring, dbar, q0, n, b);
    int32_t ibatch = -1, known = 0;   // known: chunks the consumer is known to have finished
    for (int32_t ii = 0; ii < n; ++ii) {
        const int32_t b = ii / kWDBatch;
        if (b != ibatch) {
            mbar_wait(&dbar[b & (kWDSlots - 1)], (uint32_t)((b / kWDSlots) & 1));
            ibatch = b;
        }
        const WDesc d = wdesc_at(dring, ii);
        const int32_t need = ii - d.lag();
        if (need > known) {   // the ring space of this chunk is free once chunk need - 1 has been consumed
            mbar_wait(&freed[(need - 1) & (kWNB - 1)], (uint32_t)(((need - 1) / kWNB) & 1));
            known = need;
        }
        uint64_t *bar = &full[ii & (kWNB - 1)];
        unsigned char *dst = ring + d.smem_off();
        const uint32_t bytes = d.bytes();
        const uint32_t fl = d.flags();
        const int32_t cnt = (fl & 1u) ? d.m() : d.ncols();
        const int32_t xr = d.x_ref();
        // x values of the chunk (a contiguous range, unless gathered through the pool by the consumer): ONE bulk copy
        // of the enclosing 16-byte aligned range when it lies in one array (peer mode: inside one owner's slab),
        // else one cp.async per entry; either way entry x_ref lands (x_ref mod 16 bytes) behind the chunk
        const int32_t xa0 = xr & ~(kXVec - 1), xa1 = (xr + cnt + kXVec - 1) & ~(kXVec - 1);
        bool xbulk = !(fl & 2u) && xa1 <= a.x_bulk_len;
        const T *xsrc = a.x.x + xa0;
        if (xbulk && a.x.npeer) {
            xsrc = a.x.ptr(xa0);
            xbulk = (xsrc + (xa1 - 1 - xa0)) == a.x.ptr(xa1 - 1);
        }
        if (lane == 0) {
            const uint32_t xbytes = xbulk ? (uint32_t)(xa1 - xa0) * (uint32_t)sizeof(T) : 0u;
            mbar_arrive_expect_tx(bar, bytes + xbytes);
            bulk_g2s(dst, a.arena + d.src_bytes(), bytes, bar, policy);
            if (xbulk) bulk_g2s_plain(dst + bytes, xsrc, xbytes, bar);   // x is re-read: no evict-first hint
        }
        if (!(fl & 2u) && !xbulk) {
            T *xd = reinterpret_cast<T *>(dst + bytes) + (xr - xa0);
            if (lane < cnt) cp_async_elem<(int)sizeof(T)>(xd + lane, a.x.ptr(xr + lane));
            if (lane + 32 < cnt) cp_async_elem<(int)sizeof(T)>(xd + lane + 32, a.x.ptr(xr + lane + 32));
        }
        cp_async_mbar_arrive_noinc(bar);
    }
}

// FORM: 0 = the launch holds N-form chunks only, 1 = T-form only, 2 = both (the dead path is compiled out: half the
// code, which matters for small problems whose first wave also pays the instruction fetch)
template <class T, bool CONJ, int FORM>
__device__ __forceinline__ void stream_warp_consumer(const WarpArgs<T> &a, unsigned char *ring, const int4 *dring,
                                                     T *xs, T *ts, uint64_t *full, uint64_t *dbar, uint64_t *freed,
                                                     int32_t q0, int32_t n) {
    const int lane = threadIdx.x & 31;
    constexpr int32_t kXVec = (int32_t)(16 / sizeof(T));
    const int32_t nbatch = (n + kWDBatch - 1) / kWDBatch;
    T acc0 = El<T>::zero(), acc1 = El<T>::zero();
    int32_t cbatch = -1;
    for (int32_t ci = 0; ci < n; ++ci) {
        const int32_t b = ci / kWDBatch;
        if (b != cbatch) {
            mbar_wait(&dbar[b & (kWDSlots - 1)], (uint32_t)((b / kWDSlots) & 1));
            cbatch = b;
            // batch b-1 is behind both cursors (the producer issued its chunks before they were consumed): its slot
            // takes batch b+3
            if (b >= 1 && b + kWDSlots - 1 < nbatch && lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                wdesc_load_batch(a, dring, dbar, q0, n, b + kWDSlots - 1);
            }
        }
        const WDesc d0 = wdesc_at(dring, ci);
        const uint32_t fl = d0.flags();
        const int32_t m = d0.m(), nc = d0.ncols();
        const bool tform = FORM == 2 ? (fl & 1u) != 0 : FORM == 1;
        const unsigned char *cbase = ring + d0.smem_off();
        const T *xin = reinterpret_cast<const T *>(cbase + d0.bytes());
        if (!(fl & 2u)) xin += d0.x_ref() & (kXVec - 1);   // see stream_warp_producer
        if (fl & 8u) {  // first chunk of a segment
            acc0 = El<T>::zero();
            acc1 = El<T>::zero();
            ts[lane] = El<T>::zero();
            ts[lane + 32] = El<T>::zero();
        }
        if (fl & 2u) {
            // arbitrary index vector: gather x through the pool into the warp's staging array
            const int32_t cnt = tform ? m : nc;
            xs[lane] = (lane < cnt) ? a.x.at(__ldg(a.pool + d0.x_ref() + lane)) : El<T>::zero();
            xs[lane + 32] = (lane + 32 < cnt) ? a.x.at(__ldg(a.pool + d0.x_ref() + lane + 32)) : El<T>::zero();
            xin = xs;
        }
        __syncwarp();
        mbar_wait(&full[ci & (kWNB - 1)], (uint32_t)((ci / kWNB) & 1));
        const T *sm = reinterpret_cast<const T *>(cbase + d0.delta());
        bool done_tc = false;
        if constexpr (sizeof(T) == 8) {
            if (tform) {   // Float64 T-form: tensor-core path (no shuffle butterfly)
                wchunk_tform_dmma_dispatch(reinterpret_cast<const double *>(sm), reinterpret_cast<const double *>(xin),
                                           m, nc, d0.out_col(), lane, reinterpret_cast<double *>(ts));
                done_tc = true;
            }
        }
        if (done_tc) {
        } else if (m > 32)
            wchunk_compute<T, CONJ, true>(sm, xin, m, nc, tform, d0.out_col(), lane, acc0, acc1, ts);
        else
            wchunk_compute<T, CONJ, false>(sm, xin, m, nc, tform, d0.out_col(), lane, acc0, acc1, ts);
        __syncwarp();
        if (lane == 0) {   // the chunk's ring space may be overwritten (by the async proxy) from here on
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&freed[ci & (kWNB - 1)]);
        }
        if (fl & 16u) {  // last chunk of the segment: write the outputs
            const int32_t L = d0.seg_len();
            const int64_t o = d0.out();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int32_t r = lane + 32 * h;
                if (r < L) {
                    const T tot = El<T>::add(h ? acc1 : acc0, ts[r]);
                    if (fl & 64u) {          // part of a segment: the CTA sums the parts (stream_warp_kernel)
                        // addresses recomputed here (rare path) so that nothing stays live across the chunk loop
                        const int warp = threadIdx.x >> 5;
                        unsigned char *tail = ring + (kWWarps - warp) * stream_warp_smem_per_warp<T>();
                        reinterpret_cast<T *>(tail)[warp * kWSegMax + r] = tot;
                        WarpEnd *we = reinterpret_cast<WarpEnd *>(tail + kWWarps * kWSegMax * sizeof(T)) + warp;
                        we->out = o;
                        we->fl = fl;
                        we->L = L;
                    } else if (fl & 32u) {
                        const int32_t row = (fl & 4u) ? __ldg(a.pool + o + r) : (int32_t)o + r;
                        T v = El<T>::mul(a.alpha, tot);
                        if (!a.beta_false) v = El<T>::add(v, El<T>::mul(a.beta, a.y[row]));
                        a.y[row] = v;
                    } else {
                        a.scratch[o + r] = tot;
                    }
                }
            }
            __syncwarp();
        }
    }
}

template <class T>
constexpr size_t stream_warp_smem_bytes() {
    return kWWarps * stream_warp_smem_per_warp<T>() + kWWarps * kWSegMax * sizeof(T) + kWWarps * sizeof(WarpEnd);   // + the parts of CTA-part mode
}

constexpr int kWThreads = 2 * kWWarps * 32;   // consumer warps 0..3, producer warps 4..7 (warp w + 4 feeds warp w)

template <class T, int FORM>
__global__ void __launch_bounds__(kWThreads, 4) stream_warp_kernel(const WarpArgs<T> a) {
    extern __shared__ __align__(128) unsigned char wsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int32_t nctas = (a.nitems + kWWarps - 1) / kWWarps;
    if ((int32_t)blockIdx.x >= nctas) {   // rows no block touches: y <- beta*y (these CTAs exist in CTA-part mode only)
        const int32_t i = ((int32_t)blockIdx.x - nctas) * kWThreads + (int32_t)threadIdx.x;
        if (i < a.nz) {
            const int32_t row = __ldg(a.zrows + i) & 0x7fffffff;
            a.y[row] = a.beta_false ? El<T>::zero() : El<T>::mul(a.beta, a.y[row]);
        }
        return;
    }
    const bool producer = warp >= kWWarps;
    const int slot = producer ? warp - kWWarps : warp;
    const int32_t item = blockIdx.x * kWWarps + slot;
    T *parts = reinterpret_cast<T *>(wsm + kWWarps * stream_warp_smem_per_warp<T>());
    WarpEnd *ends = reinterpret_cast<WarpEnd *>(parts + kWWarps * kWSegMax);
    unsigned char *base = wsm + slot * stream_warp_smem_per_warp<T>();
    unsigned char *ring = base;
    const int4 *dring = reinterpret_cast<const int4 *>(base + kWRing);
    T *xs = reinterpret_cast<T *>(base + kWRing + kWDSlots * kWDBatch * 32);
    T *ts = xs + kWSegMax;
    uint64_t *full = reinterpret_cast<uint64_t *>(ts + kWSegMax);
    uint64_t *dbar = full + kWNB;
    uint64_t *freed = dbar + kWDSlots;
    if (!producer && lane == 0) {
        if (a.cta_mode) ends[warp].L = -1;      // no part yet
        for (int i = 0; i < kWNB; ++i) {
            mbar_init(&full[i], 33);            // the producer's lane 0 (expect_tx) + its 32 cp.async arrivals
            mbar_init(&freed[i], 1);            // the consumer's lane 0
        }
        for (int i = 0; i < kWDSlots; ++i) mbar_init(&dbar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    int32_t q0 = 0, q1 = 0;
    if (item < a.nitems) {
        q0 = __ldg(a.item_ptr + item);
        q1 = __ldg(a.item_ptr + item + 1);
    }
    if (producer) {
        if (q0 < q1) {
            if (a.x.npeer) {   // no x entry is fetched before every rank has published its slab
                if (lane == 0) peer_entry(a.x.sync);
                __syncwarp();
            }
            stream_warp_producer<T>(a, ring, dring, full, dbar, freed, q0, q1 - q0);
        }
    } else {
        if (a.x.npeer) {       // pool-gathered x values are read by the consumer itself; every warp arrives at the exit
            if (lane == 0) peer_entry(a.x.sync);
            __syncwarp();
        }
        if (q0 < q1) {
            bool conj_done = false;
            if constexpr (sizeof(T) == 16) {       // conj is the identity for real element types: one instantiation
                if (a.conj) {
                    stream_warp_consumer<T, true, FORM>(a, ring, dring, xs, ts, full, dbar, freed, q0, q1 - q0);
                    conj_done = true;
                }
            }
            if (!conj_done) stream_warp_consumer<T, false, FORM>(a, ring, dring, xs, ts, full, dbar, freed, q0, q1 - q0);
        }
    }
    if (a.cta_mode) {
        // the parts of the segment meet here: fixed order (warp 0, 1, 2, 3), one writer
        __syncthreads();
        const WarpEnd e0 = ends[0];
        if (warp == 0 && e0.L >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int32_t r = lane + 32 * h;
                if (r < e0.L) {
                    T tot = parts[r];
#pragma unroll
                    for (int w = 1; w < kWWarps; ++w)
                        if (ends[w].L >= 0) tot = El<T>::add(tot, parts[w * kWSegMax + r]);
                    if (e0.fl & 32u) {
                        const int32_t row = (e0.fl & 4u) ? __ldg(a.pool + e0.out + r) : (int32_t)e0.out + r;
                        T v = El<T>::mul(a.alpha, tot);
                        if (!a.beta_false) v = El<T>::add(v, El<T>::mul(a.beta, a.y[row]));
                        a.y[row] = v;
                    } else {
                        a.scratch[e0.out + r] = tot;
                    }
                }
            }
        }
    }
    if (a.x.npeer && !producer) {   // the consumer has seen every chunk's x values land
        __syncwarp();
        if (lane == 0) peer_exit(a.x.sync);
    }
}

// y <- beta*y (beta === false: exact zeros, NaN/Inf in y are not propagated) — first step of the
// colour-ordered variant, /root/reference/src/blockmatrix.jl:231 `y .*= β`.
template <class T>
__global__ void __launch_bounds__(256) scale_kernel(T *y, int64_t n, T beta, int beta_false) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = beta_false ? El<T>::zero() : El<T>::mul(beta, y[i]);
}

template <class T>
struct FinalizeArgs {
    const int32_t *rows;
    const int64_t *ptr;
    const int64_t *pos;
    const T *scratch;
    T *y;
    T alpha, beta;
    int64_t n;
    int64_t ldy;
    int32_t beta_false;
};

template <class T>
__global__ void __launch_bounds__(256) gather_finalize_kernel(const FinalizeArgs<T> a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int32_t rr = a.rows[i];
    const int32_t row = rr & 0x7fffffff;
    T *y = a.y + (int64_t)blockIdx.y * a.ldy;   // blockIdx.y: right-hand side (multi-RHS path, no partial sums)
    T s = El<T>::zero();
    for (int64_t k = a.ptr[i]; k < a.ptr[i + 1]; ++k) s = El<T>::add(s, a.scratch[a.pos[k]]);
    T v = El<T>::mul(a.alpha, s);
    if (rr < 0) {
        v = El<T>::add(v, y[row]);  // a direct slice already wrote alpha*acc + beta*y here
    } else if (!a.beta_false) {
        v = El<T>::add(v, El<T>::mul(a.beta, y[row]));
    }
    y[row] = v;
}

}  // namespace bsm
