// krylov.cu — a Krylov iteration driven through the operator with every scalar on the device (SURVEY.md §8f row 2:
// "solver-loop integration: sharded-x multiply and fused y = alpha*A*x + beta*y + dot / axpy").
//
// The reference is used through LinearMaps by iterative solvers (/root/reference/docs/src/block.md:56-63 times
// B*y, B'*y, transpose(B)*y — the building block of every Krylov method); it ships no solver itself. bsm_cg is that
// loop kept on the GPU: conjugate gradients with either the Hermitian inner product (CG, Hermitian positive definite
// operators) or the unconjugated bilinear form (COCG — what a complex SYMMETRIC operator such as a
// SymmetricBlockMatrix{ComplexF64} of a BEM near field calls for; identical to CG for real dtypes).
//
// Per iteration: q = A p (bsm_mul / bsm_mul_dist_peer), then three fused vector kernels
//   dot_partial   p.q                                   -> partial sums
//   cg_update     alpha = rr/pq;  x += alpha p;  r -= alpha q;  partial sums of r.r and |r|^2
//   cg_direction  beta = rr'/rr;  p = r + beta p        (p lives in the peer-mapped array when the operator is sharded)
// Reductions are two-stage with a FIXED number of partial sums and fixed trees, so results are bitwise reproducible;
// the scalars alpha, beta never visit the host. Sharded: every rank owns its block-row slab of x, r, p, q; the two
// dot products per iteration are summed over the ranks with ncclAllReduce on 2 + 3 doubles; p is read by the peers
// straight from its owners (the multiply's exit barrier orders the update of p behind the peers' reads).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bsm_b200.h"
#include "kernels.cuh"

void bsm_set_error(const std::string &msg);
int bsm_dist_allreduce_sum_f64_internal(bsm_comm c, double *dev_values, int64_t count, void *stream);
int bsm_dist_swap_debug_internal(bsm_comm c, int flags);

namespace {

using namespace bsm;

constexpr int kParts = 512;      // partial sums per reduction: fixed, so the summation order never depends on n
constexpr int kRThreads = 256;

int kfail(int code, const std::string &msg) {
    bsm_set_error(msg);
    return code;
}
#define K_TRY(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) return kfail(BSM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// device scalars (doubles): complex values take two slots
enum { S_RR0 = 0, S_RR1 = 2, S_PQ = 4, S_RN = 6, S_BN = 7, S_COUNT = 8 };

struct C2 {
    double re, im;
};
template <class T>
__device__ __forceinline__ C2 widen(T v);
template <>
__device__ __forceinline__ C2 widen<float>(float v) { return C2{(double)v, 0.0}; }
template <>
__device__ __forceinline__ C2 widen<double>(double v) { return C2{v, 0.0}; }
template <>
__device__ __forceinline__ C2 widen<cplx>(cplx v) { return C2{v.re, v.im}; }
template <class T>
__device__ __forceinline__ T narrow(C2 v);
template <>
__device__ __forceinline__ float narrow<float>(C2 v) { return (float)v.re; }
template <>
__device__ __forceinline__ double narrow<double>(C2 v) { return v.re; }
template <>
__device__ __forceinline__ cplx narrow<cplx>(C2 v) { return cplx{v.re, v.im}; }
__device__ __forceinline__ C2 cmul(C2 a, C2 b) { return C2{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ C2 cdiv(C2 a, C2 b) {
    const double d = b.re * b.re + b.im * b.im;
    return C2{(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// fixed-tree block reduction of up to 3 doubles per thread; result valid in thread 0
__device__ __forceinline__ void block_reduce3(double &a, double &b, double &c, double (*sm)[kRThreads]) {
    const int t = threadIdx.x;
    sm[0][t] = a;
    sm[1][t] = b;
    sm[2][t] = c;
    __syncthreads();
    for (int s = kRThreads / 2; s > 0; s >>= 1) {
        if (t < s) {
            sm[0][t] += sm[0][t + s];
            sm[1][t] += sm[1][t + s];
            sm[2][t] += sm[2][t + s];
        }
        __syncthreads();
    }
    a = sm[0][0];
    b = sm[1][0];
    c = sm[2][0];
}

// part[0..kParts) (re), part[kParts..2kParts) (im): sum over i in [lo, hi) of op(x_i) * y_i
template <class T>
__global__ void __launch_bounds__(kRThreads) dot_partial_kernel(const T *x, const T *y, int64_t lo, int64_t hi, int herm,
                                                                double *part) {
    __shared__ double sm[3][kRThreads];
    const int64_t n = hi - lo, chunk = (n + kParts - 1) / kParts;
    const int64_t b0 = lo + (int64_t)blockIdx.x * chunk, b1 = min(hi, b0 + chunk);
    double re = 0.0, im = 0.0, z = 0.0;
    for (int64_t i = b0 + threadIdx.x; i < b1; i += kRThreads) {
        C2 a = widen<T>(x[i]);
        if (herm) a.im = -a.im;
        const C2 p = cmul(a, widen<T>(y[i]));
        re += p.re;
        im += p.im;
    }
    block_reduce3(re, im, z, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = re;
        part[kParts + blockIdx.x] = im;
    }
}

// one block: out[k] = sum over the kParts partials of column k (ncol <= 3), fixed tree
__global__ void __launch_bounds__(kRThreads) reduce_final_kernel(const double *part, int ncol, double *out, double *hist) {
    __shared__ double sm[3][kRThreads];
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < ncol; ++k)
        for (int i = threadIdx.x; i < kParts; i += kRThreads) v[k] += part[k * kParts + i];
    block_reduce3(v[0], v[1], v[2], sm);
    if (threadIdx.x == 0) {
        for (int k = 0; k < ncol; ++k) out[k] = v[k];
        if (hist) *hist = v[ncol - 1];
    }
}

// alpha = rr / pq;  x += alpha p;  r -= alpha q;  partials of r.r (bilinear or Hermitian) and of |r|^2
template <class T>
__global__ void __launch_bounds__(kRThreads) cg_update_kernel(const T *p, const T *q, T *x, T *r, int64_t lo, int64_t hi, int herm,
                                                              const double *scal, int rr_slot, double *part) {
    __shared__ double sm[3][kRThreads];
    const C2 rr{scal[rr_slot], scal[rr_slot + 1]}, pq{scal[S_PQ], scal[S_PQ + 1]};
    const C2 alpha = cdiv(rr, pq);
    const int64_t n = hi - lo, chunk = (n + kParts - 1) / kParts;
    const int64_t b0 = lo + (int64_t)blockIdx.x * chunk, b1 = min(hi, b0 + chunk);
    double re = 0.0, im = 0.0, nn = 0.0;
    for (int64_t i = b0 + threadIdx.x; i < b1; i += kRThreads) {
        const C2 pi = widen<T>(p[i]), qi = widen<T>(q[i]);
        C2 xi = widen<T>(x[i]), ri = widen<T>(r[i]);
        const C2 ap = cmul(alpha, pi), aq = cmul(alpha, qi);
        xi.re += ap.re;
        xi.im += ap.im;
        ri.re -= aq.re;
        ri.im -= aq.im;
        x[i] = narrow<T>(xi);
        const T rs = narrow<T>(ri);
        r[i] = rs;
        const C2 rw = widen<T>(rs);     // what is stored is what the next iteration sees
        C2 rc = rw;
        if (herm) rc.im = -rc.im;
        const C2 pr = cmul(rc, rw);
        re += pr.re;
        im += pr.im;
        nn += rw.re * rw.re + rw.im * rw.im;
    }
    block_reduce3(re, im, nn, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = re;
        part[kParts + blockIdx.x] = im;
        part[2 * kParts + blockIdx.x] = nn;
    }
}

// beta = rr_new / rr_old;  p = r + beta p
template <class T>
__global__ void __launch_bounds__(kRThreads) cg_direction_kernel(const T *r, T *p, int64_t lo, int64_t hi, const double *scal,
                                                                 int rr_old, int rr_new) {
    const C2 beta = cdiv(C2{scal[rr_new], scal[rr_new + 1]}, C2{scal[rr_old], scal[rr_old + 1]});
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const C2 bp = cmul(beta, widen<T>(p[i]));
    const C2 ri = widen<T>(r[i]);
    p[i] = narrow<T>(C2{ri.re + bp.re, ri.im + bp.im});
}

template <class T>
__global__ void __launch_bounds__(kRThreads) cg_init_kernel(const T *b, T *x, T *r, T *p, int64_t lo, int64_t hi) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const T v = b[i];
    r[i] = v;
    p[i] = v;
    x[i] = El<T>::zero();
}

template <class T>
int cg_impl(bsm_comm c, bsm_handle h, const T *b, T *x, const int64_t *cuts, int rank, int nranks, const bsm_cg_options &o,
            int64_t *iters_out, double *relres_out, cudaStream_t st) {
    int64_t nr = 0, nc = 0;
    if (bsm_size(h, &nr, &nc)) return BSM_ERR_ARG;
    if (nr != nc) return kfail(BSM_ERR_ARG, "bsm_cg needs a square operator");
    const int64_t n = nr;
    const int64_t lo = c ? cuts[rank] : 0, hi = c ? cuts[rank + 1] : n;
    const int64_t maxit = std::max<int64_t>(0, o.maxit);
    const int check = o.check_every > 0 ? o.check_every : 8;
    const int herm = o.hermitian ? 1 : 0;
    T *r = nullptr, *q = nullptr, *p = nullptr;
    double *scal = nullptr, *part = nullptr, *hist = nullptr;
    void *p_shared = nullptr;
    // Sharded solve: the multiply's exit barrier (wait until every peer has finished reading this rank's slab of p) is
    // not needed here — the all-reduce of p.q sits between every multiply and the next update of p, so no rank reaches
    // that update before all ranks have finished the multiply. Bit 1 of the communicator's flags = no exit wait.
    const int old_flags = c ? bsm_dist_swap_debug_internal(c, 0) : 0;
    if (c) bsm_dist_swap_debug_internal(c, old_flags | 2);
    struct Cleanup {
        bsm_comm c;
        T **r, **q, **p;
        double **scal, **part, **hist;
        void **p_shared;
        int old_flags;
        ~Cleanup() {
            if (c) bsm_dist_swap_debug_internal(c, old_flags);
            if (*r) cudaFree(*r);
            if (*q) cudaFree(*q);
            if (*scal) cudaFree(*scal);
            if (*part) cudaFree(*part);
            if (*hist) cudaFree(*hist);
            if (c && *p_shared)
                bsm_dist_free(c, *p_shared);
            else if (*p)
                cudaFree(*p);
        }
    } cleanup{c, &r, &q, &p, &scal, &part, &hist, &p_shared, old_flags};
    K_TRY(cudaMalloc((void **)&r, (size_t)n * sizeof(T)));
    K_TRY(cudaMalloc((void **)&q, (size_t)n * sizeof(T)));
    if (c) {   // the search direction is what the peers read: it lives in a peer-mapped array (collective allocation)
        if (int rc = bsm_dist_alloc(c, (size_t)n * sizeof(T), &p_shared)) return rc;
        p = (T *)p_shared;
    } else {
        K_TRY(cudaMalloc((void **)&p, (size_t)n * sizeof(T)));
    }
    K_TRY(cudaMalloc((void **)&scal, S_COUNT * sizeof(double)));
    K_TRY(cudaMalloc((void **)&part, 3 * kParts * sizeof(double)));
    K_TRY(cudaMalloc((void **)&hist, (size_t)(maxit + 1) * sizeof(double)));
    K_TRY(cudaMemsetAsync(scal, 0, S_COUNT * sizeof(double), st));
    const int64_t rows = hi - lo;
    const unsigned eg = (unsigned)std::max<int64_t>(1, (rows + kRThreads - 1) / kRThreads);
    T one, zero;
    std::memset(&zero, 0, sizeof(T));
    std::memset(&one, 0, sizeof(T));
    if (sizeof(T) == 4) {
        const float f = 1.f;
        std::memcpy(&one, &f, 4);
    } else {
        const double d = 1.0;
        std::memcpy(&one, &d, 8);
    }
    auto allreduce = [&](double *v, int64_t cnt) -> int {
        return c ? bsm_dist_allreduce_sum_f64_internal(c, v, cnt, (void *)st) : 0;
    };
    // x = 0, r = p = b; rr = r.r, |b|^2
    cg_init_kernel<T><<<eg, kRThreads, 0, st>>>(b, x, r, p, lo, hi);
    dot_partial_kernel<T><<<kParts, kRThreads, 0, st>>>(r, r, lo, hi, herm, part);
    reduce_final_kernel<<<1, kRThreads, 0, st>>>(part, 2, scal + S_RR0, nullptr);
    dot_partial_kernel<T><<<kParts, kRThreads, 0, st>>>(r, r, lo, hi, 1, part);
    reduce_final_kernel<<<1, kRThreads, 0, st>>>(part, 1, scal + S_BN, nullptr);
    K_TRY(cudaGetLastError());
    if (int rc = allreduce(scal + S_RR0, 2)) return rc;
    if (int rc = allreduce(scal + S_BN, 1)) return rc;
    double bn2 = 0.0;
    K_TRY(cudaMemcpyAsync(&bn2, scal + S_BN, sizeof(double), cudaMemcpyDeviceToHost, st));
    K_TRY(cudaStreamSynchronize(st));
    int64_t it = 0;
    double relres = bn2 > 0.0 ? 1.0 : 0.0;
    std::vector<double> hbuf((size_t)check);
    while (it < maxit && relres > o.rtol) {
        const int64_t burst = std::min<int64_t>(check, maxit - it);
        for (int64_t k = 0; k < burst; ++k, ++it) {
            const int rr_old = (it & 1) ? S_RR1 : S_RR0, rr_new = (it & 1) ? S_RR0 : S_RR1;
            int rc;
            if (c)
                rc = bsm_mul_dist_peer(c, h, BSM_OP_N, &one, &zero, 1, p, q, cuts, (void *)st);
            else
                rc = bsm_mul(h, BSM_OP_N, &one, &zero, 1, p, n, q, n, 1, (void *)st);
            if (rc) return rc;
            dot_partial_kernel<T><<<kParts, kRThreads, 0, st>>>(p, q, lo, hi, herm, part);
            reduce_final_kernel<<<1, kRThreads, 0, st>>>(part, 2, scal + S_PQ, nullptr);
            if ((rc = allreduce(scal + S_PQ, 2))) return rc;
            cg_update_kernel<T><<<kParts, kRThreads, 0, st>>>(p, q, x, r, lo, hi, herm, scal, rr_old, part);
            // rr_new (2 doubles) and |r|^2 land in consecutive slots only for rr_new = S_RR1 ... keep them separate
            reduce_final_kernel<<<1, kRThreads, 0, st>>>(part, 2, scal + rr_new, nullptr);
            reduce_final_kernel<<<1, kRThreads, 0, st>>>(part + 2 * kParts, 1, scal + S_RN, nullptr);
            if ((rc = allreduce(scal + rr_new, 2))) return rc;
            if ((rc = allreduce(scal + S_RN, 1))) return rc;
            K_TRY(cudaMemcpyAsync(hist + it, scal + S_RN, sizeof(double), cudaMemcpyDeviceToDevice, st));
            cg_direction_kernel<T><<<eg, kRThreads, 0, st>>>(r, p, lo, hi, scal, rr_old, rr_new);
            K_TRY(cudaGetLastError());
        }
        K_TRY(cudaMemcpyAsync(hbuf.data(), hist + (it - burst), (size_t)burst * sizeof(double), cudaMemcpyDeviceToHost, st));
        K_TRY(cudaStreamSynchronize(st));
        // the first iteration of the burst that met the tolerance ends the solve (x carries the later updates too:
        // they only improve it)
        for (int64_t k = 0; k < burst; ++k) {
            relres = std::sqrt(hbuf[(size_t)k] / bn2);
            if (!(relres > o.rtol)) break;
        }
        if (!std::isfinite(relres)) return kfail(BSM_ERR_ARG, "bsm_cg broke down (non-finite residual): the operator is not definite enough for CG / COCG");
        relres = std::sqrt(hbuf[(size_t)(burst - 1)] / bn2);
    }
    if (iters_out) *iters_out = it;
    if (relres_out) *relres_out = relres;
    return 0;
}

}  // namespace

extern "C" {

void bsm_cg_default_options(bsm_cg_options *o) {
    if (!o) return;
    o->rtol = 1e-10;
    o->maxit = 200;
    o->hermitian = 0;
    o->check_every = 8;
}

int bsm_cg(bsm_handle h, const void *b_dev, void *x_dev, const bsm_cg_options *opt, int64_t *iters, double *relres, void *stream) {
    if (!h || !b_dev || !x_dev) return kfail(BSM_ERR_ARG, "null argument");
    bsm_cg_options o;
    bsm_cg_default_options(&o);
    if (opt) o = *opt;
    switch (bsm_dtype_of(h)) {
    case BSM_F32: return cg_impl<float>(nullptr, h, (const float *)b_dev, (float *)x_dev, nullptr, 0, 1, o, iters, relres, (cudaStream_t)stream);
    case BSM_F64: return cg_impl<double>(nullptr, h, (const double *)b_dev, (double *)x_dev, nullptr, 0, 1, o, iters, relres, (cudaStream_t)stream);
    case BSM_C64: return cg_impl<cplx>(nullptr, h, (const cplx *)b_dev, (cplx *)x_dev, nullptr, 0, 1, o, iters, relres, (cudaStream_t)stream);
    }
    return kfail(BSM_ERR_ARG, "bad handle");
}

int bsm_cg_dist(bsm_comm c, bsm_handle h, const void *b_dev, void *x_dev, const int64_t *cuts, const bsm_cg_options *opt,
                int64_t *iters, double *relres, void *stream) {
    if (!c || !h || !b_dev || !x_dev || !cuts) return kfail(BSM_ERR_ARG, "null argument");
    bsm_cg_options o;
    bsm_cg_default_options(&o);
    if (opt) o = *opt;
    int nranks = 1, rank = 0;
    if (int rc = bsm_dist_info(c, &nranks, &rank, nullptr)) return rc;
    switch (bsm_dtype_of(h)) {
    case BSM_F32: return cg_impl<float>(c, h, (const float *)b_dev, (float *)x_dev, cuts, rank, nranks, o, iters, relres, (cudaStream_t)stream);
    case BSM_F64: return cg_impl<double>(c, h, (const double *)b_dev, (double *)x_dev, cuts, rank, nranks, o, iters, relres, (cudaStream_t)stream);
    case BSM_C64: return cg_impl<cplx>(c, h, (const cplx *)b_dev, (cplx *)x_dev, cuts, rank, nranks, o, iters, relres, (cudaStream_t)stream);
    }
    return kfail(BSM_ERR_ARG, "bad handle");
}

}  // extern "C"
