// sparse.cu — SparseArrays.sparse(op(A)) built on the device (§8f row 3: the conversion users call next to the
// multiply path; /root/reference/src/sparse.jl:17-129: rowcolvals pushes one COO triplet per stored entry — for a
// SymmetricBlockMatrix the off-diagonal blocks twice — and `sparse(I, J, V)` canonicalises: column-major order,
// rows sorted inside a column, duplicates summed, explicit zeros kept).
//
// Device pipeline over the resident arena and the op's contribution table (every block use exactly once):
//   expand   one thread per stored entry: key = col << 32 | row, value = op-conjugated entry, in block order
//   sort     cub::DeviceRadixSort::SortPairs on (key, entry index), stable, only the significant key bits
//   heads    first entry of every distinct key, exclusive scan -> position in the CSC arrays
//   reduce   one thread per distinct key sums its duplicates in sorted (= input) order, writes rowval / nzval
//   colptr   lower_bound of every column in the distinct keys
// CUB is library code; this path is off the timed multiply and exists for parity (bit-exact structure) and for
// interop (cuSPARSE, preconditioner setup) without a round trip through the host.
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.h"

using namespace bsm;

void bsm_set_error(const std::string &msg);

// accessors into the handle (defined in abi.cu)
int bsm_sparse_source(bsm_handle h, SparseSource *out);
SparseResult *bsm_sparse_slot(bsm_handle h);

namespace {

int sfail(int code, const std::string &msg) {
    bsm_set_error(msg);
    return code;
}
#define SP_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) return sfail(BSM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct ExpContrib {
    int64_t off;        // arena element offset
    int64_t base;       // first COO slot of this contribution
    int32_t m, n;
    int32_t out_set, in_set;
    int32_t tform;
    int32_t pad;
};

__device__ __forceinline__ int32_t set_at(const int32_t *start, const int64_t *poff, const int32_t *pool, int32_t s,
                                          int32_t k) {
    const int32_t st = start[s];
    return st >= 0 ? st + k : pool[poff[s] + k];
}

template <class T>
__global__ void __launch_bounds__(256) expand_kernel(const T *arena, const ExpContrib *cs, const int32_t *set_start,
                                                     const int64_t *set_pool_off, const int32_t *pool, int conj,
                                                     uint64_t *keys, uint32_t *idx, T *vals) {
    const ExpContrib c = cs[blockIdx.y];
    const int64_t cnt = (int64_t)c.m * c.n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = (int32_t)(e % c.m), j = (int32_t)(e / c.m);
        // N-form: entry (out[i], in[j]); T-form: entry (out[j], in[i])
        const int32_t row = set_at(set_start, set_pool_off, pool, c.out_set, c.tform ? j : i);
        const int32_t col = set_at(set_start, set_pool_off, pool, c.in_set, c.tform ? i : j);
        T v = arena[c.off + e];
        if (conj) v = El<T>::conj(v);
        keys[c.base + e] = ((uint64_t)(uint32_t)col << 32) | (uint32_t)row;
        idx[c.base + e] = (uint32_t)(c.base + e);
        vals[c.base + e] = v;
    }
}

__global__ void __launch_bounds__(256) heads_kernel(const uint64_t *keys, int64_t n, int32_t *flag) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flag[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

template <class T>
__global__ void __launch_bounds__(256) reduce_kernel(const uint64_t *keys, const uint32_t *idx, const int32_t *flag,
                                                     const int64_t *pos, const T *vals, int64_t n, uint64_t *ukeys,
                                                     int64_t *rowval, T *nzval) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !flag[k]) return;
    const uint64_t key = keys[k];
    T s = vals[idx[k]];
    for (int64_t q = k + 1; q < n && keys[q] == key; ++q) s = El<T>::add(s, vals[idx[q]]);   // duplicates, in input order
    const int64_t p = pos[k];
    ukeys[p] = key;
    rowval[p] = (int64_t)(uint32_t)key + 1;   // 1-based, as Julia holds it
    nzval[p] = s;
}

__global__ void __launch_bounds__(256) colptr_kernel(const uint64_t *ukeys, int64_t nu, int64_t ncols, int64_t *colptr) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncols) return;
    const uint64_t target = (uint64_t)c << 32;
    int64_t lo = 0, hi = nu;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ukeys[mid] < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    colptr[c] = lo + 1;   // 1-based
}

template <class T>
int build_csc(const SparseSource &S, int op, SparseResult *R) {
    const HostMatrix &H = *S.H;
    const HostPlan &P = H.plan[op == BSM_OP_N ? 0 : 1];
    // contributions with their output set (group key) and COO base offsets
    std::vector<ExpContrib> cs;
    int64_t total = 0, maxcnt = 0;
    for (size_t g = 0; g + 1 < P.group_ptr.size(); ++g)
        for (int64_t c = P.group_ptr[g]; c < P.group_ptr[g + 1]; ++c) {
            const bsm_contrib &cb = P.contrib[(size_t)c];
            if (cb.m == 0 || cb.n == 0) continue;
            ExpContrib e;
            e.off = cb.off;
            e.base = total;
            e.m = cb.m;
            e.n = cb.n;
            e.out_set = P.group_set[g];
            e.in_set = cb.in_set;
            e.tform = (cb.form & kFormT) ? 1 : 0;
            e.pad = 0;
            cs.push_back(e);
            total += (int64_t)cb.m * cb.n;
            maxcnt = std::max<int64_t>(maxcnt, (int64_t)cb.m * cb.n);
        }
    const int64_t nrows = op == BSM_OP_N ? H.nrows : H.ncols, ncols = op == BSM_OP_N ? H.ncols : H.nrows;
    if (total >= (int64_t)1 << 32) return sfail(BSM_ERR_UNSUPPORTED, "more than 2^32 stored entries");
    (void)nrows;
    R->ncols = ncols;
    R->dtype = H.dtype;
    SP_TRY(cudaMalloc(&R->colptr, (size_t)(ncols + 1) * 8));
    if (total == 0) {
        std::vector<int64_t> ones((size_t)ncols + 1, 1);
        SP_TRY(cudaMemcpy(R->colptr, ones.data(), ones.size() * 8, cudaMemcpyHostToDevice));
        R->nnz = 0;
        return 0;
    }
    ExpContrib *dcs = nullptr;
    uint64_t *keys = nullptr, *keys2 = nullptr, *ukeys = nullptr;
    uint32_t *idx = nullptr, *idx2 = nullptr;
    int32_t *flag = nullptr;
    int64_t *pos = nullptr;
    T *vals = nullptr;
    void *tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(dcs); cudaFree(keys); cudaFree(keys2); cudaFree(ukeys); cudaFree(idx); cudaFree(idx2);
        cudaFree(flag); cudaFree(pos); cudaFree(vals); cudaFree(tmp);
    };
#define SP_TRY_C(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cleanup();                                                                                 \
            return sfail(BSM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));           \
        }                                                                                              \
    } while (0)
    SP_TRY_C(cudaMalloc(&dcs, cs.size() * sizeof(ExpContrib)));
    SP_TRY_C(cudaMemcpy(dcs, cs.data(), cs.size() * sizeof(ExpContrib), cudaMemcpyHostToDevice));
    SP_TRY_C(cudaMalloc(&keys, (size_t)total * 8));
    SP_TRY_C(cudaMalloc(&keys2, (size_t)total * 8));
    SP_TRY_C(cudaMalloc(&idx, (size_t)total * 4));
    SP_TRY_C(cudaMalloc(&idx2, (size_t)total * 4));
    SP_TRY_C(cudaMalloc(&vals, (size_t)total * sizeof(T)));
    {
        const unsigned gx = (unsigned)std::min<int64_t>(64, (maxcnt + 255) / 256);
        for (size_t c0 = 0; c0 < cs.size(); c0 += 65535) {   // grid.y limit
            dim3 grid(gx, (unsigned)std::min<size_t>(65535, cs.size() - c0));
            expand_kernel<T><<<grid, 256>>>((const T *)S.arena, dcs + c0, S.set_start, S.set_pool_off, S.pool,
                                            (op == BSM_OP_C && sizeof(T) == 16) ? 1 : 0, keys, idx, vals);
        }
        SP_TRY_C(cudaGetLastError());
    }
    int colbits = 1;
    while (((int64_t)1 << colbits) < ncols) ++colbits;
    size_t tbytes = 0;
    SP_TRY_C(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, keys, keys2, idx, idx2, total, 0, 32 + colbits));
    size_t sbytes = 0;
    SP_TRY_C(cudaMalloc(&flag, (size_t)total * 4));
    SP_TRY_C(cudaMalloc(&pos, (size_t)total * 8));
    SP_TRY_C(cub::DeviceScan::ExclusiveSum(nullptr, sbytes, flag, pos, total));
    SP_TRY_C(cudaMalloc(&tmp, std::max(tbytes, sbytes)));
    SP_TRY_C(cub::DeviceRadixSort::SortPairs(tmp, tbytes, keys, keys2, idx, idx2, total, 0, 32 + colbits));
    const unsigned nb = (unsigned)((total + 255) / 256);
    heads_kernel<<<nb, 256>>>(keys2, total, flag);
    SP_TRY_C(cub::DeviceScan::ExclusiveSum(tmp, sbytes, flag, pos, total));
    int64_t last_pos = 0;
    int32_t last_flag = 0;
    SP_TRY_C(cudaMemcpy(&last_pos, pos + total - 1, 8, cudaMemcpyDeviceToHost));
    SP_TRY_C(cudaMemcpy(&last_flag, flag + total - 1, 4, cudaMemcpyDeviceToHost));
    const int64_t nu = last_pos + last_flag;
    SP_TRY_C(cudaMalloc(&ukeys, (size_t)nu * 8));
    SP_TRY_C(cudaMalloc(&R->rowval, (size_t)nu * 8));
    SP_TRY_C(cudaMalloc(&R->nzval, (size_t)nu * sizeof(T)));
    reduce_kernel<T><<<nb, 256>>>(keys2, idx2, flag, pos, vals, total, ukeys, (int64_t *)R->rowval, (T *)R->nzval);
    colptr_kernel<<<(unsigned)((ncols + 1 + 255) / 256), 256>>>(ukeys, nu, ncols, (int64_t *)R->colptr);
    SP_TRY_C(cudaGetLastError());
    SP_TRY_C(cudaDeviceSynchronize());
    cleanup();
    R->nnz = nu;
    return 0;
#undef SP_TRY_C
}

void release(SparseResult *R) {
    if (R->colptr) cudaFree(R->colptr);
    if (R->rowval) cudaFree(R->rowval);
    if (R->nzval) cudaFree(R->nzval);
    *R = SparseResult();
}

}  // namespace

extern "C" {

int bsm_sparse_build(bsm_handle h, int op, int64_t *nnz_out) {
    if (!h) return sfail(BSM_ERR_ARG, "null handle");
    if (op < BSM_OP_N || op > BSM_OP_C) return sfail(BSM_ERR_ARG, "bad op");
    SparseSource S;
    if (int rc = bsm_sparse_source(h, &S)) return rc;
    if (S.restricted) return sfail(BSM_ERR_UNSUPPORTED, "sparse() of a slab handle: convert the full matrix");
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(S.device) != cudaSuccess) return sfail(BSM_ERR_CUDA, "cudaSetDevice failed");
    SparseResult *R = bsm_sparse_slot(h);
    release(R);
    int rc;
    switch (S.H->dtype) {
    case BSM_F32: rc = build_csc<float>(S, op, R); break;
    case BSM_F64: rc = build_csc<double>(S, op, R); break;
    default: rc = build_csc<cplx>(S, op, R); break;
    }
    if (rc != 0) release(R);
    if (prev >= 0) cudaSetDevice(prev);
    if (rc == 0 && nnz_out) *nnz_out = R->nnz;
    return rc;
}

int bsm_sparse_fetch(bsm_handle h, int64_t *colptr, int64_t *rowval, void *nzval) {
    if (!h) return sfail(BSM_ERR_ARG, "null handle");
    SparseResult *R = bsm_sparse_slot(h);
    if (R->nnz < 0) return sfail(BSM_ERR_ARG, "call bsm_sparse_build first");
    if (!colptr || (R->nnz > 0 && (!rowval || !nzval))) return sfail(BSM_ERR_ARG, "null output");
    const size_t s = R->dtype == BSM_F32 ? 4 : R->dtype == BSM_F64 ? 8 : 16;
    SP_TRY(cudaMemcpy(colptr, R->colptr, (size_t)(R->ncols + 1) * 8, cudaMemcpyDeviceToHost));
    if (R->nnz > 0) {
        SP_TRY(cudaMemcpy(rowval, R->rowval, (size_t)R->nnz * 8, cudaMemcpyDeviceToHost));
        SP_TRY(cudaMemcpy(nzval, R->nzval, (size_t)R->nnz * s, cudaMemcpyDeviceToHost));
    }
    release(R);
    return 0;
}

int bsm_sparse_device_pointers(bsm_handle h, void **colptr_dev, void **rowval_dev, void **nzval_dev, int64_t *nnz) {
    if (!h) return sfail(BSM_ERR_ARG, "null handle");
    SparseResult *R = bsm_sparse_slot(h);
    if (R->nnz < 0) return sfail(BSM_ERR_ARG, "call bsm_sparse_build first");
    if (colptr_dev) *colptr_dev = R->colptr;
    if (rowval_dev) *rowval_dev = R->rowval;
    if (nzval_dev) *nzval_dev = R->nzval;
    if (nnz) *nnz = R->nnz;
    return 0;
}

}  // extern "C"
