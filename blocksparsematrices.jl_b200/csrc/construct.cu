// construct.cu — device-side construction.
//
//   arena_gather_kernel     blocks that already live in HBM (a GPU assembly routine wrote them) are gathered into the
//                           arena on the device: no PCIe, no host staging. One CTA per piece of <= 64 KB of a block;
//                           16-byte vector copies when source and destination allow, lazy `transpose(parent)`
//                           wrappers (SymmetricBlockMatrix -> VBCRS conversion, /root/reference/src/vbcrs.jl:222-264)
//                           materialised on the fly.
//   bsm_vbcrs_sort_dev      the sorting constructor of VariableBlockCompressedRowStorage
//                           (/root/reference/src/vbcrs.jl:78-122) on the device: stable sort of the blocks by
//                           (row start, column start) — cub::DeviceRadixSort on the packed 64-bit key, stable —, block-row
//                           pointer and block-row start rows from the run heads.
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <string>
#include <vector>

#include "../../include/bsm_b200.h"
#include "plan.h"

void bsm_set_error(const std::string &msg);

namespace {

int cfail(int code, const std::string &msg) {
    bsm_set_error(msg);
    return code;
}
#define C_TRY(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) return cfail(BSM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct GatherJob {
    const unsigned char *src;   // device pointer of the block (or of the parent of a transposed block)
    int64_t dst;                // byte offset in the arena
    int64_t begin, end;         // byte range of the block this job copies
    int32_t m, n;               // transposed blocks only (arena block is m x n, parent n x m); m = 0: plain copy
    int32_t esize;
    int32_t pad;
};

__global__ void __launch_bounds__(256) arena_gather_kernel(const GatherJob *jobs, unsigned char *arena) {
    const GatherJob j = jobs[blockIdx.x];
    const int64_t bytes = j.end - j.begin;
    unsigned char *dst = arena + j.dst + j.begin;
    if (j.m == 0) {
        const unsigned char *src = j.src + j.begin;
        if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
            const int64_t nv = bytes >> 4;
            for (int64_t i = threadIdx.x; i < nv; i += blockDim.x)
                reinterpret_cast<int4 *>(dst)[i] = __ldcs(reinterpret_cast<const int4 *>(src) + i);
            for (int64_t i = (nv << 4) + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
        } else {
            for (int64_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
        }
    } else {
        // arena block A is m x n (ld m), parent P is n x m (ld n): A[i, c] = P[c, i]
        const int64_t e0 = j.begin / j.esize, cnt = bytes / j.esize;
        for (int64_t e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int64_t i = (e0 + e) % j.m, c = (e0 + e) / j.m;
            const unsigned char *s = j.src + (i * j.n + c) * j.esize;
            unsigned char *d = dst + e * j.esize;
            for (int b = 0; b < j.esize; ++b) d[b] = s[b];
        }
    }
}

__global__ void pack_keys_kernel(const int64_t *rs, const int64_t *cs, int64_t nb, uint64_t *keys, int64_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    keys[i] = ((uint64_t)rs[i] << 32) | (uint64_t)(uint32_t)cs[i];
    idx[i] = i;
}
__global__ void heads_kernel(const uint64_t *keys, int64_t nb, int64_t *head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    head[i] = (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0;
}
__global__ void rows_kernel(const uint64_t *keys, const int64_t *head, const int64_t *rowid, int64_t nb, int64_t *rowptr,
                            int64_t *rowindices, int64_t *colindices, int64_t *nbrows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    colindices[i] = (int64_t)(keys[i] & 0xffffffffull);
    if (head[i]) {
        rowptr[rowid[i]] = i + 1;                       // 1-based, as Julia holds it
        rowindices[rowid[i]] = (int64_t)(keys[i] >> 32);
    }
    if (i == nb - 1) {
        const int64_t nr = rowid[i] + head[i];
        rowptr[nr] = nb + 1;                            // sentinel
        *nbrows = nr;
    }
}

}  // namespace

// abi.cu: gathers DEVICE blocks into the arena (block b of `m x n` elements at element offset off[b])
int bsm_arena_gather_dev(void *arena, int esize, const std::vector<bsm::BlockSrc> &blocks, const std::vector<int64_t> &off,
                         cudaStream_t st) {
    std::vector<GatherJob> jobs;
    const int64_t piece = 64 << 10;
    for (size_t b = 0; b < blocks.size(); ++b) {
        const int64_t bytes = (int64_t)blocks[b].m * blocks[b].n * esize;
        for (int64_t p0 = 0; p0 < bytes; p0 += piece) {
            GatherJob j;
            j.src = (const unsigned char *)blocks[b].host;      // a device pointer on this path
            j.dst = off[b] * esize;
            j.begin = p0;
            j.end = std::min(bytes, p0 + piece);
            j.m = blocks[b].transposed ? blocks[b].m : 0;
            j.n = blocks[b].n;
            j.esize = esize;
            j.pad = 0;
            jobs.push_back(j);
        }
    }
    if (jobs.empty()) return 0;
    GatherJob *dj = nullptr;
    C_TRY(cudaMalloc((void **)&dj, jobs.size() * sizeof(GatherJob)));
    cudaError_t e = cudaMemcpyAsync(dj, jobs.data(), jobs.size() * sizeof(GatherJob), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        arena_gather_kernel<<<(unsigned)jobs.size(), 256, 0, st>>>(dj, (unsigned char *)arena);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // `jobs` (pageable) and dj must outlive the copy / kernel
    cudaFree(dj);
    if (e != cudaSuccess) return cfail(BSM_ERR_CUDA, std::string("arena gather: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int bsm_vbcrs_sort_dev(int64_t nb, const int64_t *rowstart_dev, const int64_t *colstart_dev, int64_t *perm_dev,
                                  int64_t *rowptr_dev, int64_t *rowindices_dev, int64_t *colindices_dev, int64_t *nbrows_out,
                                  void *stream) {
    if (nb < 0 || !nbrows_out) return cfail(BSM_ERR_ARG, "bad argument");
    *nbrows_out = 0;
    if (nb == 0) {
        if (rowptr_dev) {
            const int64_t one = 1;
            C_TRY(cudaMemcpy(rowptr_dev, &one, sizeof(one), cudaMemcpyHostToDevice));
        }
        return 0;
    }
    if (!rowstart_dev || !colstart_dev || !perm_dev || !rowptr_dev || !rowindices_dev || !colindices_dev)
        return cfail(BSM_ERR_ARG, "null array");
    if (nb >= (1ll << 31)) return cfail(BSM_ERR_UNSUPPORTED, "too many blocks");
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t *keys = nullptr, *keys2 = nullptr;
    int64_t *idx = nullptr, *head = nullptr, *rowid = nullptr, *nbrows = nullptr;
    void *tmp = nullptr;
    struct Cleanup {
        void **p[7];
        ~Cleanup() {
            for (auto q : p)
                if (*q) cudaFree(*q);
        }
    } cleanup{{(void **)&keys, (void **)&keys2, (void **)&idx, (void **)&head, (void **)&rowid, (void **)&nbrows, &tmp}};
    const size_t nbytes = (size_t)nb * 8;
    C_TRY(cudaMalloc((void **)&keys, nbytes));
    C_TRY(cudaMalloc((void **)&keys2, nbytes));
    C_TRY(cudaMalloc((void **)&idx, nbytes));
    C_TRY(cudaMalloc((void **)&head, nbytes));
    C_TRY(cudaMalloc((void **)&rowid, nbytes));
    C_TRY(cudaMalloc((void **)&nbrows, 8));
    const unsigned g = (unsigned)((nb + 255) / 256);
    pack_keys_kernel<<<g, 256, 0, st>>>(rowstart_dev, colstart_dev, nb, keys, idx);
    size_t t1 = 0, t2 = 0;
    C_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, keys, keys2, idx, perm_dev, (int)nb, 0, 64, st));
    C_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t2, head, rowid, (int)nb, st));
    C_TRY(cudaMalloc(&tmp, std::max(t1, t2)));
    // radix sort is stable: blocks with equal (row start, column start) keep their input order, as sortperm does
    C_TRY(cub::DeviceRadixSort::SortPairs(tmp, t1, keys, keys2, idx, perm_dev, (int)nb, 0, 64, st));
    heads_kernel<<<g, 256, 0, st>>>(keys2, nb, head);
    C_TRY(cub::DeviceScan::ExclusiveSum(tmp, t2, head, rowid, (int)nb, st));
    rows_kernel<<<g, 256, 0, st>>>(keys2, head, rowid, nb, rowptr_dev, rowindices_dev, colindices_dev, nbrows);
    C_TRY(cudaGetLastError());
    C_TRY(cudaMemcpyAsync(nbrows_out, nbrows, 8, cudaMemcpyDeviceToHost, st));
    C_TRY(cudaStreamSynchronize(st));
    return 0;
}
