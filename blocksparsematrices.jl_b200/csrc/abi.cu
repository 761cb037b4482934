// abi.cu — extern "C" entry points of libbsm_b200.so (see include/bsm_b200.h).
// Front-ends lower the three reference storage types to contributions, pack.cpp builds the plans,
// this file owns the device copies and launches the kernels. No CPU fallback: every compute entry
// point needs a CUDA device and fails with BSM_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "persist.cuh"
#include "plan.h"
#include "spmm.cuh"
#include "spmm_tma.cuh"

using namespace bsm;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

}  // namespace

void bsm_set_error(const std::string &msg) { g_err = msg; }   // used by dist.cu

namespace {

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(BSM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
    } while (0)

template <class U>
struct DevBuf {
    U *p = nullptr;
    int64_t n = 0;
    int upload(const std::vector<U> &h) {
        n = (int64_t)h.size();
        if (n == 0) {
            // keep a valid non-null pointer so kernels can form addresses
            CUDA_TRY(cudaMalloc((void **)&p, 16));
            return 0;
        }
        CUDA_TRY(cudaMalloc((void **)&p, (size_t)n * sizeof(U)));
        CUDA_TRY(cudaMemcpy(p, h.data(), (size_t)n * sizeof(U), cudaMemcpyHostToDevice));
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
    }
};

struct DevPlan {
    DevBuf<bsm_contrib> contrib;
    DevBuf<int64_t> contrib_toff;
    DevBuf<bsm_slice> slices;
    DevBuf<int32_t> gather_rows;
    DevBuf<int64_t> gather_ptr, gather_pos;
    DevBuf<bsm_wchunk> wchunk;
    DevBuf<int32_t> witem_ptr, mitem_ptr, muncovered;
    DevBuf<bsm_slice> mslices;
    // row pieces of tall N-form blocks (CTA-stream kernel, real element types): tensor maps, built on first use
    DevBuf<int32_t> contrib_map;
    DevBuf<unsigned char> piece_maps;
    int piece_state = 0;   // 0: not looked at, 1: maps in place, 2: the plan has no such pieces (or no tensor maps here)
    void release() {
        contrib_map.release();
        piece_maps.release();
        wchunk.release();
        witem_ptr.release();
        mitem_ptr.release();
        muncovered.release();
        mslices.release();
        contrib.release();
        contrib_toff.release();
        slices.release();
        gather_rows.release();
        gather_ptr.release();
        gather_pos.release();
    }
};

}  // namespace

struct bsm_matrix {
    HostMatrix H;  // block sources are cleared after upload; tables stay for export
    int device = 0;
    int variant = BSM_VARIANT_AUTO;
    void *arena = nullptr;
    DevBuf<int32_t> set_len, set_start, pool;
    DevBuf<int64_t> set_pool_off;
    DevPlan plan[6];
    // host-pointer path
    std::mutex host_mu;
    void *hx = nullptr, *hy = nullptr;
    int64_t hx_bytes = 0, hy_bytes = 0;
    cudaStream_t host_stream = nullptr;
    // sparse(A) result built by bsm_sparse_build (sparse.cu), device arrays
    void *sparse_slot[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool restricted = false;
    int64_t own_lo[2] = {0, 0}, own_hi[2] = {-1, -1};   // owned outputs of op N / op T, C (hi < 0: all)
    bool blocks_on_device = false;  // the block sources of the pending upload are device pointers
    int64_t plan_hints = 0;
    // tensor maps of the arena for spmm_tma_kernel (encoded on the first multi-RHS multiply)
    TmaMaps tma_maps;
    bool tma_maps_ready = false;
    std::mutex tma_mu;
    // partial-sum scratch: one persistent buffer per launch stream (sized for the largest plan), so that a multiply
    // allocates nothing and concurrent multiplies on different streams never share partial sums
    std::mutex scratch_mu;
    std::vector<std::pair<cudaStream_t, void *>> scratch_by_stream;
    size_t scratch_bytes = 0;
    // benchmarking: events around the kernels of the last bsm_mul
    bool profiling = false;
    // a ring of event triples: every bsm_mul while profiling is on records into the next slot, so a whole timed
    // region of back-to-back multiplies can be read afterwards (bsm_get_profile averages and resets)
    static constexpr int kProfSlots = 64;
    std::vector<cudaEvent_t> ev_ring;
    cudaEvent_t *ev = nullptr;      // the slot of the multiply being launched
    int64_t prof_count = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int resolve_device(const bsm_options *opt, int *dev) {
    if (opt && opt->device == BSM_DEVICE_NONE) {
        *dev = BSM_DEVICE_NONE;
        return 0;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(BSM_ERR_CUDA, std::string("no CUDA device available (libbsm_b200 has no CPU fallback): ") +
                                      cudaGetErrorString(e));
    int d = opt ? opt->device : -1;
    if (d < 0) CUDA_TRY(cudaGetDevice(&d));
    if (d >= count) return fail(BSM_ERR_ARG, "device ordinal out of range");
    *dev = d;
    return 0;
}

}  // namespace
int bsm_arena_gather_dev(void *arena, int esize, const std::vector<bsm::BlockSrc> &blocks, const std::vector<int64_t> &off,
                         cudaStream_t st);   // construct.cu
namespace {

// Copies every block into the device arena through two pinned staging buffers (host blocks), or gathers them in HBM
// (device blocks).
int upload_arena(bsm_matrix *A) {
    HostMatrix &H = A->H;
    const int64_t s = dtype_size(H.dtype);
    if (!A->arena) {   // first upload (bsm_update_values reuses the arena: padding stays zero)
        CUDA_TRY(cudaMalloc(&A->arena, (size_t)H.arena_elems * s));
        CUDA_TRY(cudaMemset(A->arena, 0, (size_t)H.arena_elems * s));
    }
    if (A->blocks_on_device) return bsm_arena_gather_dev(A->arena, (int)s, H.blocks, H.block_off, nullptr);
    const int64_t stage_bytes = 64ll << 20;
    unsigned char *stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    struct Cleanup {   // staging resources are released on every path out of this function
        unsigned char **stage;
        cudaEvent_t *done;
        cudaStream_t *st;
        ~Cleanup() {
            if (*st) cudaStreamSynchronize(*st);
            for (int i = 0; i < 2; ++i) {
                if (stage[i]) cudaFreeHost(stage[i]);
                if (done[i]) cudaEventDestroy(done[i]);
            }
            if (*st) cudaStreamDestroy(*st);
        }
    } cleanup{stage, done, &st};
    CUDA_TRY(cudaStreamCreate(&st));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaMallocHost((void **)&stage[i], (size_t)stage_bytes));
        CUDA_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    int cur = 0;
    int64_t fill = 0;        // bytes used in stage[cur]
    int64_t dst_off = -1;    // arena byte offset the staged run starts at
    // Host blocks are separate heap allocations (Julia `Vector{Matrix{T}}`): filling a pinned staging buffer is a
    // gather of many memcpys, which one thread cannot feed at PCIe speed. The copies of one staging buffer are
    // recorded as jobs (large ones cut into 1 MB pieces) and executed by a few threads while the previous buffer
    // is in flight.
    struct Job {
        unsigned char *dst;
        const unsigned char *src;
        int64_t bytes;      // plain copy when m == 0
        int64_t e0;         // transposed block: first arena element of the piece
        int32_t m, n;
    };
    std::vector<Job> jobs;
    const unsigned nthreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    auto run_jobs = [&]() {
        if (jobs.empty()) return;
        auto work = [&](unsigned t) {
            for (size_t k = t; k < jobs.size(); k += nthreads) {
                const Job &j = jobs[k];
                if (j.m == 0) {
                    std::memcpy(j.dst, j.src, (size_t)j.bytes);
                } else {
                    // arena block is m x n (ld m); host parent is n x m (ld n): A[i,j] = Parent[j,i]
                    const int64_t cnt = j.bytes / s;
                    for (int64_t e = 0; e < cnt; ++e) {
                        const int64_t i = (j.e0 + e) % j.m, c = (j.e0 + e) / j.m;
                        std::memcpy(j.dst + e * s, j.src + (i * j.n + c) * s, (size_t)s);
                    }
                }
            }
        };
        int64_t total = 0;
        for (const Job &j : jobs) total += j.bytes;
        if (nthreads == 1 || total < (4 << 20)) {
            for (unsigned t = 0; t < nthreads; ++t) work(t);
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto &th : pool) th.join();
        }
        jobs.clear();
    };
    auto flush = [&]() -> int {
        if (fill == 0) return 0;
        run_jobs();
        CUDA_TRY(cudaMemcpyAsync((unsigned char *)A->arena + dst_off, stage[cur], (size_t)fill,
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(done[cur], st));
        cur ^= 1;
        CUDA_TRY(cudaEventSynchronize(done[cur]));
        fill = 0;
        dst_off = -1;
        return 0;
    };
    const int64_t piece = 1 << 20;
    for (size_t b = 0; b < H.blocks.size(); ++b) {
        const BlockSrc &src = H.blocks[b];
        const int64_t bytes = (int64_t)src.m * src.n * s;
        const int64_t boff = H.block_off[b] * s;
        int64_t done_b = 0;
        while (done_b < bytes) {
            // staged runs must be contiguous in the arena: the gap between blocks is alignment
            // padding (already zero), so a run simply continues at boff + done_b
            if (fill > 0 && dst_off + fill != boff + done_b) {
                const int64_t gap = boff + done_b - (dst_off + fill);
                if (gap > 0 && fill + gap < stage_bytes) {
                    std::memset(stage[cur] + fill, 0, (size_t)gap);
                    fill += gap;
                } else if (int rc = flush()) {
                    return rc;
                }
            }
            if (fill == 0) dst_off = boff + done_b;
            const int64_t take = std::min(std::min(bytes - done_b, stage_bytes - fill), piece);
            Job j;
            j.dst = stage[cur] + fill;
            j.bytes = take;
            if (!src.transposed) {
                j.src = (const unsigned char *)src.host + done_b;
                j.e0 = 0;
                j.m = j.n = 0;
            } else {
                j.src = (const unsigned char *)src.host;
                j.e0 = done_b / s;
                j.m = src.m;
                j.n = src.n;
            }
            jobs.push_back(j);
            fill += take;
            done_b += take;
            if (fill == stage_bytes)
                if (int rc = flush()) return rc;
        }
    }
    if (int rc = flush()) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int upload_tables(bsm_matrix *A) {
    HostMatrix &H = A->H;
    if (int rc = A->set_len.upload(H.sets.len)) return rc;
    if (int rc = A->set_start.upload(H.sets.start)) return rc;
    if (int rc = A->set_pool_off.upload(H.sets.pool_off)) return rc;
    if (int rc = A->pool.upload(H.sets.pool)) return rc;
    for (int p = 0; p < 6; ++p) {
        if (p >= 2 && p < 4 && !H.has_fused) continue;
        if (p >= 4 && !H.plan[p].color_ok) continue;
        if (int rc = A->plan[p].contrib.upload(H.plan[p].contrib)) return rc;
        if (int rc = A->plan[p].contrib_toff.upload(H.plan[p].contrib_toff)) return rc;
        if (int rc = A->plan[p].slices.upload(H.plan[p].slices)) return rc;
        if (int rc = A->plan[p].gather_rows.upload(H.plan[p].gather_rows)) return rc;
        if (int rc = A->plan[p].gather_ptr.upload(H.plan[p].gather_ptr)) return rc;
        if (int rc = A->plan[p].gather_pos.upload(H.plan[p].gather_pos)) return rc;
        if (int rc = A->plan[p].wchunk.upload(H.plan[p].wchunk)) return rc;
        if (int rc = A->plan[p].witem_ptr.upload(H.plan[p].witem_ptr)) return rc;
        if (int rc = A->plan[p].mitem_ptr.upload(H.plan[p].mitem_ptr)) return rc;
        if (int rc = A->plan[p].muncovered.upload(H.plan[p].muncovered)) return rc;
        if (int rc = A->plan[p].mslices.upload(H.plan[p].mslices)) return rc;
    }
    return 0;
}

// Shared tail of the three create functions.
// ir[0], ir[1]: GATHER plans for op N and op T/C; ir[2], ir[3]: FUSED plans (only if H.has_fused).
int finish_create(bsm_matrix *A, const std::vector<ContribIR> *ir, const bsm_options *opt,
                  bsm_handle *out) {
    HostMatrix &H = A->H;
    if (H.nrows >= (1ll << 31) || H.ncols >= (1ll << 31)) {
        delete A;
        return fail(BSM_ERR_UNSUPPORTED, "matrix dimension does not fit Int32 device indices");
    }
    layout_arena(H);
    PlanParams pp[2];
    if (opt) {
        pp[0].own_lo = opt->own_row_lo;
        pp[0].own_hi = opt->own_row_hi;
        pp[1].own_lo = opt->own_col_lo;
        pp[1].own_hi = opt->own_col_hi;
        // op N reads x along the columns, op T/C along the rows: the x slab a rank owns before the all-gather
        pp[0].in_lo = opt->own_col_lo;
        pp[0].in_hi = opt->own_col_hi;
        pp[1].in_lo = opt->own_row_lo;
        pp[1].in_hi = opt->own_row_hi;
        A->variant = opt->variant;
        A->restricted = opt->own_row_hi >= 0 || opt->own_col_hi >= 0;
        A->own_lo[0] = opt->own_row_lo;
        A->own_hi[0] = opt->own_row_hi;
        A->own_lo[1] = opt->own_col_lo;
        A->own_hi[1] = opt->own_col_hi;
        A->blocks_on_device = opt->blocks_on_device != 0;
        A->plan_hints = opt->plan_hints;
    }
    // work-item budget of the stream plans: no CTA should hold more than ~1/8 of a resident slot's fair share (LPT tail <= ~6 %)
    const int64_t total_bytes = [&] {
        int64_t t = 0;
        for (const auto &b : H.blocks) t += (int64_t)b.m * b.n * dtype_size(H.dtype);
        return t;
    }();
    // BSM_TUNE_SPLIT_DIV / BSM_TUNE_WITEMS_PER_SLOT: development knobs for the work-item granularity (profiles/tools)
    const char *tune_split = std::getenv("BSM_TUNE_SPLIT_DIV"), *tune_witems = std::getenv("BSM_TUNE_WITEMS_PER_SLOT");
    const int64_t split_div = tune_split ? std::max(1, std::atoi(tune_split)) : 8;
    const int64_t split = std::max<int64_t>(256 << 10, total_bytes / (148 * 2 * split_div));
    if (tune_witems) pp[0].witems_per_slot = pp[1].witems_per_slot = std::max(1, std::atoi(tune_witems));
    // long segments cut into sub-ranges (1024-row blocks): 2/5 of a resident CTA slot's share per CTA, between
    // 256 KB and 1 MB — measured on the C4 matrix, transpose(A)*x: 0.151 / 0.141 / 0.130 / 0.147 ms kernel time for
    // 256 KB / 512 KB / 1 MB / 2 MB per CTA (2886 / 1600 / 820 / 426 CTAs)
    pp[0].work_target_bytes = pp[1].work_target_bytes =
        std::min<int64_t>(1 << 20, std::max<int64_t>(256 << 10, total_bytes * 2 / (148 * 2 * 5)));
    if (const char *tune_wt = std::getenv("BSM_TUNE_WORK_TARGET"))   // development knob: bytes per such CTA
        pp[0].work_target_bytes = pp[1].work_target_bytes = std::max<int64_t>(64 << 10, std::atoll(tune_wt));
    if (const char *tune_wchunk = std::getenv("BSM_TUNE_WCHUNK"))   // fixed warp-stream chunk payload, 1024 .. 5120 bytes
        pp[0].wchunk_bytes = pp[1].wchunk_bytes = std::min(5120, std::max(1024, std::atoi(tune_wchunk) / 16 * 16));
    std::string err = build_plan(H, ir[0], H.nrows, H.ncols, pp[0], H.plan[0]);
    if (err.empty()) err = build_plan(H, ir[1], H.ncols, H.nrows, pp[1], H.plan[1]);
    if (H.has_fused) {
        pp[0].fused = pp[1].fused = true;
        pp[0].split_bytes = pp[1].split_bytes = split;
        // small (L2-resident, latency-bound) problems: one warp work item per ~1/2368 of the matrix, down to
        // single blocks, so that all 148 x 16 warps have something to stream
        if (total_bytes < (int64_t)148 * 16 * (32 << 10))
            pp[0].wsplit_bytes = pp[1].wsplit_bytes = std::max<int64_t>(4 << 10, total_bytes / (148 * 16));
        pp[0].warp_stream = pp[1].warp_stream = H.kind != BSM_KIND_SYMMETRIC;
        if (opt && (opt->plan_hints & 1)) pp[0].warp_stream = pp[1].warp_stream = false;
        if (opt && (opt->plan_hints & 2)) pp[0].wsplit_bytes = pp[1].wsplit_bytes = 0;
        if (opt && (opt->plan_hints & 4)) pp[0].wcta = pp[1].wcta = false;
        if (err.empty()) err = build_plan(H, ir[2], H.nrows, H.ncols, pp[0], H.plan[2]);
        if (err.empty()) err = build_plan(H, ir[3], H.ncols, H.nrows, pp[1], H.plan[3]);
    }
    // colour-ordered comparison variant: same contributions as the GATHER plans, launched colour by colour
    if (err.empty()) err = build_color_plan(H, ir[0], H.nrows, H.ncols, pp[0], H.plan[4]);
    if (err.empty()) err = build_color_plan(H, ir[1], H.ncols, H.nrows, pp[1], H.plan[5]);
    for (int p = 0; p < 4; ++p)
        A->scratch_bytes = std::max(A->scratch_bytes, (size_t)H.plan[p].scratch_elems * (size_t)dtype_size(H.dtype));
    if (!err.empty()) {
        delete A;
        return fail(BSM_ERR_ARG, err);
    }
    if (A->device == BSM_DEVICE_NONE) {  // host-only handle: tables only
        for (auto &b : H.blocks) b.host = nullptr;
        *out = A;
        return 0;
    }
    DeviceGuard g(A->device);
    if (!g.ok) {
        delete A;
        return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    }
    int rc = upload_arena(A);
    if (rc == 0) rc = upload_tables(A);
    if (rc != 0) {
        std::string keep = g_err;
        bsm_destroy(A);
        g_err = keep;
        return rc;
    }
    for (auto &b : H.blocks) b.host = nullptr;  // host blocks are not referenced after create
    A->blocks_on_device = false;
    *out = A;
    return 0;
}

// GATHER plans are 0/1, FUSED plans 2/3 (symmetric matrices; AUTO picks FUSED when it exists).
int plan_index(const bsm_matrix *A, int op) {
    const int base = (op == BSM_OP_N) ? 0 : 1;
    if (A->variant == BSM_VARIANT_COLOR && A->H.plan[4 + base].color_ok) return 4 + base;
    const bool fused = A->H.has_fused && (A->variant == BSM_VARIANT_AUTO || A->variant == BSM_VARIANT_FUSED ||
                                          A->variant == BSM_VARIANT_FUSED_TMA);
    return base + (fused ? 2 : 0);
}

int sm_count(int device) {
    static int cached[64] = {};
    int &c = cached[device & 63];
    if (c == 0) {
        int v = 0;
        c = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ? v : 148;
    }
    return c;
}

// Persistent partial-sum scratch of (handle, stream): allocated on the first multiply a stream sees.
int get_scratch(bsm_matrix *A, cudaStream_t st, void **out) {
    *out = nullptr;
    if (A->scratch_bytes == 0) return 0;
    std::lock_guard<std::mutex> lk(A->scratch_mu);
    for (auto &e : A->scratch_by_stream)
        if (e.first == st) {
            *out = e.second;
            return 0;
        }
    if (A->scratch_by_stream.size() >= 8) {   // streams come and go: recycle the oldest buffer (cudaFree waits for the device)
        cudaFree(A->scratch_by_stream.front().second);
        A->scratch_by_stream.erase(A->scratch_by_stream.begin());
    }
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, A->scratch_bytes));
    A->scratch_by_stream.emplace_back(st, p);
    *out = p;
    return 0;
}

// Dynamic shared-memory opt-in of every kernel instantiated for T, once per device (the attribute is per device).
template <class T>
int ensure_attrs(int device) {
    static std::mutex mu;
    static bool done[64] = {};
    std::lock_guard<std::mutex> lk(mu);
    if (done[device & 63]) return 0;
    CUDA_TRY(cudaFuncSetAttribute(sym_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fused_smem_bytes<T>()));
    CUDA_TRY(cudaFuncSetAttribute(sym_fused_tma_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fused_tma_smem_bytes<T>()));
    CUDA_TRY(cudaFuncSetAttribute(sym_fused_tma_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fused_tma_smem_bytes<T>()));
    CUDA_TRY(cudaFuncSetAttribute(sym_persist_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)PersistSmem<T>::total));
    CUDA_TRY(cudaFuncSetAttribute(stream_warp_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)stream_warp_smem_bytes<T>()));
    CUDA_TRY(cudaFuncSetAttribute(stream_warp_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)stream_warp_smem_bytes<T>()));
    CUDA_TRY(cudaFuncSetAttribute(stream_warp_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)stream_warp_smem_bytes<T>()));
    if constexpr (sizeof(T) == 8)
        CUDA_TRY(cudaFuncSetAttribute(spmm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)std::max(spmm_smem_bytes(true), spmm_smem_bytes(false))));
    done[device & 63] = true;
    return 0;
}

// ---- spmm_tma_kernel: tensor maps and launch -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int encode_2d(CUtensorMap *map, CUtensorMapDataType dt, void *base, uint64_t d0, uint64_t d1, uint64_t stride1_bytes,
              uint32_t b0, uint32_t b1, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(BSM_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t dims[2] = {d0, d1};
    const cuuint64_t strides[1] = {stride1_bytes};
    const cuuint32_t box[2] = {b0, b1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BSM_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 0;
}

template <class T>
int ensure_arena_maps(bsm_matrix *A) {
    std::lock_guard<std::mutex> lk(A->tma_mu);
    if (A->tma_maps_ready) return 0;
    const uint64_t rows = (uint64_t)A->H.arena_elems * sizeof(T) / 128;
    for (int q = 0; q < 4; ++q)   // the arena as rows of 128 bytes; box q covers ARows >> q rows
        if (int rc = encode_2d(&A->tma_maps.a[q], CU_TENSOR_MAP_DATA_TYPE_UINT8, A->arena, 128, rows, 128, 128,
                               (uint32_t)(TmaGeom<T>::ARows >> q)))
            return rc;
    A->tma_maps_ready = true;
    return 0;
}

// Row pieces of tall N-form blocks in the CTA-stream plans (pack.cpp step 4, "long segment fed by N-form contributions
// only"): per-column bulk copies of 1-2 KB run into the request rate of the copy engine (measured: 0.26 of the roofline
// on the C4 matrix), so for Float32 / Float64 a piece of a whole chunk of columns is fetched as ONE box of a 2-D tensor
// map. A map describes the arena as columns of m entries starting at the block's phase (offset mod m), so blocks of the
// same height and phase share it; contributions beyond kMaxPieceMaps maps keep the per-column copies.
constexpr size_t kMaxPieceMaps = 4096;
template <class T>
int ensure_piece_maps(bsm_matrix *A, const HostPlan &HP, DevPlan &DP) {
    std::lock_guard<std::mutex> lk(A->tma_mu);
    if (DP.piece_state) return 0;
    DP.piece_state = 2;
    if (sizeof(T) > 8 || !encode_tiled_fn()) return 0;
    std::vector<int32_t> cmap(HP.contrib.size(), -1);
    std::map<std::pair<int64_t, int64_t>, int32_t> keys;     // (m, phase in entries) -> map index
    std::vector<CUtensorMap> maps;
    const int64_t cc = kPChunk / (kFMaxRows * (int64_t)sizeof(T));
    for (int64_t i = 0; i < HP.n_fused_slices; ++i) {
        const bsm_slice &sl = HP.slices[(size_t)i];
        for (int32_t c = sl.c_begin; c < sl.c_end; ++c) {
            const bsm_contrib &cb = HP.contrib[(size_t)c];
            if ((cb.form & kFormT) || cb.m == 0 || cb.n == 0 || cmap[(size_t)c] >= 0) continue;
            if (!(sl.r0 > 0 || std::min<int32_t>(sl.r1, cb.out_len) < cb.m)) continue;
            if (cb.m < kFMaxRows || ((int64_t)cb.m * sizeof(T)) % 16 != 0) continue;
            const std::pair<int64_t, int64_t> key(cb.m, cb.off % cb.m);
            auto found = keys.find(key);
            int32_t k = found == keys.end() ? -1 : found->second;
            if (k < 0) {
                if (keys.size() >= kMaxPieceMaps) continue;
                const int64_t ncol = (A->H.arena_elems - key.second) / cb.m;
                CUtensorMap tm;
                if (int rc = encode_2d(&tm, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                       (unsigned char *)A->arena + key.second * sizeof(T), (uint64_t)cb.m, (uint64_t)ncol,
                                       (uint64_t)cb.m * sizeof(T), kFMaxRows, (uint32_t)cc, CU_TENSOR_MAP_SWIZZLE_NONE))
                    return rc;
                k = (int32_t)maps.size();
                keys.emplace(key, k);
                maps.push_back(tm);
            }
            cmap[(size_t)c] = k;
        }
    }
    if (maps.empty()) return 0;
    std::vector<unsigned char> raw(maps.size() * sizeof(CUtensorMap));
    std::memcpy(raw.data(), maps.data(), raw.size());
    if (int rc = DP.contrib_map.upload(cmap)) return rc;
    if (int rc = DP.piece_maps.upload(raw)) return rc;
    DP.piece_state = 1;
    return 0;
}

template <class T, int NB>
int launch_spmm_tma_nb(bsm_matrix *A, const HostPlan &HP, const DevPlan &DP, const TmaSpmmArgs<T> &args, const T *x, int64_t ldx,
                       int64_t nin, int64_t nrhs, cudaStream_t st) {
    static std::mutex mu;
    static bool attr_done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!attr_done[A->device & 63]) {
            CUDA_TRY(cudaFuncSetAttribute(spmm_tma_kernel<T, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)spmm_tma_smem_bytes<T, NB>()));
            attr_done[A->device & 63] = true;
        }
    }
    TmaMaps maps = A->tma_maps;
    constexpr bool is_c = sizeof(T) == 16;
    constexpr bool is_f = sizeof(T) == 4;
    // X as (contraction entries) x (right-hand sides); ComplexF64 is described as interleaved doubles
    if (int rc = encode_2d(&maps.x, is_f ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (void *)x,
                           (uint64_t)nin * (is_c ? 2 : 1), (uint64_t)nrhs, (uint64_t)ldx * sizeof(T), is_f ? 32 : 16, NB))
        return rc;
    TmaSpmmArgs<T> a = args;
    a.nstages = spmm_tma_stages<T, NB>();
    dim3 grid((unsigned)(HP.mitem_ptr.size() - 1), (unsigned)((nrhs + NB - 1) / NB));
    spmm_tma_kernel<T, NB><<<grid, kTThreads, spmm_tma_smem_bytes<T, NB>(), st>>>(a, maps);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Multi-RHS multiply through spmm_tma_kernel. Returns 1 when the call is not eligible (the caller falls back).
template <class T>
int launch_spmm_tma(bsm_matrix *A, int op, const HostPlan &HP, const DevPlan &DP, const void *alpha, const void *beta,
                    int beta_is_false, const T *x, int64_t ldx, T *y, int64_t ldy, int64_t nrhs, cudaStream_t st) {
    if (!HP.spmm_tma || !encode_tiled_fn()) return 1;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || ((uint64_t)ldx * sizeof(T)) % 16 != 0) return 1;   // TMA alignment rules
    if (int rc = ensure_arena_maps<T>(A)) return rc;
    TmaSpmmArgs<T> a;
    a.contrib = DP.contrib.p;
    a.slices = DP.mslices.p;
    a.item_ptr = DP.mitem_ptr.p;
    a.set_start = A->set_start.p;
    a.set_pool_off = A->set_pool_off.p;
    a.pool = A->pool.p;
    a.y = y;
    a.ldy = ldy;
    std::memcpy(&a.alpha, alpha, sizeof(T));
    std::memset(&a.beta, 0, sizeof(T));
    if (!beta_is_false) std::memcpy(&a.beta, beta, sizeof(T));
    a.nrhs = (int32_t)nrhs;
    a.beta_false = beta_is_false ? 1 : 0;
    a.conj = (op == BSM_OP_C && sizeof(T) == 16) ? 1 : 0;
    a.nstages = 0;
    a.elem_shift = sizeof(T) == 4 ? 2 : sizeof(T) == 8 ? 3 : 4;
    const int64_t nin = HP.in_dim;
    const bool prof = A->profiling;
    if (prof) CUDA_TRY(cudaEventRecord(A->ev[0], st));
    int rc;
    constexpr int kMaxNB = sizeof(T) == 16 ? 32 : 64;
    const int64_t want = std::min<int64_t>(nrhs, kMaxNB);
    if (want <= 8)
        rc = launch_spmm_tma_nb<T, 8>(A, HP, DP, a, x, ldx, nin, nrhs, st);
    else if (want <= 16)
        rc = launch_spmm_tma_nb<T, 16>(A, HP, DP, a, x, ldx, nin, nrhs, st);
    else if (want <= 32 || kMaxNB == 32)
        rc = launch_spmm_tma_nb<T, 32>(A, HP, DP, a, x, ldx, nin, nrhs, st);
    else
        rc = launch_spmm_tma_nb<T, (sizeof(T) == 16 ? 32 : 64)>(A, HP, DP, a, x, ldx, nin, nrhs, st);
    if (rc) return rc;
    if (prof) CUDA_TRY(cudaEventRecord(A->ev[1], st));
    const int64_t nu = (int64_t)HP.muncovered.size();
    if (nu > 0) {   // rows no block touches: y = beta*y
        for (int64_t jb = 0; jb < nrhs; jb += 65535) {   // grid.y limit
            dim3 fg((unsigned)((nu + 255) / 256), (unsigned)std::min<int64_t>(65535, nrhs - jb));
            spmm_uncovered_kernel_t<T><<<fg, 256, 0, st>>>(DP.muncovered.p, nu, y + jb * ldy, ldy, a.beta, a.beta_false);
        }
        CUDA_TRY(cudaGetLastError());
    }
    if (prof) CUDA_TRY(cudaEventRecord(A->ev[2], st));
    return 0;
}

// phase 0: the whole multiply. Slab handles under bsm_mul_dist (nrhs = 1, scratch owned by the caller and passed
// through *scratch_io): phase 1 = the slices whose inputs are rank-local (run while x is being all-gathered),
// phase 2 = the remote slices (on a second stream, so the two grids fill each other's tails), phase 3 = the
// gather pass.
template <class T>
int launch_mul(bsm_matrix *A, int op, const void *alpha, const void *beta, int beta_is_false,
               const T *x, int64_t ldx, T *y, int64_t ldy, int64_t nrhs, cudaStream_t st, int phase = 0,
               void **scratch_io = nullptr, const PeerX *px = nullptr) {
    const int p = plan_index(A, op);
    const HostPlan &HP = A->H.plan[p];
    if (HP.fused_general && HP.n_fused_slices > 0 && !A->plan[p].piece_state)
        if (int rc = ensure_piece_maps<T>(A, HP, A->plan[p])) return rc;
    const DevPlan &DP = A->plan[p];
    T *scratch = nullptr;
    if (A->profiling && phase <= 1) {   // next slot of the event ring (one per multiply)
        A->ev = A->ev_ring.data() + 3 * (A->prof_count % bsm_matrix::kProfSlots);
        A->prof_count++;
    }
    if (phase != 0 && (nrhs != 1 || p >= 4 || !scratch_io)) return fail(BSM_ERR_ARG, "phased multiply needs nrhs = 1");
    if (px && px->npeer > 0 && (nrhs != 1 || p >= 4)) return fail(BSM_ERR_ARG, "peer-mode multiply needs nrhs = 1");
    const int32_t nfused = (int32_t)HP.n_fused_slices;
    const int32_t nwarp = (int32_t)HP.n_warp_slices;
    // the direct-load comparison kernel only knows whole segments with short T-form blocks
    const bool use_tma = A->variant != BSM_VARIANT_FUSED || HP.fused_general;
    if (int rc = ensure_attrs<T>(A->device)) return rc;
    if (phase != 0)
        scratch = (T *)*scratch_io;
    else if (HP.scratch_elems > 0) {
        void *sp = nullptr;
        if (int rc = get_scratch(A, st, &sp)) return rc;
        scratch = (T *)sp;
    }
    // slice / item ranges of this phase
    const int32_t nitems_all = HP.witem_ptr.empty() ? 0 : (int32_t)HP.witem_ptr.size() - 1;
    const int32_t ngather_all = (int32_t)HP.slices.size() - nfused - nwarp;
    const int32_t f0 = phase == 2 ? (int32_t)HP.n_fused_local : 0;
    const int32_t f1 = phase == 3 ? 0 : phase == 1 ? (int32_t)HP.n_fused_local : nfused;
    const int32_t w0 = phase == 2 ? (int32_t)HP.n_warp_items_local : 0;
    const int32_t w1 = phase == 3 ? 0 : phase == 1 ? (int32_t)HP.n_warp_items_local : nitems_all;
    const int32_t g0 = phase == 2 ? (int32_t)HP.n_gather_local : 0;
    const int32_t g1 = phase == 3 ? 0 : phase == 1 ? (int32_t)HP.n_gather_local : ngather_all;
    constexpr int VMAX = 16 / (int)sizeof(T);
    // CTA-part plans whose only leftover work is "rows no block touches": folded into the warp-stream launch
    const bool fold_zero_rows = HP.wcta && HP.scratch_elems == 0 && phase == 0 && nfused == 0 && ngather_all == 0 &&
                                nitems_all > 0 && !HP.gather_rows.empty();
    if (p >= 4) {
        // colour-ordered variant: y <- beta*y, then one launch per (sweep, colour); the slices of a launch
        // touch disjoint rows, so each accumulates straight into y (beta = 1), as the reference's tasks do
        const int64_t nout = HP.out_dim;
        for (int64_t j = 0; j < nrhs; ++j) {
            MulArgs<T> a;
            a.arena = (const T *)A->arena;
            a.contrib = DP.contrib.p;
            a.contrib_toff = DP.contrib_toff.p;
            a.set_len = A->set_len.p;
            a.set_start = A->set_start.p;
            a.set_pool_off = A->set_pool_off.p;
            a.pool = A->pool.p;
            a.x.x = x + j * ldx;
            a.x.npeer = 0;
            a.y = y + j * ldy;
            a.scratch = nullptr;
            std::memcpy(&a.alpha, alpha, sizeof(T));
            T one;
            std::memset(&one, 0, sizeof(T));
            if constexpr (sizeof(T) == 4) {
                const float o = 1.f;
                std::memcpy(&one, &o, 4);
            } else {
                const double o = 1.0;
                std::memcpy(&one, &o, 8);
            }
            T b;
            std::memset(&b, 0, sizeof(T));
            if (!beta_is_false) std::memcpy(&b, beta, sizeof(T));
            const bool prof = A->profiling && nrhs == 1;
            if (prof) CUDA_TRY(cudaEventRecord(A->ev[0], st));
            if (nout > 0) {
                scale_kernel<T><<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(a.y, nout, b, beta_is_false ? 1 : 0);
                CUDA_TRY(cudaGetLastError());
            }
            a.beta = one;
            a.beta_false = 0;
            a.conj = (op == BSM_OP_C && sizeof(T) == 16) ? 1 : 0;
            for (size_t l = 0; l + 1 < HP.color_ptr.size(); ++l) {
                const int32_t s0 = HP.color_ptr[l], s1 = HP.color_ptr[l + 1];
                if (s1 == s0) continue;
                a.slices = DP.slices.p + s0;
                a.nslices = s1 - s0;
                gather_gemv_kernel<T, VMAX><<<a.nslices, kThreads, 0, st>>>(a);
                CUDA_TRY(cudaGetLastError());
            }
            if (prof) {
                CUDA_TRY(cudaEventRecord(A->ev[1], st));
                CUDA_TRY(cudaEventRecord(A->ev[2], st));
            }
        }
        return 0;
    }
    // many right-hand sides: one pass over A on the FP64 tensor cores instead of nrhs SpMV passes. Regular plans (blocks
    // of <= 32 rows, contiguous input sets) of every dtype: spmm_tma_kernel; other Float64 plans: spmm_dmma_kernel
    if (nrhs >= kSpmmMinRhs && HP.spmm_ok && HP.spmm_tma && A->variant != BSM_VARIANT_GATHER && p >= 2 && p < 4 && phase == 0 &&
        !(px && px->npeer > 0) && A->variant != BSM_VARIANT_FUSED) {   // FUSED: comparison, round-1 SpMM kernel
        const int rc = launch_spmm_tma<T>(A, op, HP, DP, alpha, beta, beta_is_false, x, ldx, y, ldy, nrhs, st);
        if (rc <= 0) return rc;
    }
    if constexpr (sizeof(T) == 8) {
        if (nrhs >= kSpmmMinRhs && HP.spmm_ok && A->variant != BSM_VARIANT_GATHER && p >= 2) {
            SpmmArgs m;
            m.arena = (const double *)A->arena;
            m.contrib = DP.contrib.p;
            m.slices = DP.mslices.p;
            m.item_ptr = DP.mitem_ptr.p;
            m.set_start = A->set_start.p;
            m.set_pool_off = A->set_pool_off.p;
            m.pool = A->pool.p;
            m.x = (const double *)x;
            m.y = (double *)y;
            m.ldx = ldx;
            m.ldy = ldy;
            std::memcpy(&m.alpha, alpha, 8);
            m.beta = 0.0;
            if (!beta_is_false) std::memcpy(&m.beta, beta, 8);
            m.nrhs = (int32_t)nrhs;
            m.beta_false = beta_is_false ? 1 : 0;
            m.abytes = HP.spmm_small ? kMABytesSmall : kMABytesBig;
            m.nstages = spmm_stages(HP.spmm_small);
            const bool prof = A->profiling;
            if (prof) CUDA_TRY(cudaEventRecord(A->ev[0], st));
            dim3 grid((unsigned)(HP.mitem_ptr.size() - 1), (unsigned)((nrhs + kMRhs - 1) / kMRhs));
            spmm_dmma_kernel<<<grid, kMThreads, spmm_smem_bytes(HP.spmm_small), st>>>(m);
            CUDA_TRY(cudaGetLastError());
            if (prof) CUDA_TRY(cudaEventRecord(A->ev[1], st));
            const int64_t nu = (int64_t)HP.muncovered.size();
            if (nu > 0) {   // rows no block touches: y = beta*y
                for (int64_t jb = 0; jb < nrhs; jb += 65535) {   // grid.y limit
                    dim3 fg((unsigned)((nu + 255) / 256), (unsigned)std::min<int64_t>(65535, nrhs - jb));
                    spmm_uncovered_kernel<<<fg, 256, 0, st>>>(DP.muncovered.p, nu, (double *)y + jb * ldy, ldy, m.beta,
                                                              m.beta_false);
                }
                CUDA_TRY(cudaGetLastError());
            }
            if (prof) CUDA_TRY(cudaEventRecord(A->ev[2], st));
            return 0;
        }
    }
    for (int64_t j = 0; j < nrhs; ++j) {
        MulArgs<T> a;
        a.arena = (const T *)A->arena;
        a.contrib = DP.contrib.p;
        a.contrib_toff = DP.contrib_toff.p;
        a.slices = DP.slices.p;
        a.set_len = A->set_len.p;
        a.set_start = A->set_start.p;
        a.set_pool_off = A->set_pool_off.p;
        a.pool = A->pool.p;
        a.x.x = x + j * ldx;
        a.x.npeer = 0;
        // the last kernel that reads x runs the exit barrier of peer mode
        const int last_kind = (g1 > g0) ? 2 : (w1 > w0) ? 1 : (f1 > f0) ? 0 : -1;
        if (px && px->npeer > 0) {   // peer mode (nrhs = 1): element i comes from its owner's array
            a.x.npeer = px->npeer;
            for (int r = 0; r < px->npeer; ++r) a.x.peer[r] = (const T *)px->peer[r];
            for (int r = 0; r <= px->npeer; ++r) a.x.cuts[r] = px->cuts[r];
            a.x.sync.peer_flags = px->peer_flags;
            a.x.sync.my_flags = px->my_flags;
            a.x.sync.state = px->state;
            a.x.sync.nranks = px->npeer;
            a.x.sync.rank = px->rank;
            a.x.sync.do_exit = 0;
            a.x.sync.arrivals = 1;
            a.x.sync.debug = px->debug;
            a.x.sync.dbg = px->dbg;
            if (last_kind < 0) {   // nothing to multiply on this rank: the barriers alone
                PeerSync ps = a.x.sync;
                ps.do_exit = 1;
                peer_sync_kernel<<<1, 32, 0, st>>>(ps);
                CUDA_TRY(cudaGetLastError());
            }
        }
        a.y = y + j * ldy;
        a.scratch = scratch;
        std::memcpy(&a.alpha, alpha, sizeof(T));
        if (beta_is_false)
            std::memset(&a.beta, 0, sizeof(T));
        else
            std::memcpy(&a.beta, beta, sizeof(T));
        a.nslices = (int32_t)HP.slices.size();
        a.beta_false = beta_is_false ? 1 : 0;
        a.conj = (op == BSM_OP_C && sizeof(T) == 16) ? 1 : 0;
        if (DP.piece_state == 1) {
            a.contrib_map = DP.contrib_map.p;
            a.piece_maps = DP.piece_maps.p;
        }
        const bool prof = A->profiling && nrhs == 1;
        if (A->profiling && nrhs > 1 && j == 0) CUDA_TRY(cudaEventRecord(A->ev[0], st));
        if (prof) CUDA_TRY(cudaEventRecord(A->ev[0], st));
        if (f1 > f0) {
            MulArgs<T> b = a;
            b.slices = a.slices + f0;
            if (b.x.npeer && last_kind == 0) {
                b.x.sync.do_exit = 1;
                b.x.sync.arrivals = f1 - f0;
            }
            if (use_tma)
                if (HP.fused_general)
                    sym_fused_tma_kernel<T, true><<<f1 - f0, kPThreads, fused_tma_smem_bytes<T>(), st>>>(b);
                else if (!(A->plan_hints & 8) && !(std::getenv("BSM_TUNE_PERSIST") && std::atoi(std::getenv("BSM_TUNE_PERSIST"))))   // default: one CTA per work item
                    sym_fused_tma_kernel<T, false><<<f1 - f0, kPThreads, fused_tma_smem_bytes<T>(), st>>>(b);
                else {
                    // experimental (plan_hints bit 3 / BSM_TUNE_PERSIST): persistent CTAs (two per SM) walking the work
                    // items, helper warps staging x and metadata ahead — correct, but 5-7 % slower than the per-item CTAs on
                    // C2 (profiles/README.md, "measured and rejected")
                    b.nslices = f1 - f0;
                    const char *tune_grid = std::getenv("BSM_TUNE_PERSIST_CTAS_PER_SM");   // 0: one CTA per item
                    const int per_sm = tune_grid ? std::atoi(tune_grid) : 2;
                    const int32_t pgrid = per_sm > 0 ? std::min<int32_t>(f1 - f0, per_sm * sm_count(A->device)) : f1 - f0;
                    if (b.x.npeer && last_kind == 0) b.x.sync.arrivals = pgrid;
                    sym_persist_kernel<T><<<pgrid, kQThreads, PersistSmem<T>::total, st>>>(b);
                }
            else
                sym_fused_kernel<T><<<f1 - f0, kFThreads, fused_smem_bytes<T>(), st>>>(b);
            CUDA_TRY(cudaGetLastError());
        }
        if (w1 > w0) {
            WarpArgs<T> w;
            w.arena = (const unsigned char *)A->arena;
            w.chunks = DP.wchunk.p;
            w.item_ptr = DP.witem_ptr.p + w0;
            w.pool = A->pool.p;
            w.x = a.x;
            w.y = a.y;
            w.scratch = scratch;
            w.alpha = a.alpha;
            w.beta = a.beta;
            w.nitems = w1 - w0;
            w.beta_false = a.beta_false;
            w.conj = a.conj;
            w.cta_mode = HP.wcta ? 1 : 0;
            w.nz = 0;
            w.zrows = nullptr;
            w.x_bulk_len = 0;
            {
                bool aligned = (reinterpret_cast<uintptr_t>(w.x.x) & 15) == 0;
                for (int r = 0; r < w.x.npeer; ++r) aligned = aligned && (reinterpret_cast<uintptr_t>(w.x.peer[r]) & 15) == 0;
                static const bool no_peer_bulk = std::getenv("BSM_TUNE_NO_XBULK_PEER") && std::atoi(std::getenv("BSM_TUNE_NO_XBULK_PEER"));
                if (w.x.npeer && no_peer_bulk) aligned = false;   // development knob: per-entry copies of remote x
                if (aligned) w.x_bulk_len = (int32_t)(HP.in_dim / (16 / (int64_t)sizeof(T)) * (16 / (int64_t)sizeof(T)));
            }
            if (fold_zero_rows) {   // single launch: the rows no block touches are set by extra CTAs of this kernel
                w.nz = (int32_t)HP.gather_rows.size();
                w.zrows = DP.gather_rows.p;
            }
            if (w.x.npeer && last_kind == 1) {
                w.x.sync.do_exit = 1;
                w.x.sync.arrivals = (int32_t)((w.nitems + kWWarps - 1) / kWWarps) * kWWarps;
            }
            const unsigned zctas = (unsigned)((w.nz + kWThreads - 1) / kWThreads);
            const unsigned wgrid = (unsigned)((w.nitems + kWWarps - 1) / kWWarps) + zctas;
            if (HP.wform == 0)
                stream_warp_kernel<T, 0><<<wgrid, kWThreads, stream_warp_smem_bytes<T>(), st>>>(w);
            else if (HP.wform == 1)
                stream_warp_kernel<T, 1><<<wgrid, kWThreads, stream_warp_smem_bytes<T>(), st>>>(w);
            else
                stream_warp_kernel<T, 2><<<wgrid, kWThreads, stream_warp_smem_bytes<T>(), st>>>(w);
            CUDA_TRY(cudaGetLastError());
        }
        if (g1 > g0) {
            MulArgs<T> b = a;
            b.slices = a.slices + nfused + nwarp + g0;
            b.nslices = g1 - g0;
            if (b.x.npeer) {
                b.x.sync.do_exit = 1;
                b.x.sync.arrivals = b.nslices;
            }
            gather_gemv_kernel<T, VMAX><<<b.nslices, kThreads, 0, st>>>(b);
            CUDA_TRY(cudaGetLastError());
        }
        if (prof) CUDA_TRY(cudaEventRecord(A->ev[1], st));
        const int64_t ng = (phase == 1 || phase == 2 || fold_zero_rows) ? 0 : (int64_t)HP.gather_rows.size();
        if (ng > 0) {
            FinalizeArgs<T> f;
            f.rows = DP.gather_rows.p;
            f.ptr = DP.gather_ptr.p;
            f.pos = DP.gather_pos.p;
            f.scratch = scratch;
            f.y = y + j * ldy;
            f.alpha = a.alpha;
            f.beta = a.beta;
            f.n = ng;
            f.ldy = 0;
            f.beta_false = a.beta_false;
            gather_finalize_kernel<T><<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(f);
            CUDA_TRY(cudaGetLastError());
        }
        if (prof) CUDA_TRY(cudaEventRecord(A->ev[2], st));
        if (A->profiling && nrhs > 1 && j == nrhs - 1) {   // column loop: one bracket around all columns
            CUDA_TRY(cudaEventRecord(A->ev[1], st));
            CUDA_TRY(cudaEventRecord(A->ev[2], st));
        }
    }
    return 0;
}

int check_handle(bsm_handle h) {
    if (!h) return fail(BSM_ERR_ARG, "null handle");
    return 0;
}

}  // namespace

// ================================================================================= C ABI

extern "C" {

const char *bsm_last_error(void) { return g_err.c_str(); }
const char *bsm_version(void) { return "bsm_b200 0.1.0 (sm_100a)"; }

void bsm_default_options(bsm_options *opt) {
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->device = -1;
    opt->variant = BSM_VARIANT_AUTO;
    opt->own_row_lo = 0;
    opt->own_row_hi = -1;
    opt->own_col_lo = 0;
    opt->own_col_hi = -1;
}

int bsm_create_blocksparse(int dtype, int64_t nrows, int64_t ncols, int64_t nb,
                           const void *const *blocks, const int64_t *m, const int64_t *n,
                           const int64_t *rowidx, const int64_t *rowptr, const int64_t *colidx,
                           const int64_t *colptr, const bsm_options *opt, bsm_handle *out) {
    if (!out) return fail(BSM_ERR_ARG, "out is null");
    *out = nullptr;
    if (dtype < 0 || dtype > 2) return fail(BSM_ERR_ARG, "bad dtype");
    if (nrows < 0 || ncols < 0 || nb < 0) return fail(BSM_ERR_ARG, "negative size");
    if (nb > 0 && (!blocks || !m || !n || !rowidx || !rowptr || !colidx || !colptr))
        return fail(BSM_ERR_ARG, "null array");
    int dev = 0;
    if (int rc = resolve_device(opt, &dev)) return rc;
    bsm_matrix *A = new (std::nothrow) bsm_matrix();
    if (!A) return fail(BSM_ERR_ALLOC, "out of host memory");
    A->device = dev;
    HostMatrix &H = A->H;
    H.dtype = dtype;
    H.kind = BSM_KIND_BLOCKSPARSE;
    H.nrows = nrows;
    H.ncols = ncols;
    H.has_fused = true;
    std::vector<ContribIR> ir[4];
    for (int64_t b = 0; b < nb; ++b) {
        if (m[b] < 0 || n[b] < 0 || rowptr[b + 1] - rowptr[b] != m[b] || colptr[b + 1] - colptr[b] != n[b] ||
            (m[b] * n[b] > 0 && !blocks[b])) {
            delete A;
            return fail(BSM_ERR_ARG, "block " + std::to_string(b) + ": size does not match its index vectors");
        }
        const int32_t rs = H.sets.add_vector(rowidx + rowptr[b], m[b], nrows);
        const int32_t cs = H.sets.add_vector(colidx + colptr[b], n[b], ncols);
        if (rs < 0 || cs < 0) {
            delete A;
            return fail(BSM_ERR_ARG, "block " + std::to_string(b) + ": index out of range");
        }
        H.blocks.push_back(BlockSrc{blocks[b], (int32_t)m[b], (int32_t)n[b], false});
        H.nnz += m[b] * n[b];
        // y[R_b] += op(B_b) x[C_b]   (/root/reference/src/blockmatrix.jl:236-242)
        ir[0].push_back(ContribIR{(int32_t)b, 0, rs, cs, (int32_t)m[b]});
        // wrappers swap the index vectors (/root/reference/src/symmetricblockmatrix.jl:345-365)
        ir[1].push_back(ContribIR{(int32_t)b, 1, cs, rs, (int32_t)n[b]});
    }
    ir[2] = ir[0];  // stream plans: same contributions, TMA-staged kernels
    ir[3] = ir[1];
    return finish_create(A, ir, opt, out);
}

int bsm_create_symmetric(int dtype, int64_t nrows, int64_t ncols, int64_t ndiag,
                         const void *const *diag, const int64_t *dsize, const int64_t *didx,
                         const int64_t *dptr, int64_t noff, const void *const *off,
                         const int64_t *om, const int64_t *on, const int64_t *rowidx,
                         const int64_t *rowptr, const int64_t *colidx, const int64_t *colptr,
                         const bsm_options *opt, bsm_handle *out) {
    if (!out) return fail(BSM_ERR_ARG, "out is null");
    *out = nullptr;
    if (dtype < 0 || dtype > 2) return fail(BSM_ERR_ARG, "bad dtype");
    if (nrows < 0 || ncols < 0 || ndiag < 0 || noff < 0) return fail(BSM_ERR_ARG, "negative size");
    if (nrows != ncols) return fail(BSM_ERR_ARG, "a SymmetricBlockMatrix must be square");
    if (ndiag > 0 && (!diag || !dsize || !didx || !dptr)) return fail(BSM_ERR_ARG, "null diagonal array");
    if (noff > 0 && (!off || !om || !on || !rowidx || !rowptr || !colidx || !colptr))
        return fail(BSM_ERR_ARG, "null off-diagonal array");
    int dev = 0;
    if (int rc = resolve_device(opt, &dev)) return rc;
    bsm_matrix *A = new (std::nothrow) bsm_matrix();
    if (!A) return fail(BSM_ERR_ALLOC, "out of host memory");
    A->device = dev;
    HostMatrix &H = A->H;
    H.dtype = dtype;
    H.kind = BSM_KIND_SYMMETRIC;
    H.nrows = nrows;
    H.ncols = ncols;
    H.has_fused = true;
    std::vector<ContribIR> ir[4];
    // diagonal sweep (/root/reference/src/symmetricblockmatrix.jl:420-432); emitted first so the
    // leaf segments claim their rows and are written directly
    for (int64_t d = 0; d < ndiag; ++d) {
        if (dsize[d] < 0 || dptr[d + 1] - dptr[d] != dsize[d] || (dsize[d] > 0 && !diag[d])) {
            delete A;
            return fail(BSM_ERR_ARG, "diagonal block " + std::to_string(d) + ": size mismatch");
        }
        const int32_t ds = H.sets.add_vector(didx + dptr[d], dsize[d], nrows);
        if (ds < 0) {
            delete A;
            return fail(BSM_ERR_ARG, "diagonal block " + std::to_string(d) + ": index out of range");
        }
        H.blocks.push_back(BlockSrc{diag[d], (int32_t)dsize[d], (int32_t)dsize[d], false});
        H.nnz += dsize[d] * dsize[d];
        ir[0].push_back(ContribIR{(int32_t)d, 0, ds, ds, (int32_t)dsize[d]});
        ir[1].push_back(ContribIR{(int32_t)d, 1, ds, ds, (int32_t)dsize[d]});  // transpose(D)/adjoint(D), :225-237
        ir[0].back().sweep = ir[1].back().sweep = 0;
        ir[2].push_back(ir[0].back());
        ir[3].push_back(ir[1].back());
    }
    std::vector<int32_t> rs((size_t)noff), cs((size_t)noff);
    for (int64_t b = 0; b < noff; ++b) {
        if (om[b] < 0 || on[b] < 0 || rowptr[b + 1] - rowptr[b] != om[b] || colptr[b + 1] - colptr[b] != on[b] ||
            (om[b] * on[b] > 0 && !off[b])) {
            delete A;
            return fail(BSM_ERR_ARG, "off-diagonal block " + std::to_string(b) + ": size mismatch");
        }
        rs[b] = H.sets.add_vector(rowidx + rowptr[b], om[b], nrows);
        cs[b] = H.sets.add_vector(colidx + colptr[b], on[b], ncols);
        if (rs[b] < 0 || cs[b] < 0) {
            delete A;
            return fail(BSM_ERR_ARG, "off-diagonal block " + std::to_string(b) + ": index out of range");
        }
        H.blocks.push_back(BlockSrc{off[b], (int32_t)om[b], (int32_t)on[b], false});
        H.nnz += 2 * om[b] * on[b];
    }
    // y[R_b] += O_b x[C_b]: sweep 1 of A (:394-405) and sweep 2 of the wrappers (:407-418, where
    // transpose(op(O_b)) is O_b or conj(O_b))
    for (int64_t b = 0; b < noff; ++b) {
        const int32_t blk = (int32_t)(ndiag + b);
        ir[0].push_back(ContribIR{blk, 0, rs[b], cs[b], (int32_t)om[b]});
        ir[1].push_back(ContribIR{blk, 0, rs[b], cs[b], (int32_t)om[b]});
        ir[0].back().sweep = ir[1].back().sweep = 1;
    }
    // y[C_b] += transpose(O_b) x[R_b]: sweep 2 of A and sweep 1 of the wrappers
    for (int64_t b = 0; b < noff; ++b) {
        const int32_t blk = (int32_t)(ndiag + b);
        ir[0].push_back(ContribIR{blk, 1, cs[b], rs[b], (int32_t)on[b]});
        ir[1].push_back(ContribIR{blk, 1, cs[b], rs[b], (int32_t)on[b]});
        ir[0].back().sweep = ir[1].back().sweep = 2;
    }
    // FUSED plans: both sweeps of a half-stored block from ONE pass over it, whenever its row segment
    // fits the fused kernel; taller blocks keep the two separate contributions
    for (int pl = 2; pl < 4; ++pl) {
        for (int64_t b = 0; b < noff; ++b) {
            const int32_t blk = (int32_t)(ndiag + b);
            ContribIR c{blk, 0, rs[b], cs[b], (int32_t)om[b]};
            if (om[b] <= kFusedMaxRows) c.fuse_tset = cs[b];
            ir[pl].push_back(c);
        }
        for (int64_t b = 0; b < noff; ++b)
            if (om[b] > kFusedMaxRows)
                ir[pl].push_back(ContribIR{(int32_t)(ndiag + b), 1, cs[b], rs[b], (int32_t)on[b]});
    }
    return finish_create(A, ir, opt, out);
}

int bsm_create_vbcrs(int dtype, int64_t nrows, int64_t ncols, int64_t nbrows, int64_t nb,
                     const int64_t *rowptr, const int64_t *colstart, const int64_t *rowstart,
                     const void *const *blocks, const int64_t *m, const int64_t *n,
                     const uint8_t *is_transposed, const bsm_options *opt, bsm_handle *out) {
    if (!out) return fail(BSM_ERR_ARG, "out is null");
    *out = nullptr;
    if (dtype < 0 || dtype > 2) return fail(BSM_ERR_ARG, "bad dtype");
    if (nrows < 0 || ncols < 0 || nbrows < 0 || nb < 0) return fail(BSM_ERR_ARG, "negative size");
    if (!rowptr) return fail(BSM_ERR_ARG, "rowptr is null");
    if (nb > 0 && (!colstart || !rowstart || !blocks || !m || !n)) return fail(BSM_ERR_ARG, "null array");
    if (rowptr[0] != 1 || rowptr[nbrows] != nb + 1) return fail(BSM_ERR_ARG, "rowptr must be 1-based with sentinel nb+1");
    int dev = 0;
    if (int rc = resolve_device(opt, &dev)) return rc;
    bsm_matrix *A = new (std::nothrow) bsm_matrix();
    if (!A) return fail(BSM_ERR_ALLOC, "out of host memory");
    A->device = dev;
    HostMatrix &H = A->H;
    H.dtype = dtype;
    H.kind = BSM_KIND_VBCRS;
    H.nrows = nrows;
    H.ncols = ncols;
    // segment lengths: a block row is keyed by its start row only and the height is taken per block
    // (/root/reference/src/vbcrs.jl:108, :280-282) → the row segment is as long as its tallest
    // block; likewise column segments are keyed by the start column
    std::vector<int32_t> brow_of((size_t)nb);
    std::vector<int64_t> rowlen((size_t)nbrows, 0);
    std::unordered_map<int64_t, int64_t> collen;
    for (int64_t r = 0; r < nbrows; ++r) {
        if (rowptr[r + 1] < rowptr[r]) {
            delete A;
            return fail(BSM_ERR_ARG, "rowptr must be non-decreasing");
        }
        for (int64_t b = rowptr[r] - 1; b < rowptr[r + 1] - 1; ++b) {
            if (m[b] < 0 || n[b] < 0 || (m[b] * n[b] > 0 && !blocks[b]) || rowstart[r] < 1 ||
                rowstart[r] + m[b] - 1 > nrows || colstart[b] < 1 || colstart[b] + n[b] - 1 > ncols) {
                delete A;
                return fail(BSM_ERR_ARG, "block " + std::to_string(b) + ": range outside the matrix");
            }
            brow_of[b] = (int32_t)r;
            rowlen[r] = std::max(rowlen[r], m[b]);
            auto it = collen.find(colstart[b]);
            if (it == collen.end())
                collen[colstart[b]] = n[b];
            else
                it->second = std::max(it->second, n[b]);
        }
    }
    // block rows with the same start row (legal, if unusual) share one segment
    std::unordered_map<int64_t, int64_t> rowlen_by_start;
    for (int64_t r = 0; r < nbrows; ++r) {
        auto it = rowlen_by_start.find(rowstart[r]);
        if (it == rowlen_by_start.end())
            rowlen_by_start[rowstart[r]] = rowlen[r];
        else
            it->second = std::max(it->second, rowlen[r]);
    }
    H.has_fused = true;
    std::vector<ContribIR> ir[4];
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t r = brow_of[b];
        const bool tr = is_transposed && is_transposed[b];
        H.blocks.push_back(BlockSrc{blocks[b], (int32_t)m[b], (int32_t)n[b], tr});
        H.nnz += m[b] * n[b];
        const int32_t out_rows = H.sets.add_range(rowstart[r] - 1, rowlen_by_start[rowstart[r]]);
        const int32_t in_cols = H.sets.add_range(colstart[b] - 1, n[b]);
        const int32_t out_cols = H.sets.add_range(colstart[b] - 1, collen[colstart[b]]);
        const int32_t in_rows = H.sets.add_range(rowstart[r] - 1, m[b]);
        ir[0].push_back(ContribIR{(int32_t)b, 0, out_rows, in_cols, (int32_t)m[b]});   // src/vbcrs.jl:277-284
        ir[1].push_back(ContribIR{(int32_t)b, 1, out_cols, in_rows, (int32_t)n[b]});   // src/vbcrs.jl:315-326
    }
    ir[2] = ir[0];  // stream plans: same contributions, TMA-staged kernels
    ir[3] = ir[1];
    return finish_create(A, ir, opt, out);
}

int bsm_update_values(bsm_handle h, const void *const *blocks, int64_t nb) {
    if (int rc = check_handle(h)) return rc;
    if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle has no arena");
    if (nb != (int64_t)h->H.blocks.size()) return fail(BSM_ERR_ARG, "block count differs from the handle's");
    if (nb > 0 && !blocks) return fail(BSM_ERR_ARG, "blocks is null");
    for (int64_t b = 0; b < nb; ++b) {
        if (!blocks[b] && (int64_t)h->H.blocks[b].m * h->H.blocks[b].n > 0) return fail(BSM_ERR_ARG, "null block");
        h->H.blocks[b].host = blocks[b];
    }
    DeviceGuard g(h->device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaDeviceSynchronize());   // no multiply may still be reading the arena
    const int rc = upload_arena(h);
    for (auto &b : h->H.blocks) b.host = nullptr;
    return rc;
}

int bsm_update_values_dev(bsm_handle h, const void *const *dev_blocks, int64_t nb) {
    if (int rc = check_handle(h)) return rc;
    h->blocks_on_device = true;
    const int rc = bsm_update_values(h, dev_blocks, nb);
    h->blocks_on_device = false;
    return rc;
}

int bsm_destroy(bsm_handle h) {
    if (!h) return 0;
    if (h->device == BSM_DEVICE_NONE) {
        delete h;
        return 0;
    }
    DeviceGuard g(h->device);
    if (h->arena) cudaFree(h->arena);
    if (h->sparse_slot[7]) {
        SparseResult *R = reinterpret_cast<SparseResult *>(h->sparse_slot);
        if (R->colptr) cudaFree(R->colptr);
        if (R->rowval) cudaFree(R->rowval);
        if (R->nzval) cudaFree(R->nzval);
    }
    h->set_len.release();
    h->set_start.release();
    h->set_pool_off.release();
    h->pool.release();
    for (int p = 0; p < 6; ++p) h->plan[p].release();
    if (h->hx) cudaFree(h->hx);
    if (h->hy) cudaFree(h->hy);
    for (auto &e : h->scratch_by_stream) cudaFree(e.second);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    for (cudaEvent_t e : h->ev_ring) cudaEventDestroy(e);
    delete h;
    return 0;
}

int bsm_set_variant(bsm_handle h, int variant) {
    if (int rc = check_handle(h)) return rc;
    if (variant < BSM_VARIANT_AUTO || variant > BSM_VARIANT_FUSED_TMA) return fail(BSM_ERR_ARG, "bad variant");
    if (variant == BSM_VARIANT_COLOR && !(h->H.plan[4].color_ok && h->H.plan[5].color_ok))
        return fail(BSM_ERR_UNSUPPORTED, "the colour-ordered variant needs an unrestricted handle, no repeated "
                                         "index inside a block's index vector and at most 64 colours per sweep");
    h->variant = variant;
    return 0;
}

int bsm_set_profiling(bsm_handle h, int on) {
    if (int rc = check_handle(h)) return rc;
    if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle");
    DeviceGuard g(h->device);
    if (on && h->ev_ring.empty()) {
        h->ev_ring.resize(3 * bsm_matrix::kProfSlots);
        for (auto &e : h->ev_ring) CUDA_TRY(cudaEventCreate(&e));
    }
    h->profiling = on != 0;
    h->prof_count = 0;
    h->ev = h->ev_ring.empty() ? nullptr : h->ev_ring.data();
    return 0;
}

int bsm_get_profile(bsm_handle h, double *main_ms, double *finalize_ms) {
    if (int rc = check_handle(h)) return rc;
    if (!h->profiling || h->ev_ring.empty()) return fail(BSM_ERR_ARG, "profiling is off");
    if (h->prof_count == 0) return fail(BSM_ERR_ARG, "no multiply has been recorded since profiling was switched on");
    DeviceGuard g(h->device);
    const int64_t n = std::min<int64_t>(h->prof_count, bsm_matrix::kProfSlots);
    double sa = 0.0, sb = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        cudaEvent_t *e = h->ev_ring.data() + 3 * ((h->prof_count - 1 - k) % bsm_matrix::kProfSlots);
        CUDA_TRY(cudaEventSynchronize(e[2]));
        float a = 0.f, b = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&a, e[0], e[1]));
        CUDA_TRY(cudaEventElapsedTime(&b, e[1], e[2]));
        sa += a;
        sb += b;
    }
    if (main_ms) *main_ms = sa / (double)n;
    if (finalize_ms) *finalize_ms = sb / (double)n;
    h->prof_count = 0;
    return 0;
}

int bsm_mul(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
            const void *x_dev, int64_t ldx, void *y_dev, int64_t ldy, int64_t nrhs, void *stream) {
    if (int rc = check_handle(h)) return rc;
    if (op < BSM_OP_N || op > BSM_OP_C) return fail(BSM_ERR_ARG, "bad op");
    if (!alpha || (!beta && !beta_is_false)) return fail(BSM_ERR_ARG, "alpha/beta is null");
    if (nrhs < 0) return fail(BSM_ERR_ARG, "negative nrhs");
    if (nrhs == 0) return 0;
    if (!x_dev || !y_dev) return fail(BSM_ERR_ARG, "x or y is null");
    if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle: no device, and there is no CPU fallback");
    const int64_t nout = (op == BSM_OP_N) ? h->H.nrows : h->H.ncols;
    const int64_t nin = (op == BSM_OP_N) ? h->H.ncols : h->H.nrows;
    if (nrhs > 1 && (ldx < nin || ldy < nout)) return fail(BSM_ERR_ARG, "leading dimension too small");
    DeviceGuard g(h->device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    switch (h->H.dtype) {
    case BSM_F32:
        return launch_mul<float>(h, op, alpha, beta, beta_is_false, (const float *)x_dev, ldx,
                                 (float *)y_dev, ldy, nrhs, st);
    case BSM_F64:
        return launch_mul<double>(h, op, alpha, beta, beta_is_false, (const double *)x_dev, ldx,
                                  (double *)y_dev, ldy, nrhs, st);
    default:
        return launch_mul<cplx>(h, op, alpha, beta, beta_is_false, (const cplx *)x_dev, ldx,
                                (cplx *)y_dev, ldy, nrhs, st);
    }
}

}  // extern "C"

// internal (dist.cu): one phase of a slab multiply, see launch_mul. Returns 1 through *has_remote when the plan
// of `op` holds remote slices at all (otherwise a plain bsm_mul after the gather is just as good).
int bsm_plan_has_remote(bsm_handle h, int op) {
    if (!h || op < BSM_OP_N || op > BSM_OP_C) return 0;
    const int p = plan_index(h, op);
    return (p < 4 && h->H.plan[p].has_remote) ? 1 : 0;
}
// ---- sparse.cu hooks
static_assert(sizeof(SparseResult) <= sizeof(void *) * 8, "sparse slot too small");
int bsm_sparse_source(bsm_handle h, SparseSource *out) {
    if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle: no device, and there is no CPU fallback");
    out->H = &h->H;
    out->arena = h->arena;
    out->set_start = h->set_start.p;
    out->set_pool_off = h->set_pool_off.p;
    out->pool = h->pool.p;
    out->device = h->device;
    out->restricted = h->restricted;
    return 0;
}
SparseResult *bsm_sparse_slot(bsm_handle h) {
    SparseResult *R = reinterpret_cast<SparseResult *>(h->sparse_slot);
    if (!h->sparse_slot[7]) {     // first use: construct in place (slot 7 doubles as the "initialised" mark)
        new (R) SparseResult();
        h->sparse_slot[7] = (void *)1;
    }
    return R;
}

int bsm_get_scratch(bsm_handle h, void *stream, void **out) {
    if (int rc = check_handle(h)) return rc;
    DeviceGuard g(h->device);
    return get_scratch(h, (cudaStream_t)stream, out);
}
int64_t bsm_plan_scratch_bytes(bsm_handle h, int op) {
    if (!h || op < BSM_OP_N || op > BSM_OP_C) return 0;
    return h->H.plan[plan_index(h, op)].scratch_elems * (int64_t)dtype_size(h->H.dtype);
}
int bsm_mul_phase(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false, const void *x_dev,
                  void *y_dev, void *stream, int phase, void **scratch_io, const PeerX *px) {
    if (int rc = check_handle(h)) return rc;
    if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle");
    if (op < BSM_OP_N || op > BSM_OP_C) return fail(BSM_ERR_ARG, "bad op");
    if (!alpha || (!beta && !beta_is_false) || !x_dev || !y_dev) return fail(BSM_ERR_ARG, "null argument");
    DeviceGuard g(h->device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    switch (h->H.dtype) {
    case BSM_F32:
        return launch_mul<float>(h, op, alpha, beta, beta_is_false, (const float *)x_dev, 0, (float *)y_dev, 0, 1, st,
                                 phase, scratch_io, px);
    case BSM_F64:
        return launch_mul<double>(h, op, alpha, beta, beta_is_false, (const double *)x_dev, 0, (double *)y_dev, 0, 1,
                                  st, phase, scratch_io, px);
    default:
        return launch_mul<cplx>(h, op, alpha, beta, beta_is_false, (const cplx *)x_dev, 0, (cplx *)y_dev, 0, 1, st,
                                phase, scratch_io, px);
    }
}

extern "C" {

int bsm_mul_host(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                 const void *x_host, int64_t ldx, void *y_host, int64_t ldy, int64_t nrhs) {
    if (int rc = check_handle(h)) return rc;
    if (op < BSM_OP_N || op > BSM_OP_C) return fail(BSM_ERR_ARG, "bad op");
    if (nrhs < 0) return fail(BSM_ERR_ARG, "negative nrhs");
    if (nrhs == 0) return 0;
    if (!x_host || !y_host) return fail(BSM_ERR_ARG, "x or y is null");
    if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle: no device, and there is no CPU fallback");
    const int64_t s = dtype_size(h->H.dtype);
    const int64_t nout = (op == BSM_OP_N) ? h->H.nrows : h->H.ncols;
    const int64_t nin = (op == BSM_OP_N) ? h->H.ncols : h->H.nrows;
    if (nrhs > 1 && (ldx < nin || ldy < nout)) return fail(BSM_ERR_ARG, "leading dimension too small");
    std::lock_guard<std::mutex> lock(h->host_mu);
    DeviceGuard g(h->device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    if (!h->host_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    const int64_t xb = nin * nrhs * s, yb = nout * nrhs * s;
    if (h->hx_bytes < xb) {
        if (h->hx) cudaFree(h->hx);
        h->hx = nullptr;
        h->hx_bytes = 0;
        CUDA_TRY(cudaMalloc(&h->hx, (size_t)xb));
        h->hx_bytes = xb;
    }
    if (h->hy_bytes < yb) {
        if (h->hy) cudaFree(h->hy);
        h->hy = nullptr;
        h->hy_bytes = 0;
        CUDA_TRY(cudaMalloc(&h->hy, (size_t)yb));
        h->hy_bytes = yb;
    }
    cudaStream_t st = h->host_stream;
    // slab handles (bsm_options.own_*): only the owned rows of y exist for this handle — they alone cross PCIe
    const int ob = (op == BSM_OP_N) ? 0 : 1;
    const int64_t olo = h->own_hi[ob] >= 0 ? std::max<int64_t>(0, h->own_lo[ob]) : 0;
    const int64_t ohi = h->own_hi[ob] >= 0 ? std::min<int64_t>(nout, h->own_hi[ob]) : nout;
    const size_t ypitch_h = (size_t)((nrhs > 1 ? ldy : nout) * s), ypitch_d = (size_t)(nout * s);
    const size_t ywidth = (size_t)(std::max<int64_t>(0, ohi - olo) * s);
    CUDA_TRY(cudaMemcpy2DAsync(h->hx, (size_t)(nin * s), x_host, (size_t)((nrhs > 1 ? ldx : nin) * s),
                               (size_t)(nin * s), (size_t)nrhs, cudaMemcpyHostToDevice, st));
    if (!beta_is_false && ywidth > 0)
        CUDA_TRY(cudaMemcpy2DAsync((char *)h->hy + olo * s, ypitch_d, (const char *)y_host + olo * s, ypitch_h, ywidth,
                                   (size_t)nrhs, cudaMemcpyHostToDevice, st));
    if (int rc = bsm_mul(h, op, alpha, beta, beta_is_false, h->hx, nin, h->hy, nout, nrhs, (void *)st)) return rc;
    if (ywidth > 0)
        CUDA_TRY(cudaMemcpy2DAsync((char *)y_host + olo * s, ypitch_h, (const char *)h->hy + olo * s, ypitch_d, ywidth,
                                   (size_t)nrhs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ---- queries -------------------------------------------------------------------------------
int64_t bsm_nnz(bsm_handle h) { return h ? h->H.nnz : BSM_ERR_ARG; }
int64_t bsm_stored_entries(bsm_handle h) { return h ? h->H.stored : BSM_ERR_ARG; }
int bsm_size(bsm_handle h, int64_t *nrows, int64_t *ncols) {
    if (int rc = check_handle(h)) return rc;
    if (nrows) *nrows = h->H.nrows;
    if (ncols) *ncols = h->H.ncols;
    return 0;
}
int bsm_dtype_of(bsm_handle h) { return h ? h->H.dtype : BSM_ERR_ARG; }
int bsm_kind_of(bsm_handle h) { return h ? h->H.kind : BSM_ERR_ARG; }

int bsm_work(bsm_handle h, int op, int64_t nrhs, int beta_used, double *bytes, double *flops,
             double *index_table_bytes) {
    if (int rc = check_handle(h)) return rc;
    if (op < BSM_OP_N || op > BSM_OP_C) return fail(BSM_ERR_ARG, "bad op");
    const HostMatrix &H = h->H;
    const HostPlan &P = H.plan[plan_index(h, op)];
    const double s = dtype_size(H.dtype);
    const double tab = (double)P.contrib.size() * sizeof(bsm_contrib) + (double)P.slices.size() * sizeof(bsm_slice) +
                       (double)P.wchunk.size() * sizeof(bsm_wchunk) +
                       (double)H.sets.pool.size() * 4 + (double)H.sets.len.size() * 16 +
                       (double)P.gather_rows.size() * 12 + (double)P.gather_pos.size() * 8;
    if (bytes)
        *bytes = (double)H.stored * s + ((double)P.in_dim + (double)P.out_dim * (beta_used ? 2.0 : 1.0)) * s * (double)nrhs + tab;
    if (flops) *flops = (H.dtype == BSM_C64 ? 8.0 : 2.0) * (double)P.applied_entries * (double)nrhs;
    if (index_table_bytes) *index_table_bytes = tab;
    return 0;
}

int bsm_plan_stats(bsm_handle h, int op, int64_t out[12]) {
    if (int rc = check_handle(h)) return rc;
    if (op < BSM_OP_N || op > BSM_OP_C || !out) return fail(BSM_ERR_ARG, "bad op or null output");
    const HostPlan &P = h->H.plan[plan_index(h, op)];
    const int64_t s = dtype_size(h->H.dtype);
    for (int i = 0; i < 12; ++i) out[i] = 0;
    for (size_t i = 0; i < P.slices.size(); ++i) {
        const bsm_slice &sl = P.slices[i];
        const int cls = (sl.flags & kSliceFused) ? 0 : (sl.flags & kSliceWarp) ? 1 : 2;
        out[cls]++;
        for (int32_t c = sl.c_begin; c < sl.c_end; ++c) {
            const bsm_contrib &cb = P.contrib[c];
            const int64_t lo = sl.r0, hi = std::min<int64_t>(sl.r1, cb.out_len);
            if (hi > lo) out[3 + cls] += (hi - lo) * ((cb.form & kFormT) ? cb.m : cb.n) * s;
        }
    }
    out[6] = P.witem_ptr.empty() ? 0 : (int64_t)P.witem_ptr.size() - 1;
    out[7] = (int64_t)P.wchunk.size();
    out[8] = P.scratch_elems;
    out[9] = (int64_t)P.gather_rows.size();
    out[10] = P.spmm_ok ? (int64_t)P.mitem_ptr.size() - 1 : 0;
    out[11] = (P.spmm_ok && P.spmm_tma) ? 1 : 0;
    return 0;
}

int bsm_launch_count(bsm_handle h, int op) {
    if (!h || op < BSM_OP_N || op > BSM_OP_C) return BSM_ERR_ARG;
    const HostPlan &P = h->H.plan[plan_index(h, op)];
    if (P.color_ok) return (int)P.color_ptr.size();      // scale kernel + one launch per (sweep, colour)
    const bool only_warp = P.n_fused_slices == 0 && (int64_t)P.slices.size() == P.n_warp_slices && P.n_warp_slices > 0;
    const bool folded = P.wcta && P.scratch_elems == 0 && only_warp;      // zero rows set inside the warp-stream launch
    return (P.n_fused_slices > 0 ? 1 : 0) + (P.n_warp_slices > 0 ? 1 : 0) +
           ((int64_t)P.slices.size() > P.n_fused_slices + P.n_warp_slices ? 1 : 0) + ((P.gather_rows.empty() || folded) ? 0 : 1);
}

}  // extern "C"

// ---- table export --------------------------------------------------------------------------
namespace {
struct TabView {
    const void *p = nullptr;
    int64_t count = -1;
    int64_t elem = 0;
    bool device_arena = false;
};
template <class U>
TabView view(const std::vector<U> &v) {
    TabView t;
    t.p = v.data();
    t.count = (int64_t)v.size();
    t.elem = (int64_t)sizeof(U);
    return t;
}
TabView table_view(bsm_handle h, int table, int plan, std::vector<int32_t> &tmp32) {
    const HostMatrix &H = h->H;
    TabView t;
    if (plan < 0 || plan > 5) return t;
    const HostPlan &P = H.plan[plan];
    switch (table) {
    case BSM_TAB_ARENA:
        t.count = H.arena_elems;
        t.elem = dtype_size(H.dtype);
        t.device_arena = true;
        return t;
    case BSM_TAB_BLOCK_OFF: return view(H.block_off);
    case BSM_TAB_BLOCK_M:
        tmp32.clear();
        for (auto &b : H.blocks) tmp32.push_back(b.m);
        return view(tmp32);
    case BSM_TAB_BLOCK_N:
        tmp32.clear();
        for (auto &b : H.blocks) tmp32.push_back(b.n);
        return view(tmp32);
    case BSM_TAB_SET_LEN: return view(H.sets.len);
    case BSM_TAB_SET_START: return view(H.sets.start);
    case BSM_TAB_SET_POOL_OFF: return view(H.sets.pool_off);
    case BSM_TAB_POOL: return view(H.sets.pool);
    case BSM_TAB_CONTRIB: return view(P.contrib);
    case BSM_TAB_SLICE: return view(P.slices);
    case BSM_TAB_GATHER_ROWS: return view(P.gather_rows);
    case BSM_TAB_GATHER_PTR: return view(P.gather_ptr);
    case BSM_TAB_GATHER_POS: return view(P.gather_pos);
    case BSM_TAB_GROUP_PTR: return view(P.group_ptr);
    case BSM_TAB_GROUP_SET: return view(P.group_set);
    case BSM_TAB_CONTRIB_TOFF: return view(P.contrib_toff);
    case BSM_TAB_WCHUNK: return view(P.wchunk);
    case BSM_TAB_WITEM_PTR: return view(P.witem_ptr);
    case BSM_TAB_COLOR_PTR: return view(P.color_ptr);
    }
    return t;
}
}  // namespace

extern "C" {

int64_t bsm_table_count(bsm_handle h, int table, int plan) {
    if (!h) return BSM_ERR_ARG;
    std::vector<int32_t> tmp;
    return table_view(h, table, plan, tmp).count;
}

int bsm_table_copy(bsm_handle h, int table, int plan, void *dst, int64_t dst_bytes) {
    if (int rc = check_handle(h)) return rc;
    std::vector<int32_t> tmp;
    TabView t = table_view(h, table, plan, tmp);
    if (t.count < 0) return fail(BSM_ERR_ARG, "unknown table");
    if (dst_bytes < t.count * t.elem) return fail(BSM_ERR_ARG, "destination too small");
    if (t.count == 0) return 0;
    if (!dst) return fail(BSM_ERR_ARG, "dst is null");
    if (t.device_arena) {
        if (h->device == BSM_DEVICE_NONE) return fail(BSM_ERR_CUDA, "host-only handle has no arena");
        DeviceGuard g(h->device);
        CUDA_TRY(cudaMemcpy(dst, h->arena, (size_t)(t.count * t.elem), cudaMemcpyDeviceToHost));
    } else {
        std::memcpy(dst, t.p, (size_t)(t.count * t.elem));
    }
    return 0;
}

// ---- device helpers ------------------------------------------------------------------------
int bsm_device_count(int *count) {
    if (!count) return fail(BSM_ERR_ARG, "count is null");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(BSM_ERR_CUDA, cudaGetErrorString(e));
    }
    return 0;
}
int bsm_malloc(int device, size_t bytes, void **dev_ptr) {
    if (!dev_ptr) return fail(BSM_ERR_ARG, "dev_ptr is null");
    DeviceGuard g(device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaMalloc(dev_ptr, bytes ? bytes : 16));
    return 0;
}
int bsm_free(int device, void *dev_ptr) {
    DeviceGuard g(device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaFree(dev_ptr));
    return 0;
}
int bsm_memcpy_h2d(void *dev_dst, const void *host_src, size_t bytes) {
    CUDA_TRY(cudaMemcpy(dev_dst, host_src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int bsm_memcpy_d2h(void *host_dst, const void *dev_src, size_t bytes) {
    CUDA_TRY(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int bsm_synchronize(int device) {
    DeviceGuard g(device);
    if (!g.ok) return fail(BSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
