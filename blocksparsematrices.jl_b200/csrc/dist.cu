// dist.cu — single-box multi-GPU layer of libbsm_b200: one process per GPU, block-row slabs, x replicated
// by an all-gather of the slab slices over NCCL (NVLink 5 / NVSwitch), every rank writes its own y slice.
//
// The reference is single-process (nothing to cite); this is the §8(e) design. NCCL is bound at run time
// with dlopen("libnccl.so.2") so the library carries no link-time dependency and, inside a PyTorch
// process, shares the NCCL build PyTorch already loaded. Slabs are uneven (nnz-balanced): every rank packs
// its slab into one of nranks equal chunks of a staging buffer, ONE ncclAllGather, strided 2-D device copies
// back into x (a group of in-place broadcasts per rank and column is kept as the comparison).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bsm_b200.h"
#include "plan.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string err;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy PyTorch loaded, if any
            if (api.lib) break;
        }
        for (const char *n : names) {
            if (api.lib) break;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!api.lib) {
            api.err = std::string("libnccl.so.2 not found: ") + dlerror();
            return;
        }
#define BSM_SYM(field, name)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));   \
    if (!api.field) api.err = std::string("NCCL symbol missing: ") + name;
        BSM_SYM(GetUniqueId, "ncclGetUniqueId")
        BSM_SYM(CommInitRank, "ncclCommInitRank")
        BSM_SYM(CommDestroy, "ncclCommDestroy")
        BSM_SYM(GroupStart, "ncclGroupStart")
        BSM_SYM(GroupEnd, "ncclGroupEnd")
        BSM_SYM(Broadcast, "ncclBroadcast")
        BSM_SYM(AllGather, "ncclAllGather")
        BSM_SYM(AllReduce, "ncclAllReduce")
        BSM_SYM(GetErrorString, "ncclGetErrorString")
        BSM_SYM(GetVersion, "ncclGetVersion")
#undef BSM_SYM
    });
    return api;
}

thread_local std::string g_dist_err;

}  // namespace

// abi.cu owns the thread-local error string of bsm_last_error(); this hook lets dist.cu set it.
void bsm_set_error(const std::string &msg);

int bsm_plan_has_remote(bsm_handle h, int op);
int64_t bsm_plan_scratch_bytes(bsm_handle h, int op);
int bsm_get_scratch(bsm_handle h, void *stream, void **out);
int bsm_mul_phase(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false, const void *x_dev,
                  void *y_dev, void *stream, int phase, void **scratch_io, const bsm::PeerX *px);

struct bsm_comm_s {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cudaStream_t comm_stream = nullptr;     // the all-gather runs here, beside the rank-local slices
    cudaStream_t aux_stream = nullptr;      // the remote slices run here, beside the tail of the local ones
    cudaEvent_t ev_ready = nullptr, ev_gathered = nullptr, ev_remote = nullptr;
    int overlap = 1;
    int use_broadcasts = 0;                 // comparison: one grouped in-place broadcast per rank and column
    void *stage = nullptr;                  // nranks equal chunks of max-slab size for ncclAllGather
    size_t stage_bytes = 0;
    // peer mode: allocations mapped into every rank (CUDA IPC), and the flag words of the two barriers that
    // bracket a peer-mode multiply
    struct Shared {
        void *local = nullptr;
        void *peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    };
    std::vector<Shared> shared;
    int32_t *flags = nullptr;               // this rank's flag array: ready[nranks], done[nranks]
    int32_t **peer_flags_dev = nullptr;     // device array: every rank's flag array
    int32_t *sync_state = nullptr;          // local device words of the in-kernel barriers (kernels.cuh PeerSync)
    int debug = 0;                          // benchmarking only: bit0 no entry wait, bit1 no exit wait
};

namespace {
int dfail(int code, const std::string &msg) {
    bsm_set_error(msg);
    return code;
}
#define NCCL_TRY(expr)                                                                             \
    do {                                                                                           \
        ncclResult_t r__ = (expr);                                                                 \
        if (r__ != ncclSuccess) return dfail(BSM_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(r__)); \
    } while (0)
int api_ok() {
    if (!nccl().err.empty() || !nccl().lib) return dfail(BSM_ERR_UNSUPPORTED, nccl().err.empty() ? "NCCL unavailable" : nccl().err);
    return 0;
}
}  // namespace

extern "C" {

int bsm_dist_unique_id(void *id128) {
    if (!id128) return dfail(BSM_ERR_ARG, "id128 is null");
    if (int rc = api_ok()) return rc;
    static_assert(sizeof(ncclUniqueId) == BSM_DIST_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}

int bsm_dist_init(const void *id128, int nranks, int rank, int device, bsm_comm *out) {
    if (!out) return dfail(BSM_ERR_ARG, "out is null");
    *out = nullptr;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return dfail(BSM_ERR_ARG, "bad communicator arguments");
    if (int rc = api_ok()) return rc;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return dfail(BSM_ERR_CUDA, "no CUDA device");
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return dfail(BSM_ERR_CUDA, "cudaGetDevice failed");
    if (cudaSetDevice(device) != cudaSuccess) return dfail(BSM_ERR_CUDA, "cudaSetDevice failed");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    bsm_comm_s *c = new bsm_comm_s();
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return dfail(BSM_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    }
    // highest priority: the NCCL blocks must get SM slots as soon as compute blocks retire, not after the
    // whole compute grid has drained
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_remote, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_gathered, cudaEventDisableTiming) != cudaSuccess) {
        bsm_dist_destroy(c);
        return dfail(BSM_ERR_CUDA, "could not create the communication stream");
    }
    *out = c;
    return 0;
}

int bsm_dist_set_overlap(bsm_comm c, int on) {
    if (!c) return dfail(BSM_ERR_ARG, "null communicator");
    c->overlap = on ? 1 : 0;
    return 0;
}

int bsm_dist_destroy(bsm_comm c) {
    if (!c) return 0;
    for (auto &sh : c->shared) {
        for (int p = 0; p < c->nranks; ++p)
            if (p != c->rank && sh.peer[p]) cudaIpcCloseMemHandle(sh.peer[p]);
        if (sh.local) cudaFree(sh.local);
    }
    if (c->peer_flags_dev) cudaFree(c->peer_flags_dev);
    if (c->sync_state) cudaFree(c->sync_state);
    if (c->stage) cudaFree(c->stage);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_remote) cudaEventDestroy(c->ev_remote);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_gathered) cudaEventDestroy(c->ev_gathered);
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
    delete c;
    return 0;
}

int bsm_dist_info(bsm_comm c, int *nranks, int *rank, int *nccl_version) {
    if (!c) return dfail(BSM_ERR_ARG, "null communicator");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (nccl().GetVersion) nccl().GetVersion(nccl_version);
    }
    return 0;
}

int bsm_dist_allgather_rows(bsm_comm c, int dtype, void *x_dev, int64_t ldx, int64_t nrhs, const int64_t *cuts,
                            void *stream) {
    if (!c || !x_dev || !cuts) return dfail(BSM_ERR_ARG, "null argument");
    if (dtype < 0 || dtype > 2) return dfail(BSM_ERR_ARG, "bad dtype");
    if (nrhs < 1) return 0;
    if (int rc = api_ok()) return rc;
    const int64_t s = dtype == BSM_F32 ? 4 : dtype == BSM_F64 ? 8 : 16;
    for (int r = 0; r < c->nranks; ++r)
        if (cuts[r + 1] < cuts[r]) return dfail(BSM_ERR_ARG, "cuts must be non-decreasing");
    if (nrhs > 1 && ldx < cuts[c->nranks]) return dfail(BSM_ERR_ARG, "leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *base = (unsigned char *)x_dev;
    int64_t maxrows = 0;
    for (int r = 0; r < c->nranks; ++r) maxrows = std::max(maxrows, cuts[r + 1] - cuts[r]);
    if (maxrows == 0) return 0;
    if (!c->use_broadcasts) {
        // Slabs are uneven (nnz-balanced) but ncclAllGather wants equal counts and is by far the fastest
        // collective on NVSwitch (8 grouped broadcasts of 4 MB cost 0.4 ms on 8 GPUs, the all-gather 10x
        // less): every rank packs its slab into chunk `rank` of a staging buffer of nranks max-size chunks,
        // one in-place all-gather, then the other ranks' chunks are copied to their rows of x. The copies are
        // strided 2-D device copies (one per rank), so a column-major multi-RHS x needs no kernel either.
        const size_t chunk = (size_t)(maxrows * nrhs * s);
        const size_t need = chunk * (size_t)c->nranks;
        if (c->stage_bytes < need) {
            // only the staging buffer is replaced; earlier collectives may still read it on either stream
            if (c->stage) {
                cudaStreamSynchronize(st);
                cudaStreamSynchronize(c->comm_stream);
                cudaFree(c->stage);
            }
            c->stage = nullptr;
            c->stage_bytes = 0;
            if (cudaMalloc(&c->stage, need) != cudaSuccess) return dfail(BSM_ERR_ALLOC, "staging buffer of the all-gather");
            c->stage_bytes = need;
        }
        unsigned char *stage = (unsigned char *)c->stage;
        const size_t spitch = (size_t)(ldx * s), dpitch = (size_t)(maxrows * s);
        const int64_t myrows = cuts[c->rank + 1] - cuts[c->rank];
        if (myrows > 0 &&
            cudaMemcpy2DAsync(stage + chunk * c->rank, dpitch, base + cuts[c->rank] * s, nrhs > 1 ? spitch : dpitch,
                              (size_t)(myrows * s), (size_t)nrhs, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return dfail(BSM_ERR_CUDA, "packing the x slab failed");
        NCCL_TRY(nccl().AllGather(stage + chunk * c->rank, stage, chunk, ncclChar, c->comm, st));
        for (int r = 0; r < c->nranks; ++r) {
            const int64_t rows = cuts[r + 1] - cuts[r];
            if (r == c->rank || rows == 0) continue;
            if (cudaMemcpy2DAsync(base + cuts[r] * s, nrhs > 1 ? spitch : dpitch, stage + chunk * r, dpitch,
                                  (size_t)(rows * s), (size_t)nrhs, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                return dfail(BSM_ERR_CUDA, "unpacking the gathered x failed");
        }
        return 0;
    }
    NCCL_TRY(nccl().GroupStart());
    for (int r = 0; r < c->nranks; ++r) {
        const int64_t rows = cuts[r + 1] - cuts[r];
        if (rows == 0) continue;
        for (int64_t j = 0; j < nrhs; ++j) {
            unsigned char *p = base + (j * ldx + cuts[r]) * s;
            ncclResult_t rr = nccl().Broadcast(p, p, (size_t)(rows * s), ncclChar, r, c->comm, st);
            if (rr != ncclSuccess) {
                nccl().GroupEnd();
                return dfail(BSM_ERR_CUDA, std::string("ncclBroadcast: ") + nccl().GetErrorString(rr));
            }
        }
    }
    NCCL_TRY(nccl().GroupEnd());
    return 0;
}

int bsm_dist_set_debug(bsm_comm c, int flags) {
    if (!c) return dfail(BSM_ERR_ARG, "null communicator");
    c->debug = flags;
    return 0;
}

int bsm_dist_debug_read(bsm_comm c, int64_t out[8]) {
    if (!c || !out) return dfail(BSM_ERR_ARG, "null argument");
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (!c->sync_state) return 0;
    if (cudaMemcpy(out, c->sync_state + 8, sizeof(long long) * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemset(c->sync_state + 8, 0, sizeof(long long) * 8) != cudaSuccess)
        return dfail(BSM_ERR_CUDA, "reading the barrier timers failed");
    return 0;
}

int bsm_dist_set_collective(bsm_comm c, int use_broadcasts) {
    if (!c) return dfail(BSM_ERR_ARG, "null communicator");
    c->use_broadcasts = use_broadcasts ? 1 : 0;
    return 0;
}

// ---- peer mode: x is read straight from its owners over NVLink -------------------------------------------
namespace {

// collective: cudaMalloc on every rank, IPC handles exchanged through the communicator, peers mapped
int shared_alloc(bsm_comm c, size_t bytes, bsm_comm_s::Shared *out) {
    if (c->nranks > 8) return dfail(BSM_ERR_UNSUPPORTED, "peer mode supports up to 8 ranks");
    if (cudaSetDevice(c->device) != cudaSuccess) return dfail(BSM_ERR_CUDA, "cudaSetDevice failed");
    bsm_comm_s::Shared sh;
    if (cudaMalloc(&sh.local, bytes ? bytes : 16) != cudaSuccess) return dfail(BSM_ERR_ALLOC, "cudaMalloc of a peer-mapped array failed");
    // the zero fill must have landed before any peer can learn the handle (a peer's flag store must not be overwritten)
    if (cudaMemset(sh.local, 0, bytes ? bytes : 16) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
        return dfail(BSM_ERR_CUDA, "cudaMemset failed");
    sh.peer[c->rank] = sh.local;
    if (c->nranks > 1) {
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, sh.local) != cudaSuccess) return dfail(BSM_ERR_CUDA, "cudaIpcGetMemHandle failed");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        unsigned char *dbuf = nullptr;
        if (cudaMalloc(&dbuf, 64 * (size_t)c->nranks) != cudaSuccess) return dfail(BSM_ERR_ALLOC, "cudaMalloc failed");
        cudaMemcpy(dbuf + 64 * c->rank, &mine, 64, cudaMemcpyHostToDevice);
        ncclResult_t r = nccl().AllGather(dbuf + 64 * c->rank, dbuf, 64, ncclChar, c->comm, c->comm_stream);
        if (r != ncclSuccess || cudaStreamSynchronize(c->comm_stream) != cudaSuccess) {
            cudaFree(dbuf);
            return dfail(BSM_ERR_CUDA, "exchange of the IPC handles failed");
        }
        std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
        cudaMemcpy(all.data(), dbuf, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost);
        cudaFree(dbuf);
        for (int p = 0; p < c->nranks; ++p) {
            if (p == c->rank) continue;
            cudaError_t e = cudaIpcOpenMemHandle(&sh.peer[p], all[(size_t)p], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return dfail(BSM_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer access between the GPUs of this box is required): ") +
                                               cudaGetErrorString(e));
        }
    }
    *out = sh;
    return 0;
}

int ensure_flags(bsm_comm c) {
    if (c->flags) return 0;
    bsm_comm_s::Shared sh;
    if (int rc = shared_alloc(c, sizeof(int32_t) * 2 * (size_t)c->nranks, &sh)) return rc;
    c->shared.push_back(sh);
    if (cudaMalloc(&c->peer_flags_dev, sizeof(int32_t *) * 8) != cudaSuccess ||
        cudaMalloc(&c->sync_state, sizeof(int32_t) * 8 + sizeof(long long) * 8) != cudaSuccess)
        return dfail(BSM_ERR_ALLOC, "cudaMalloc failed");
    if (cudaMemcpy(c->peer_flags_dev, sh.peer, sizeof(void *) * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(c->sync_state, 0, sizeof(int32_t) * 8 + sizeof(long long) * 8) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess)
        return dfail(BSM_ERR_CUDA, "initialising the barrier state failed");
    c->flags = (int32_t *)sh.local;   // set last: a failed setup is retried, never half-used
    return 0;
}

}  // namespace

int bsm_dist_alloc(bsm_comm c, size_t bytes, void **dev_ptr) {
    if (!c || !dev_ptr) return dfail(BSM_ERR_ARG, "null argument");
    if (int rc = api_ok()) return rc;
    bsm_comm_s::Shared sh;
    if (int rc = shared_alloc(c, bytes, &sh)) return rc;
    c->shared.push_back(sh);
    *dev_ptr = sh.local;
    return 0;
}

int bsm_dist_free(bsm_comm c, void *dev_ptr) {
    if (!c) return dfail(BSM_ERR_ARG, "null communicator");
    for (size_t i = 0; i < c->shared.size(); ++i) {
        if (c->shared[i].local != dev_ptr) continue;
        cudaDeviceSynchronize();
        for (int p = 0; p < c->nranks; ++p)
            if (p != c->rank && c->shared[i].peer[p]) cudaIpcCloseMemHandle(c->shared[i].peer[p]);
        cudaFree(c->shared[i].local);
        c->shared.erase(c->shared.begin() + (long)i);
        return 0;
    }
    return dfail(BSM_ERR_ARG, "not an array of bsm_dist_alloc");
}

int bsm_mul_dist_peer(bsm_comm c, bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                      void *x_shared, void *y_dev, const int64_t *in_cuts, void *stream) {
    if (!c || !h || !x_shared || !y_dev || !in_cuts) return dfail(BSM_ERR_ARG, "null argument");
    if (op < BSM_OP_N || op > BSM_OP_C) return dfail(BSM_ERR_ARG, "bad op");
    if (!alpha || (!beta && !beta_is_false)) return dfail(BSM_ERR_ARG, "alpha/beta is null");
    if (c->nranks == 1) return bsm_mul(h, op, alpha, beta, beta_is_false, x_shared, 0, y_dev, 0, 1, stream);
    for (int p = 0; p < c->nranks; ++p)
        if (in_cuts[p + 1] < in_cuts[p] || in_cuts[p + 1] >= (1ll << 31)) return dfail(BSM_ERR_ARG, "bad cuts");
    if (int rc = ensure_flags(c)) return rc;   // may grow c->shared: look x up afterwards
    const bsm_comm_s::Shared *sh = nullptr;
    for (const auto &s : c->shared)
        if (s.local == x_shared) sh = &s;
    if (!sh) return dfail(BSM_ERR_ARG, "x must be an array of bsm_dist_alloc");
    bsm::PeerX px;
    px.npeer = c->nranks;
    for (int p = 0; p < c->nranks; ++p) {
        px.peer[p] = sh->peer[p];
        px.cuts[p] = (int32_t)in_cuts[p];
    }
    px.cuts[c->nranks] = (int32_t)in_cuts[c->nranks];
    // The two barriers of a peer-mode multiply live INSIDE the multiply kernels (kernels.cuh PeerSync): "every x slab
    // of this epoch is written" is signalled by the first CTA to start and awaited before the first x fetch; "every
    // rank has finished reading" is signalled and awaited by the last CTA to finish — no extra launches. Nothing that
    // can fail is left between the argument checks above and the launches, so a rank either enters the collective
    // multiply or reports an error before any peer could wait on it.
    px.peer_flags = c->peer_flags_dev;
    px.my_flags = c->flags;
    px.state = c->sync_state;
    px.rank = c->rank;
    px.debug = c->debug;
    px.dbg = reinterpret_cast<long long *>(c->sync_state + 8);
    void *scratch = nullptr;
    return bsm_mul_phase(h, op, alpha, beta, beta_is_false, x_shared, y_dev, stream, 0, &scratch, &px);
}

int bsm_mul_dist_peer_host(bsm_comm c, bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                           const void *x_host_slab, void *x_shared, void *y_dev, void *y_host_slab,
                           const int64_t *in_cuts, int64_t out_lo, int64_t out_hi, void *stream) {
    if (!c || !h || !x_host_slab || !x_shared || !y_dev || !y_host_slab || !in_cuts)
        return dfail(BSM_ERR_ARG, "null argument");
    if (out_lo < 0 || out_hi < out_lo) return dfail(BSM_ERR_ARG, "bad output range");
    const int dt = bsm_dtype_of(h);
    if (dt < 0) return dfail(BSM_ERR_ARG, "bad handle");
    const int64_t s = dt == BSM_F32 ? 4 : dt == BSM_F64 ? 8 : 16;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaSetDevice(c->device) != cudaSuccess) return dfail(BSM_ERR_CUDA, "cudaSetDevice failed");
    const int64_t lo = in_cuts[c->rank], hi = in_cuts[c->rank + 1];
    // a failed copy must not keep this rank out of the collective multiply (the peers would wait for it): the
    // error is remembered, the multiply still runs, and the failure is reported afterwards
    cudaError_t e = cudaSuccess;
    if (hi > lo)
        e = cudaMemcpyAsync((unsigned char *)x_shared + lo * s, x_host_slab, (size_t)((hi - lo) * s), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && !beta_is_false && out_hi > out_lo)
        e = cudaMemcpyAsync((unsigned char *)y_dev + out_lo * s, y_host_slab, (size_t)((out_hi - out_lo) * s),
                            cudaMemcpyHostToDevice, st);
    const int rc = bsm_mul_dist_peer(c, h, op, alpha, beta, beta_is_false, x_shared, y_dev, in_cuts, stream);
    if (rc == 0 && e == cudaSuccess && out_hi > out_lo)
        e = cudaMemcpyAsync(y_host_slab, (unsigned char *)y_dev + out_lo * s, (size_t)((out_hi - out_lo) * s),
                            cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    if (rc != 0) return rc;
    if (e != cudaSuccess || es != cudaSuccess)
        return dfail(BSM_ERR_CUDA, std::string("host copies of the slab multiply: ") + cudaGetErrorString(e != cudaSuccess ? e : es));
    return 0;
}

int bsm_dist_allreduce_max_f64(bsm_comm c, double *dev_values, int64_t count, void *stream) {
    if (!c || !dev_values) return dfail(BSM_ERR_ARG, "null argument");
    if (int rc = api_ok()) return rc;
    NCCL_TRY(nccl().AllReduce(dev_values, dev_values, (size_t)count, ncclDouble, ncclMax, c->comm, (cudaStream_t)stream));
    return 0;
}

int bsm_dist_allreduce_sum_f64(bsm_comm c, double *dev_values, int64_t count, void *stream) {
    if (!c || !dev_values) return dfail(BSM_ERR_ARG, "null argument");
    if (c->nranks == 1) return 0;
    if (int rc = api_ok()) return rc;
    NCCL_TRY(nccl().AllReduce(dev_values, dev_values, (size_t)count, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream));
    return 0;
}

int bsm_mul_dist(bsm_comm c, bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                 void *x_dev, int64_t ldx, void *y_dev, int64_t ldy, int64_t nrhs, const int64_t *in_cuts,
                 void *stream) {
    if (!c || !h) return dfail(BSM_ERR_ARG, "null communicator or handle");
    if (c->nranks > 1 && c->overlap && nrhs == 1 && bsm_plan_has_remote(h, op)) {
        // overlap: the all-gather runs on the communicator's stream while the slices fed by this rank's own
        // x slab run on the caller's stream; the remote slices and the gather pass follow the all-gather
        cudaStream_t st = (cudaStream_t)stream;
        void *scratch = nullptr;
        if (int rc = bsm_get_scratch(h, stream, &scratch)) return rc;
        // x slab and scratch are ready at ev_ready: the all-gather (comm stream) and the remote slices (aux
        // stream, after the gather) hang off it, the local slices stay on the caller's stream
        if (cudaEventRecord(c->ev_ready, st) != cudaSuccess ||
            cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0) != cudaSuccess ||
            cudaStreamWaitEvent(c->aux_stream, c->ev_ready, 0) != cudaSuccess)
            return dfail(BSM_ERR_CUDA, "event record/wait failed");
        if (int rc = bsm_dist_allgather_rows(c, bsm_dtype_of(h), x_dev, ldx, 1, in_cuts, (void *)c->comm_stream)) return rc;
        if (cudaEventRecord(c->ev_gathered, c->comm_stream) != cudaSuccess) return dfail(BSM_ERR_CUDA, "event record failed");
        if (int rc = bsm_mul_phase(h, op, alpha, beta, beta_is_false, x_dev, y_dev, stream, 1, &scratch, nullptr)) return rc;
        if (cudaStreamWaitEvent(c->aux_stream, c->ev_gathered, 0) != cudaSuccess) return dfail(BSM_ERR_CUDA, "event wait failed");
        if (int rc = bsm_mul_phase(h, op, alpha, beta, beta_is_false, x_dev, y_dev, (void *)c->aux_stream, 2, &scratch, nullptr)) return rc;
        if (cudaEventRecord(c->ev_remote, c->aux_stream) != cudaSuccess ||
            cudaStreamWaitEvent(st, c->ev_remote, 0) != cudaSuccess ||
            cudaStreamWaitEvent(st, c->ev_gathered, 0) != cudaSuccess)
            return dfail(BSM_ERR_CUDA, "event record/wait failed");
        if (int rc = bsm_mul_phase(h, op, alpha, beta, beta_is_false, x_dev, y_dev, stream, 3, &scratch, nullptr)) return rc;
        return 0;
    }
    if (c->nranks > 1) {
        if (int rc = bsm_dist_allgather_rows(c, bsm_dtype_of(h), x_dev, ldx, nrhs, in_cuts, stream)) return rc;
    }
    return bsm_mul(h, op, alpha, beta, beta_is_false, x_dev, ldx, y_dev, ldy, nrhs, stream);
}

}  // extern "C"

// krylov.cu hook
// internal (krylov.cu): sets the flags and returns the previous ones
int bsm_dist_swap_debug_internal(bsm_comm c, int flags) {
    if (!c) return 0;
    const int old = c->debug;
    c->debug = flags;
    return old;
}

int bsm_dist_allreduce_sum_f64_internal(bsm_comm c, double *dev_values, int64_t count, void *stream) {
    return bsm_dist_allreduce_sum_f64(c, dev_values, count, stream);
}
