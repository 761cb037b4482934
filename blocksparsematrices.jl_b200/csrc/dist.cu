// dist.cu — single-box multi-GPU layer of libbsm_b200: one process per GPU, block-row slabs, x replicated
// by an all-gather of the slab slices over NCCL (NVLink 5 / NVSwitch), every rank writes its own y slice.
//
// The reference is single-process (nothing to cite); this is the §8(e) design. NCCL is bound at run time
// with dlopen("libnccl.so.2") so the library carries no link-time dependency and, inside a PyTorch
// process, shares the NCCL build PyTorch already loaded. Slabs are uneven (nnz-balanced), so the
// all-gather is one NCCL group of in-place broadcasts — one per (rank, right-hand side column) — directly
// on the caller's column-major x: no packing, no staging copy.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bsm_b200.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
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
int bsm_mul_phase(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false, const void *x_dev,
                  void *y_dev, void *stream, int phase, void **scratch_io);

struct bsm_comm_s {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cudaStream_t comm_stream = nullptr;     // the all-gather runs here, beside the rank-local slices
    cudaEvent_t ev_ready = nullptr, ev_gathered = nullptr;
    int overlap = 1;
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
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
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
    // contiguous case (one column, or columns packed back to back over all rows): one broadcast per rank
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

int bsm_dist_allreduce_max_f64(bsm_comm c, double *dev_values, int64_t count, void *stream) {
    if (!c || !dev_values) return dfail(BSM_ERR_ARG, "null argument");
    if (int rc = api_ok()) return rc;
    NCCL_TRY(nccl().AllReduce(dev_values, dev_values, (size_t)count, ncclDouble, ncclMax, c->comm, (cudaStream_t)stream));
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
        if (cudaEventRecord(c->ev_ready, st) != cudaSuccess || cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0) != cudaSuccess)
            return dfail(BSM_ERR_CUDA, "event record/wait failed");
        if (int rc = bsm_dist_allgather_rows(c, bsm_dtype_of(h), x_dev, ldx, 1, in_cuts, (void *)c->comm_stream)) return rc;
        if (cudaEventRecord(c->ev_gathered, c->comm_stream) != cudaSuccess) return dfail(BSM_ERR_CUDA, "event record failed");
        void *scratch = nullptr;
        if (int rc = bsm_mul_phase(h, op, alpha, beta, beta_is_false, x_dev, y_dev, stream, 1, &scratch)) return rc;
        if (cudaStreamWaitEvent(st, c->ev_gathered, 0) != cudaSuccess) return dfail(BSM_ERR_CUDA, "event wait failed");
        return bsm_mul_phase(h, op, alpha, beta, beta_is_false, x_dev, y_dev, stream, 2, &scratch);
    }
    if (c->nranks > 1) {
        if (int rc = bsm_dist_allgather_rows(c, bsm_dtype_of(h), x_dev, ldx, nrhs, in_cuts, stream)) return rc;
    }
    return bsm_mul(h, op, alpha, beta, beta_is_false, x_dev, ldx, y_dev, ldy, nrhs, stream);
}

}  // extern "C"
