// plan.h — host-side intermediate representation and device plan of libbsm_b200.
//
// The three reference storage types (/root/reference/src/blockmatrix.jl:26-34,
// src/symmetricblockmatrix.jl:33-44, src/vbcrs.jl:36-43) are lowered to ONE device layout:
//
//   arena      every dense block, column-major as Julia stores it, verbatim, each block starting at
//              a 128-byte aligned offset of a single HBM allocation
//   sets       deduplicated index vectors ("index sets"): a contiguous range is (start, len), an
//              arbitrary vector is a slice of one Int32 pool (0-based)
//   plan[2]    plan 0 serves op N, plan 1 serves op T and op C (conj applied on the fly).
//              A plan is a list of *contributions*  y[out] += op(B) x[in]  (N-form: outputs along the
//              block's rows; T-form: outputs along its columns), grouped by output segment (= the
//              block-row pointer for plan 0, the transposed index for plan 1), cut into work items
//              ("slices"). A slice either owns its output rows (writes y directly, alpha/beta fused)
//              or writes a partial vector to scratch that a gather pass reduces in a fixed order —
//              no atomics anywhere, results are run-to-run deterministic.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bsm_b200.h"

namespace bsm {

static_assert(sizeof(bsm_contrib) == 32, "bsm_contrib must be 32 bytes");
static_assert(sizeof(bsm_slice) == 32, "bsm_slice must be 32 bytes");
static_assert(sizeof(bsm_wchunk) == 32, "bsm_wchunk must be 32 bytes");

constexpr int kSliceDirect = 1;
constexpr int kSliceVecOk = 2;
constexpr int kSliceFused = 4;          // handled by sym_fused_kernel (whole segment, <= kFusedMaxRows rows)
constexpr int kSliceRemote = 16;        // some input of the slice lies outside the rank's own x range (slab handles):
                                        // it must wait for the all-gather, the other slices overlap with it
constexpr int kSliceWarp = 8;           // handled by stream_warp_kernel (whole segment, <= kWarpMaxRows rows)
constexpr int kFusedMaxRows = 256;
constexpr int kFusedMaxTRows = 1024;    // tallest T-form block the CTA kernel stages (x window in shared memory)
constexpr int kWarpMaxRows = 64;        // segment length and block height limit of the warp-stream kernel
constexpr int kWChunkBytes = 4096;      // round-1 payload limit of a warp-stream chunk (plan_hints bit 2 / BSM_TUNE_WCHUNK)
constexpr int kWRingBytes = 11264;      // shared-memory byte ring of one warp (chunks + their x values)
constexpr int kWSlots = 16;             // mbarrier slots of one warp: chunks in flight + the one being consumed
constexpr int kWMaxCols = 64;           // columns per warp-stream chunk (two prefetched x values per lane)
// bsm_wchunk.flags
constexpr int kWcT = 1, kWcXPool = 2, kWcOutPool = 4, kWcSegBegin = 8, kWcSegEnd = 16, kWcDirect = 32;
constexpr int kWcCtaPart = 64;          // the item is one of the kWItemsPerCta parts of a segment: at its end the warp
                                        // hands its partial vector to the CTA, warp 0 sums the parts in order and writes
constexpr int kWItemsPerCta = 4;        // == kernels.cuh kWWarps
constexpr int kWCtaMaxSegments = 592;   // CTA-part mode only when every segment's CTA is resident at once (148 SMs x 4)
constexpr int kFormT = 1;               // bsm_contrib.form bit0: T-form
constexpr int kFormFusedT = 2;          // bit1: also emits the transposed partial of the same block
constexpr int kMaxSliceHeight = 128;   // outputs per work item (one CTA of 128 threads)
constexpr int64_t kArenaAlignBytes = 128;

inline int dtype_size(int dtype) { return dtype == BSM_F32 ? 4 : dtype == BSM_F64 ? 8 : 16; }
// elements per 16-byte vector load
inline int dtype_vec(int dtype) { return 16 / dtype_size(dtype); }

// ---- index sets ------------------------------------------------------------------------------
struct IndexSets {
    std::vector<int32_t> len, start;
    std::vector<int64_t> pool_off;
    std::vector<int32_t> pool;
    std::unordered_map<uint64_t, std::vector<int32_t>> by_hash;

    // 1-based Int64 vector as Julia holds it; every value must be in [1, limit]. Returns the set id
    // (existing id if an identical vector was added before) or -1 on a bad index.
    int32_t add_vector(const int64_t *idx1, int64_t n, int64_t limit);
    // 0-based contiguous range
    int32_t add_range(int64_t start0, int64_t n);
    // positions [k0, k0+n) of an existing set as a set of its own (shares the pool entries)
    int32_t add_subset(int32_t set, int64_t k0, int64_t n);
    inline int64_t at(int32_t s, int64_t k) const {
        return start[s] >= 0 ? (int64_t)start[s] + k : (int64_t)pool[pool_off[s] + k];
    }
};

// ---- blocks ----------------------------------------------------------------------------------
struct BlockSrc {
    const void *host;   // host pointer (read during create only)
    int32_t m, n;       // logical size of the block as it will sit in the arena
    bool transposed;    // host holds the n x m parent of a lazy transpose wrapper
};

// ---- plan ------------------------------------------------------------------------------------
struct ContribIR {
    int32_t block;
    int32_t form;      // 0 N-form, 1 T-form
    int32_t out_set;   // output segment (group key)
    int32_t in_set;
    int32_t out_len;
    int32_t sweep = 0;       // sweep the contribution belongs to, numbered in IR order (symmetric: 0 diagonals,
                             // 1 forward off-diagonals, 2 transposed off-diagonals — the reference runs the same
                             // three sweeps, each with its own colouring); only the colour-ordered variant uses it
    int32_t fuse_tset = -1;  // >= 0: one pass over the block also yields y[fuse_tset] += op(B)^T x[out_set]
                             // (half-stored symmetric off-diagonal block); delivered through scratch
};

struct HostPlan {
    std::vector<bsm_contrib> contrib;   // grouped by output segment
    std::vector<int64_t> group_ptr;     // CSR over contrib
    std::vector<int32_t> group_set;
    std::vector<uint8_t> group_direct;
    std::vector<int64_t> contrib_toff;  // scratch offset of the fused transposed partial, -1 if none
    std::vector<bsm_slice> slices;      // fused slices first, each class sorted by decreasing work
    int64_t n_fused_slices = 0;
    // slab handles: within every kernel class the slices whose inputs are all rank-local come first
    int64_t n_fused_local = 0, n_warp_items_local = 0, n_gather_local = 0;
    bool has_remote = false;
    bool fused_general = false;         // some CTA-kernel slice is a column sub-range or holds a tall T-form block
    int64_t n_warp_slices = 0;          // follow the fused slices; the rest go to gather_gemv_kernel
    std::vector<bsm_wchunk> wchunk;     // chunk stream of the warp slices, in slice order
    std::vector<int32_t> witem_ptr;     // warp work items: chunk ranges cut at segment boundaries
    int wform = 2;                      // forms of the warp-stream chunks: 0 N-form only, 1 T-form only, 2 both
    bool wcta = false;                  // small problems: every segment is ONE CTA of stream_warp_kernel, its chunk list
                                        // dealt to the kWItemsPerCta warps (items 4c .. 4c+3 = the parts of segment c,
                                        // possibly empty); single launch, no partial sums through global memory
    // multi-RHS (SpMM) path: usable when every slice is a direct warp-class slice (short segments that
    // own their rows); CTA work items = ranges of those slices, balanced by bytes
    // colour-ordered variant (plans 4/5): slices sorted by (sweep, colour); launch l runs slices
    // [color_ptr[l], color_ptr[l+1]) and accumulates straight into y (no two slices of a launch share a row)
    std::vector<int32_t> color_ptr;
    bool color_ok = false;
    bool spmm_ok = false;
    std::vector<bsm_slice> mslices;     // SpMM: one (direct, whole-segment) slice per block row, never split
    std::vector<int32_t> muncovered;    // SpMM: owned rows no block touches (y <- beta*y there)
    bool spmm_small = false;            // no block has more than 32 rows or columns (4-stage ring)
    bool spmm_tma = false;              // eligible for spmm_tma_kernel: segments and blocks of <= 32 rows, T-form blocks of
                                        // <= 32 columns, every input index set a contiguous range (X tiles are TMA boxes)
    std::vector<int32_t> mitem_ptr;
    std::vector<int32_t> gather_rows;
    std::vector<int64_t> gather_ptr;
    std::vector<int64_t> gather_pos;
    int64_t scratch_elems = 0;
    int64_t out_dim = 0, in_dim = 0;
    int64_t applied_entries = 0;        // Σ m*n over contributions
};

struct HostMatrix {
    int dtype = BSM_F64;
    int kind = BSM_KIND_BLOCKSPARSE;
    int64_t nrows = 0, ncols = 0;
    int64_t nnz = 0;             // SparseArrays.nnz semantics
    int64_t stored = 0;          // entries in the arena
    std::vector<BlockSrc> blocks;
    std::vector<int64_t> block_off;    // element offsets, 128-byte aligned
    int64_t arena_elems = 0;           // including tail slack
    IndexSets sets;
    // plan[0], plan[1]: GATHER variant for op N, op T/C. plan[2], plan[3]: stream (TMA-staged) variant.
    // plan[4], plan[5]: colour-ordered variant (the reference's schedule, for comparison).
    HostPlan plan[6];
    bool has_fused = false;
};

struct PlanParams {
    int64_t own_lo = 0, own_hi = -1;   // owned output range, hi < 0: everything
    int64_t in_lo = 0, in_hi = -1;     // range of x this rank owns before the all-gather, hi < 0: everything
    int64_t work_target_bytes = 512 << 10;
    bool fused = false;                // stream plan: segments of <= kFusedMaxRows rows go to the TMA-staged
                                       // kernels (CTA kernel, or warp-stream kernel when <= kWarpMaxRows), unsplit
    bool warp_stream = true;           // short segments may go to the warp-stream kernel (off for symmetric
                                       // matrices: their leaf segments stay with the CTA kernel, one launch)
    int64_t split_bytes = 0;           // stream plans: a segment of the CTA kernel streaming more than ~1.5x this is
                                       // cut into several work items (wide blocks by column ranges, partial sums
                                       // through the gather lists) so that no single CTA sets the makespan; 0: off
    int64_t wsplit_bytes = 0;          // small (L2-resident) problems: a warp-stream segment streaming more than ~1.5x
                                       // this is cut along its block list into several work items (partial sums
                                       // through the gather lists) to expose enough parallelism; 0: off
    int64_t witem_bytes = 0;           // target bytes per warp work item (0: derived from the total)
    int64_t witems_per_slot = 0;       // warp work items per resident warp slot (0: automatic)
    int64_t wchunk_bytes = 0;          // largest warp-stream chunk payload; 0: as much as lets two chunks share the ring
    bool wcta = true;                  // small problems with few segments may use the CTA-part mode (HostPlan::wcta)
};

// Lays the blocks out in the arena (fills block_off, arena_elems, stored).
void layout_arena(HostMatrix &M);
// Groups contributions, decides direct vs scratch ownership, cuts slices, builds the gather lists.
// Returns an empty string on success, else an error message.
// (adds index sets for column sub-ranges of split blocks: call before the tables are uploaded)
std::string build_plan(HostMatrix &M, const std::vector<ContribIR> &ir, int64_t out_dim,
                       int64_t in_dim, const PlanParams &pp, HostPlan &P);

// ---- dist.cu -> abi.cu: x read straight from its owners' peer-mapped arrays (bsm_mul_dist_peer) -----------
struct PeerX {
    const void *peer[8];
    int32_t cuts[9];
    int32_t npeer = 0;
    // barriers folded into the kernels (kernels.cuh PeerSync): every rank's flag array, this rank's own,
    // and the three local state words
    int32_t *const *peer_flags = nullptr;
    const int32_t *my_flags = nullptr;
    int32_t *state = nullptr;
    int32_t rank = 0;
    int32_t debug = 0;
    long long *dbg = nullptr;
};

// ---- sparse.cu <-> abi.cu (the handle's internals stay in abi.cu) --------------------------------------
struct SparseSource {
    const HostMatrix *H;
    const void *arena;
    const int32_t *set_start;
    const int64_t *set_pool_off;
    const int32_t *pool;
    int device;
    bool restricted;
};
struct SparseResult {     // device arrays of the last bsm_sparse_build: int64 colptr / rowval (1-based), T nzval
    int64_t nnz = -1, ncols = 0;
    int dtype = 0;
    void *colptr = nullptr, *rowval = nullptr, *nzval = nullptr;
};

// Colour-ordered plan: the reference's schedule (/root/reference/src/coloring.jl:20-61 builds the conflict
// graph "two blocks share an output row" and colours it; src/blockmatrix.jl:231-244 runs colour by
// colour). Greedy first-fit colouring per sweep stands in for GraphsColoring (it only changes the grouping,
// never the result). Leaves P.color_ok = false when the plan cannot be built (slab-restricted handle,
// repeated indices inside one block's output vector, more than 64 colours).
std::string build_color_plan(const HostMatrix &M, const std::vector<ContribIR> &ir, int64_t out_dim,
                             int64_t in_dim, const PlanParams &pp, HostPlan &P);

}  // namespace bsm
