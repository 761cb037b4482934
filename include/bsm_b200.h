/* bsm_b200.h — C ABI of libbsm_b200.so: the B200-native replacement for the multiply hot path of
 * BlockSparseMatrices.jl (v0.3.1). Plain pointers and sizes only; no CUDA/torch types.
 *
 * What each entry point replaces (file:line relative to the reference repository):
 *
 *   bsm_create_blocksparse   BlockSparseMatrix struct + ctor          src/blockmatrix.jl:26-34, 81-109
 *   bsm_create_symmetric     SymmetricBlockMatrix struct + ctor       src/symmetricblockmatrix.jl:33-44, 94-126
 *   bsm_create_vbcrs         VariableBlockCompressedRowStorage struct src/vbcrs.jl:36-43 (already sorted,
 *                            i.e. the output of the sorting ctor src/vbcrs.jl:78-122)
 *   bsm_mul                  LinearMaps._unsafe_mul!(y, A|A'|transpose(A), x, α, β) for the three types
 *                            src/abstractblockmatrix.jl:27-34, src/blockmatrix.jl:225-247,
 *                            src/symmetricblockmatrix.jl:386-435, src/vbcrs.jl:266-288, 303-354
 *   bsm_mul_host             the same call with HOST x / y (copies inside) — what `mul!(y, A, x)` on
 *                            Julia Arrays maps to when the caller holds no device arrays
 *   bsm_update_values        the same constructors when only the block values changed (no re-planning)
 *   bsm_nnz                  SparseArrays.nnz                         src/blockmatrix.jl:208-223,
 *                            src/symmetricblockmatrix.jl:367-384, src/vbcrs.jl:290-296
 *   bsm_size                 Base.size                                src/abstractblockmatrix.jl:23-25
 *
 * Conventions
 *   - Indices are 1-based Int64 at this boundary, exactly as Julia holds them; the packer converts
 *     once to 0-based Int32 device tables.
 *   - Blocks are column-major (Julia `Matrix{T}`), leading dimension = number of rows.
 *   - Host block / index pointers are read only during bsm_create_*; the handle owns device copies.
 *   - Every function returns 0 on success or a negative bsm_status; it never throws or exits.
 *     bsm_last_error() returns a thread-local message for the last failure.
 *   - A handle is immutable after creation: bsm_mul may be called concurrently from several host
 *     threads on different streams.
 *   - alpha / beta are passed by pointer as ONE element of the matrix dtype.
 *   - beta_is_false != 0 reproduces Julia's `β = false` strong zero (y is overwritten, NaN/Inf in y
 *     are not propagated: src/abstractblockmatrix.jl:33, src/blockmatrix.jl:231); beta is then ignored.
 *   - There is no CPU fallback: without a CUDA device every create/mul call fails with
 *     BSM_ERR_CUDA.
 */
#ifndef BSM_B200_H
#define BSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsm_matrix *bsm_handle;

typedef enum { BSM_F32 = 0, BSM_F64 = 1, BSM_C64 = 2 /* ComplexF64 */ } bsm_dtype;
typedef enum { BSM_OP_N = 0, BSM_OP_T = 1, BSM_OP_C = 2 } bsm_op; /* A, transpose(A), A' */
typedef enum { BSM_KIND_BLOCKSPARSE = 0, BSM_KIND_SYMMETRIC = 1, BSM_KIND_VBCRS = 2 } bsm_kind;

typedef enum {
    BSM_OK = 0,
    BSM_ERR_ARG = -1,     /* bad argument (null pointer, bad dtype/op, index out of range) */
    BSM_ERR_CUDA = -2,    /* CUDA runtime failure or no device */
    BSM_ERR_ALLOC = -3,   /* host or device allocation failed */
    BSM_ERR_UNSUPPORTED = -4
} bsm_status;

/* Kernel variants (benchmarked against each other; BSM_VARIANT_AUTO picks per plan). */
typedef enum {
    BSM_VARIANT_AUTO = 0,
    BSM_VARIANT_GATHER = 1,  /* owner-computes gather GEMV, direct global loads (two passes over
                                half-stored symmetric blocks) */
    BSM_VARIANT_FUSED = 2,   /* comparison: the stream plan with DIRECT global loads in the CTA kernel (no TMA); plans
                                with column sub-range slices or tall T-form blocks still use the TMA kernel */
    BSM_VARIANT_COLOR = 3,   /* colour-ordered multi-launch (the reference's schedule, for comparison):
                                y <- beta*y, then per sweep one launch per colour of a greedy colouring of the
                                "blocks share an output row" graph, each block accumulating straight into y */
    BSM_VARIANT_FUSED_TMA = 4 /* FUSED with the blocks streamed by cp.async.bulk (TMA) into an mbarrier
                                ring of shared-memory stages; what AUTO picks for symmetric matrices */
} bsm_variant;

/* Creation options; pass NULL for defaults. */
/* device = BSM_DEVICE_NONE builds a host-only handle: the packer runs and every table except the
 * arena can be exported (packing tests on machines without a GPU); bsm_mul* on it fail with
 * BSM_ERR_CUDA. */
#define BSM_DEVICE_NONE (-2)

typedef struct {
    int32_t device;        /* CUDA device ordinal; -1 = current device; BSM_DEVICE_NONE = host-only */
    int32_t variant;       /* bsm_variant used by bsm_mul unless overridden by bsm_set_variant */
    /* Owned output range of this rank's slab (0-based, half-open). Contributions whose outputs
     * fall outside are dropped; [0, -1) = everything. own_row_* applies to op N (outputs are
     * rows), own_col_* to op T/C (outputs are columns). */
    int64_t own_row_lo, own_row_hi;
    int64_t own_col_lo, own_col_hi;
    int64_t plan_hints;    /* 0 = automatic. Bits, for experiments and comparison runs: 1 = short segments go to the CTA-stream
                              kernel, not the warp-stream kernel; 2 = never cut the segments of a small (L2-resident) problem
                              into per-block work items; 4 = small problems keep the round-1 schedule (per-block work items, partial
                              sums through scratch, gather pass) instead of the single-launch CTA-part mode; 8 = experimental persistent
                              form of the CTA-stream kernel (sym_persist_kernel) */
    int64_t blocks_on_device; /* != 0: the block pointers handed to bsm_create_* are DEVICE pointers (blocks assembled on the
                              GPU): the arena is filled by a gather kernel in HBM, nothing crosses PCIe */
    int64_t reserved[2];
} bsm_options;

void bsm_default_options(bsm_options *opt);

/* BlockSparseMatrix: nb blocks; block b is m[b] x n[b]; its row / column index vectors are
 * rowidx[rowptr[b] .. rowptr[b+1]) and colidx[colptr[b] .. colptr[b+1]) (pool offsets 0-based,
 * index VALUES 1-based), with rowptr[b+1]-rowptr[b] == m[b] and colptr[b+1]-colptr[b] == n[b]. */
int bsm_create_blocksparse(int dtype, int64_t nrows, int64_t ncols, int64_t nb,
                           const void *const *blocks, const int64_t *m, const int64_t *n,
                           const int64_t *rowidx, const int64_t *rowptr, const int64_t *colidx,
                           const int64_t *colptr, const bsm_options *opt, bsm_handle *out);

/* SymmetricBlockMatrix: ndiag square diagonal blocks (dsize[d] x dsize[d], index vector
 * didx[dptr[d] .. dptr[d+1])) and noff half-stored off-diagonal blocks (om[b] x on[b], row / column
 * index vectors as above). */
int bsm_create_symmetric(int dtype, int64_t nrows, int64_t ncols, int64_t ndiag,
                         const void *const *diag, const int64_t *dsize, const int64_t *didx,
                         const int64_t *dptr, int64_t noff, const void *const *off,
                         const int64_t *om, const int64_t *on, const int64_t *rowidx,
                         const int64_t *rowptr, const int64_t *colidx, const int64_t *colptr,
                         const bsm_options *opt, bsm_handle *out);

/* VBCRS: nbrows block rows; rowptr (1-based, length nbrows+1, sentinel nb+1), colstart[b] (1-based
 * first column of block b), rowstart[r] (1-based first row of block row r); block b is m[b] x n[b].
 * is_transposed (nullable): is_transposed[b] != 0 means blocks[b] points at the PARENT of a lazy
 * `transpose(parent)` wrapper (parent is n[b] x m[b] column-major) as produced by the
 * SymmetricBlockMatrix → VBCRS conversion (src/vbcrs.jl:222-264); it is materialised while packing. */
int bsm_create_vbcrs(int dtype, int64_t nrows, int64_t ncols, int64_t nbrows, int64_t nb,
                     const int64_t *rowptr, const int64_t *colstart, const int64_t *rowstart,
                     const void *const *blocks, const int64_t *m, const int64_t *n,
                     const uint8_t *is_transposed, const bsm_options *opt, bsm_handle *out);

/* New VALUES, same structure (a BEM matrix re-assembled for another frequency, a time step, ...): re-uploads
 * the blocks into the existing arena without re-planning. `blocks` holds nb host pointers in creation order
 * (symmetric: the diagonal blocks, then the off-diagonal blocks), each of the shape (and, for VBCRS, the
 * is_transposed flag) given at creation. Synchronises the device first. */
int bsm_update_values(bsm_handle h, const void *const *blocks, int64_t nb);
/* The same with DEVICE block pointers (a host array of nb device pointers): gathered into the arena in HBM. */
int bsm_update_values_dev(bsm_handle h, const void *const *dev_blocks, int64_t nb);

/* The sorting constructor of VariableBlockCompressedRowStorage on the device (src/vbcrs.jl:78-122): from the unsorted
 * block list (1-based Int64 row start and column start per block, DEVICE arrays of nb entries) computes the STABLE
 * sort permutation by (row start, column start) — perm[k] = 0-based input index of the block that becomes block k —,
 * the block-row pointer (1-based, nbrows+1 entries used, sentinel nb+1), the start row of every block row and the
 * sorted column starts; all outputs are DEVICE arrays of nb (+1) entries. Bit-identical to the host constructor. */
int bsm_vbcrs_sort_dev(int64_t nb, const int64_t *rowstart_dev, const int64_t *colstart_dev, int64_t *perm_dev,
                       int64_t *rowptr_dev, int64_t *rowindices_dev, int64_t *colindices_dev, int64_t *nbrows_out,
                       void *stream);

int bsm_destroy(bsm_handle h);

/* y[:, j] = alpha * op(A) * x[:, j] + beta * y[:, j], j < nrhs; x and y are DEVICE pointers of the
 * matrix dtype, column-major with leading dimensions ldx / ldy (elements). `stream` is a
 * cudaStream_t cast to void* (NULL = default stream). Asynchronous with respect to the host. */
int bsm_mul(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
            const void *x_dev, int64_t ldx, void *y_dev, int64_t ldy, int64_t nrhs, void *stream);

/* Same with HOST x / y (pageable or pinned): H2D of x (and of y when beta is used) into device buffers owned by
 * the handle, multiply, D2H of y, then synchronises (serialised per handle). Slab handles (bsm_options.own_*): only the
 * owned rows of y are read and written, the other rows of y_host are left alone. */
int bsm_mul_host(bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                 const void *x_host, int64_t ldx, void *y_host, int64_t ldy, int64_t nrhs);

int bsm_set_variant(bsm_handle h, int variant);

/* Per-kernel device timing for roofline reports: when on, bsm_mul brackets its multiply kernels and its gather
 * (finalize) kernel with CUDA events on the launch stream, one event triple per call in a ring of 64.
 * bsm_get_profile synchronises on the recorded events and returns the AVERAGE milliseconds of the main multiply
 * kernel(s) and of the gather kernel over the calls made since profiling was switched on / last read (at most
 * the last 64), then resets — so a timed region of back-to-back multiplies can be read afterwards. Not
 * thread-safe; benchmarking only. */
int bsm_set_profiling(bsm_handle h, int on);
int bsm_get_profile(bsm_handle h, double *main_ms, double *finalize_ms);

/* ---- queries ---------------------------------------------------------------------------- */
int64_t bsm_nnz(bsm_handle h);          /* SparseArrays.nnz semantics (symmetric: 2*off + diag) */
int64_t bsm_stored_entries(bsm_handle h); /* entries held in the arena (half-stored counted once) */
int bsm_size(bsm_handle h, int64_t *nrows, int64_t *ncols);
int bsm_dtype_of(bsm_handle h);
int bsm_kind_of(bsm_handle h);
/* Algorithmic bytes / flops of one multiply with nrhs right-hand sides:
 * bytes = stored*s + (ncols + nrows*(1 + beta_used))*s*nrhs + index_table_bytes. */
int bsm_work(bsm_handle h, int op, int64_t nrhs, int beta_used, double *bytes, double *flops,
             double *index_table_bytes);
/* Number of kernels one bsm_mul launches for `op` with the current variant (nrhs = 1). */
int bsm_launch_count(bsm_handle h, int op);
/* Work split of the plan bsm_mul uses for `op` with the current variant. out[0..2] = slices handled by
 * the CTA-stream kernel (sym_fused_tma_kernel), the warp-stream kernel (stream_warp_kernel) and the
 * gather kernel (gather_gemv_kernel); out[3..5] = block bytes each of them streams; out[6] = warp work
 * items; out[7] = warp-stream chunks; out[8] = scratch elements; out[9] = rows finalised by the gather pass;
 * out[10] = CTA work items of the multi-RHS (SpMM) kernels, 0 if the plan is not eligible for them (then
 * nrhs > 1 loops over the columns); out[11] = 1 if the plan is eligible for spmm_tma_kernel (every dtype), else
 * multi-RHS products use spmm_dmma_kernel (Float64 only). */
int bsm_plan_stats(bsm_handle h, int op, int64_t out[12]);

/* ---- table export (bit-exact packing checks) ---------------------------------------------- */
typedef enum {
    BSM_TAB_ARENA = 0,        /* dtype elements, device arena copied back */
    BSM_TAB_BLOCK_OFF = 1,    /* int64: element offset of every block in the arena (128-byte aligned) */
    BSM_TAB_BLOCK_M = 2,      /* int32 */
    BSM_TAB_BLOCK_N = 3,      /* int32 */
    BSM_TAB_SET_LEN = 4,      /* int32: index-set table (deduplicated row/column index vectors) */
    BSM_TAB_SET_START = 5,    /* int32: 0-based first index if the set is a contiguous range, else -1 */
    BSM_TAB_SET_POOL_OFF = 6, /* int64: offset into the index pool (only if SET_START < 0) */
    BSM_TAB_POOL = 7,         /* int32: 0-based indices */
    /* per-plan tables (GATHER variant: plan 0 = op N, plan 1 = op T/C; FUSED variant, symmetric
     * matrices only: plan 2 = op N, plan 3 = op T/C): */
    BSM_TAB_CONTRIB = 8,      /* bsm_contrib records, grouped by output segment */
    BSM_TAB_SLICE = 9,        /* bsm_slice records (work items) */
    BSM_TAB_GATHER_ROWS = 10, /* int32: rows finalised by the gather pass (bit 31: already written) */
    BSM_TAB_GATHER_PTR = 11,  /* int64 */
    BSM_TAB_GATHER_POS = 12,  /* int64: positions in the partial-sum scratch */
    BSM_TAB_GROUP_PTR = 13,   /* int64: CSR over contributions per output segment ("block-row pointer";
                                 for plan 1 this is the transposed index) */
    BSM_TAB_GROUP_SET = 14,   /* int32: index-set id of every output segment */
    BSM_TAB_CONTRIB_TOFF = 15,/* int64: scratch offset of the fused transposed partial of a contribution, -1 if none */
    BSM_TAB_WCHUNK = 16,      /* bsm_wchunk records: the chunk stream of the warp-stream kernel (plans 2/3) */
    BSM_TAB_WITEM_PTR = 17,   /* int32: chunk range [ptr[i], ptr[i+1]) of warp work item i */
    BSM_TAB_COLOR_PTR = 18    /* int32: colour-ordered plans (4 = op N, 5 = op T/C): launch l runs slices
                                 [ptr[l], ptr[l+1]) */
} bsm_table;

/* 32-byte device records (exported verbatim). */
typedef struct {
    int64_t off;       /* arena element offset of the block */
    int32_t m, n;      /* block is m x n, column-major, ld = m */
    int32_t in_set;    /* index set gathered from x */
    int32_t form;      /* bit0 = 0: y[out] += op(B) x[in] along block rows ("N-form");
                          bit0 = 1: y[out] += op(B)^T x[in] along block columns ("T-form");
                          bit1: the same pass also emits t = op(B)^T x[out rows] to scratch (fused
                          transposed partial of a half-stored symmetric block) */
    int32_t out_len;   /* outputs this block covers inside its segment (m or n) */
    int32_t block;     /* block id */
} bsm_contrib;

typedef struct {
    int32_t out_set;     /* index set of the output segment */
    int32_t r0, r1;      /* output sub-range [r0, r1) of the segment handled by this work item */
    int32_t c_begin, c_end; /* contributions [c_begin, c_end) */
    int32_t flags;       /* bit4: slab handle only — some input lies outside the rank's own x range (runs after the
                            all-gather; the other slices overlap with it);
                            bit0: direct (writes y), else partial sums to scratch; bit1: vector loads ok;
                            bit2: handled by the TMA-staged CTA kernel (whole segment of <= 256 rows, or a
                            column sub-range of an all-T-form segment); bit3: whole segment of <= 64 rows,
                            handled by the warp-stream kernel through the bsm_wchunk stream */
    int64_t scratch_off; /* element offset of the partial vector when not direct */
} bsm_slice;

/* One TMA bulk copy of the warp-stream kernel: `ncols` whole columns of one block (m <= 64 rows), with
 * everything the consuming warp needs so that it never chases a pointer on its critical path. The
 * shared-memory placement of every chunk in the warp's byte ring is decided by the packer (the chunk
 * sequence of a work item is static): the chunk occupies [smem16*16, smem16*16 + bytes16*16 + x area)
 * and may be issued once all but `lag` of the chunks before it in its work item have been consumed. */
typedef struct {
    uint32_t src16;      /* arena byte offset / 16 of the copy (start rounded down to 16 bytes), low 32 bits */
    uint16_t bytes16;    /* copy size / 16 */
    uint16_t ncols;      /* whole columns in the chunk (<= 64) */
    uint8_t m;           /* rows of the block (<= 64) */
    uint8_t flags;       /* bit0 T-form; bit1 x_ref is a pool offset; bit2 out is a pool offset;
                            bit3 first chunk of its segment; bit4 last chunk of its segment; bit5 direct */
    uint8_t delta;       /* byte offset of the first column inside the copy (0..15) */
    uint8_t seg_len;     /* rows of the output segment (<= 64) */
    int32_t x_ref;       /* N-form: position in x of the chunk's first column; T-form: of the block's
                            first row (an index into x, or into the index pool when bit1) */
    uint8_t out_col;     /* T-form: position of the chunk's first column inside the segment */
    uint8_t src16_hi;    /* bits 32..39 of the 16-byte unit offset */
    uint16_t smem16;     /* offset / 16 of the chunk in the warp's shared-memory ring */
    uint8_t lag;         /* chunk i of a work item may be issued once i - lag chunks have been consumed */
    uint8_t reserved[3];
    int64_t out;         /* last chunk of a segment: first output row (index into y, or pool offset when
                            bit2) if direct, else element offset of the partial vector in the scratch */
} bsm_wchunk;

int64_t bsm_table_count(bsm_handle h, int table, int plan); /* number of elements/records, <0 on error */
int bsm_table_copy(bsm_handle h, int table, int plan, void *dst, int64_t dst_bytes);

/* ---- sparse(A) on the device ---------------------------------------------------
 * SparseArrays.sparse(op(A)) (src/sparse.jl:17-129) built from the resident arena: canonical CSC — columns in
 * order, rows sorted inside a column, duplicates (overlapping blocks) summed, explicit zeros kept — with
 * 1-based Int64 colptr / rowval exactly as a Julia SparseMatrixCSC holds them. bsm_sparse_build leaves the
 * three arrays on the device (bsm_sparse_device_pointers, e.g. for cuSPARSE); bsm_sparse_fetch copies them to
 * host arrays of ncols+1, nnz and nnz elements and releases the device copy. Not available on slab handles. */
int bsm_sparse_build(bsm_handle h, int op, int64_t *nnz_out);
int bsm_sparse_fetch(bsm_handle h, int64_t *colptr, int64_t *rowval, void *nzval);
int bsm_sparse_device_pointers(bsm_handle h, void **colptr_dev, void **rowval_dev, void **nzval_dev, int64_t *nnz);

/* ---- single-box multi-GPU: block-row slabs + NCCL all-gather of x ------------------
 * One process per GPU. Rank r builds its handle from the blocks of its slab with
 * bsm_options.own_row_* / own_col_* = its output range, keeps a FULL-length x on its device and owns
 * rows [cuts[r], cuts[r+1]) of it. bsm_mul_dist replicates x (slabs are uneven: equal-chunk staging buffer, one
 * ncclAllGather, strided copies back) while the slices fed by the rank's own x slab already run, then
 * multiplies the rest;
 * each rank writes only its own slice of y. NCCL is bound at run time (dlopen "libnccl.so.2"). The
 * reference has no counterpart (single process). */
typedef struct bsm_comm_s *bsm_comm;
#define BSM_DIST_ID_BYTES 128
int bsm_dist_unique_id(void *id128);       /* rank 0: ncclGetUniqueId; ship the 128 bytes to every rank */
int bsm_dist_init(const void *id128, int nranks, int rank, int device, bsm_comm *out);
int bsm_dist_destroy(bsm_comm c);
/* on (default): bsm_mul_dist runs the all-gather on the communicator's own stream while the slices whose
 * inputs lie in this rank's x slab run on the caller's stream; the remote slices follow. off: gather, then
 * multiply, on one stream (comparison). */
int bsm_dist_set_overlap(bsm_comm c, int on);
/* 0 (default): equal-chunk staging + one ncclAllGather; 1: one NCCL group of in-place broadcasts per rank and
 * right-hand side (no staging; 10x slower on 8 GPUs, kept for comparison). */
int bsm_dist_set_collective(bsm_comm c, int use_broadcasts);
/* Benchmarking only: bit0 = peer-mode multiplies skip the wait of the entry barrier, bit1 = of the exit barrier
 * (results are then only valid when x does not change between multiplies, or when a collective sits between every
 * multiply and the next write of x — bsm_cg_dist sets this bit itself: the all-reduce of p.q is that collective);
 * bit2 = every arrival waits at system
 * scope; bit3 = the kernels stamp %globaltimer around the barriers; bit4 = release instead of relaxed signal stores.
 * bsm_dist_debug_read returns (and resets) the sums over the multiplies since the last read: out[0] ns between a
 * kernel's first arrival and the end of its entry wait, out[1] ns first arrival -> last arrival, out[2] ns of the
 * exit wait, out[3] multiplies counted. */
int bsm_dist_set_debug(bsm_comm c, int flags);
int bsm_dist_debug_read(bsm_comm c, int64_t out[8]);
int bsm_dist_info(bsm_comm c, int *nranks, int *rank, int *nccl_version);
/* In-place all-gather of the row slabs of a column-major (rows x nrhs, leading dimension ldx) DEVICE
 * array: on return every rank holds all rows. cuts has nranks+1 entries (0-based, non-decreasing). */
int bsm_dist_allgather_rows(bsm_comm c, int dtype, void *x_dev, int64_t ldx, int64_t nrhs, const int64_t *cuts,
                            void *stream);
int bsm_dist_allreduce_max_f64(bsm_comm c, double *dev_values, int64_t count, void *stream);
int bsm_dist_allreduce_sum_f64(bsm_comm c, double *dev_values, int64_t count, void *stream);
/* Peer mode — the all-gather fused into the multiply. bsm_dist_alloc is collective: every rank allocates `bytes`
 * on its GPU and maps every peer's allocation (CUDA IPC, NVLink peer access). Keep a full-length x in such an
 * array; a rank only ever writes its own slab. bsm_mul_dist_peer (nrhs = 1) runs NO collective and launches NO
 * extra kernel: the multiply kernels themselves signal "my x slab of this epoch is written" (first CTA to start),
 * wait for every peer's signal before their first x fetch, fetch each x element from its owner's array over NVLink
 * with the same asynchronous copies that stage it from local HBM (only the entries the rank's blocks touch ever
 * cross the link), and the last CTA to finish signals "done reading" and waits for every peer's — so the kernel
 * completes only when whatever follows on the stream (the solver's update of the slab) may overwrite x. Every rank
 * must call it the same number of times, one multiply at a time per communicator. */
int bsm_dist_alloc(bsm_comm c, size_t bytes, void **dev_ptr);
int bsm_dist_free(bsm_comm c, void *dev_ptr);   /* only after every rank has finished its last multiply on it */
int bsm_mul_dist_peer(bsm_comm c, bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                      void *x_shared, void *y_dev, const int64_t *in_cuts, void *stream);
/* The same collective multiply with this rank's slabs in HOST memory (what `mul!` on host arrays maps to when the
 * operator is sharded): H2D of x_host_slab (rows in_cuts[rank] .. in_cuts[rank+1]) into x_shared (and of y_host_slab
 * when beta is used), bsm_mul_dist_peer, D2H of rows [out_lo, out_hi) of y_dev into y_host_slab, then synchronises
 * the stream. y_dev is a full-length device work array. */
int bsm_mul_dist_peer_host(bsm_comm c, bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                           const void *x_host_slab, void *x_shared, void *y_dev, void *y_host_slab,
                           const int64_t *in_cuts, int64_t out_lo, int64_t out_hi, void *stream);
/* all-gather of x over in_cuts, then bsm_mul on this rank's slab. */
int bsm_mul_dist(bsm_comm c, bsm_handle h, int op, const void *alpha, const void *beta, int beta_is_false,
                 void *x_dev, int64_t ldx, void *y_dev, int64_t ldy, int64_t nrhs, const int64_t *in_cuts,
                 void *stream);

/* ---- solver loop on the device ----------------------------------------------------
 * The reference is an operator for Krylov solvers driven through LinearMaps (docs/src/block.md:56-63 times the three
 * products they are built from); bsm_cg keeps such a loop on the GPU: conjugate gradients for A x = b, x0 = 0, with
 * the Hermitian inner product (hermitian != 0: CG, Hermitian positive definite operators) or the unconjugated
 * bilinear form (hermitian = 0: COCG, for the complex SYMMETRIC operators a SymmetricBlockMatrix{ComplexF64} holds;
 * the same as CG for real dtypes). Per iteration: one bsm_mul (q = A p) and three fused vector kernels
 * (p.q | x += alpha p, r -= alpha q, r.r, |r|^2 | p = r + beta p); alpha, beta and the residual history stay on the
 * device, the host synchronises once per `check_every` iterations; reductions use a fixed number of partial sums and
 * fixed trees (bitwise reproducible). Stops when |r| <= rtol * |b| or after maxit iterations; returns the iterations
 * run and the last |r| / |b|.
 * bsm_cg_dist: the operator is sharded into block-row slabs (slab handle + communicator, cuts as in
 * bsm_mul_dist_peer); b_dev / x_dev are full-length device arrays of which this rank reads / writes rows
 * cuts[rank] .. cuts[rank+1]; the search direction lives in a peer-mapped array the peers read over NVLink, the two
 * dot products per iteration are summed over the ranks with ncclAllReduce. Collective. */
typedef struct {
    double rtol;
    int64_t maxit;
    int32_t hermitian;
    int32_t check_every;
} bsm_cg_options;
void bsm_cg_default_options(bsm_cg_options *opt);   /* rtol 1e-10, maxit 200, hermitian 0, check_every 8 */
int bsm_cg(bsm_handle h, const void *b_dev, void *x_dev, const bsm_cg_options *opt, int64_t *iters, double *relres,
           void *stream);
int bsm_cg_dist(bsm_comm c, bsm_handle h, const void *b_dev, void *x_dev, const int64_t *cuts, const bsm_cg_options *opt,
                int64_t *iters, double *relres, void *stream);

/* ---- device memory helpers for callers without a CUDA array package ------------------------ */
int bsm_device_count(int *count);
int bsm_malloc(int device, size_t bytes, void **dev_ptr);
int bsm_free(int device, void *dev_ptr);
int bsm_memcpy_h2d(void *dev_dst, const void *host_src, size_t bytes);
int bsm_memcpy_d2h(void *host_dst, const void *dev_src, size_t bytes);
int bsm_synchronize(int device);

const char *bsm_last_error(void);
const char *bsm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BSM_B200_H */
