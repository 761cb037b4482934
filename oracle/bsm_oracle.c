/* bsm_oracle.c — CPU restatement of BlockSparseMatrices.jl's multiply path.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the checker for the B200 path, never the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. The product (blocksparsematrices.jl_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned" for output vectors. The reference is Julia; no Julia runtime
 * exists in this image or on the GPU box, and the reference's tests hold no golden output
 * vectors (they are property tests against sparse(A) products with unseeded randn inputs,
 * /root/reference/test/test_blockmatrix.jl:51-81, test/test_vbcrs.jl:33-48). The oracle is
 * therefore pinned by (a) the reference's own test properties re-run on the one fixture the
 * reference ships (test/assets/symmetricblockexamples.jld2 → tests/golden/ npz files): agreement with
 * an independent CSC product of the restated sparse(A) to 1e-13, issymmetric(sparse(A)), the nnz
 * identities; (b) agreement between this C restatement and a separate NumPy restatement
 * (oracle/oracle_np.py). The per-block arithmetic (LinearAlgebra.mul! on views) lives in the Julia
 * stdlib, which is not under /root/reference; its loop order is restated, not verified.
 *
 * Algorithms restated (file:line are relative to /root/reference):
 *   src/abstractblockmatrix.jl:27-34   3-arg → 5-arg with (α, β) = (true, false)
 *   src/blockmatrix.jl:225-247         BlockSparseMatrix multiply, colour by colour
 *   src/symmetricblockmatrix.jl:386-435 SymmetricBlockMatrix multiply, three sweeps
 *   src/vbcrs.jl:266-288, 303-354      VBCRS forward (block-row parallel) / transposed (serial)
 *   src/coloring.jl:45-61              conflict definition (two blocks conflict iff their index
 *                                      vectors share an index); the colouring ALGORITHM
 *                                      (GraphsColoring.WorkstreamDSATUR, external) is replaced by
 *                                      a first-fit greedy colouring — results do not depend on it.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared).
 */
#include <complex.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define T float
#define SFX f32
#define CONJ(v) (v)
#include "oracle_body.inc"
#undef T
#undef SFX
#undef CONJ

#define T double
#define SFX f64
#define CONJ(v) (v)
#include "oracle_body.inc"
#undef T
#undef SFX
#undef CONJ

#define T double _Complex
#define SFX c64
#define CONJ(v) conj(v)
#include "oracle_body.inc"
#undef T
#undef SFX
#undef CONJ

enum { ORACLE_F32 = 0, ORACLE_F64 = 1, ORACLE_C64 = 2 };

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* alpha / beta are passed by pointer as one element of the matrix dtype. */
int oracle_bsm_mul(int dtype, int op, int64_t nb, const void *const *blocks, const int64_t *m,
                   const int64_t *n, const int64_t *ridx, const int64_t *rptr, const int64_t *cidx,
                   const int64_t *cptr, int64_t ncolors, const int64_t *color_ptr,
                   const int64_t *color_blocks, const void *alpha, const void *beta,
                   int beta_is_false, const void *x, void *y, int64_t ny, int nthreads) {
    switch (dtype) {
    case ORACLE_F32:
        bsm_mul_f32(op, nb, (const float *const *)blocks, m, n, ridx, rptr, cidx, cptr, ncolors,
                    color_ptr, color_blocks, *(const float *)alpha, *(const float *)beta,
                    beta_is_false, (const float *)x, (float *)y, ny, nthreads);
        return 0;
    case ORACLE_F64:
        bsm_mul_f64(op, nb, (const double *const *)blocks, m, n, ridx, rptr, cidx, cptr, ncolors,
                    color_ptr, color_blocks, *(const double *)alpha, *(const double *)beta,
                    beta_is_false, (const double *)x, (double *)y, ny, nthreads);
        return 0;
    case ORACLE_C64:
        bsm_mul_c64(op, nb, (const double _Complex *const *)blocks, m, n, ridx, rptr, cidx, cptr,
                    ncolors, color_ptr, color_blocks, *(const double _Complex *)alpha,
                    *(const double _Complex *)beta, beta_is_false, (const double _Complex *)x,
                    (double _Complex *)y, ny, nthreads);
        return 0;
    }
    return -1;
}

int oracle_sbm_mul(int dtype, int op, int64_t ndiag, const void *const *diag, const int64_t *dsz,
                   const int64_t *didx, const int64_t *dptr, int64_t noff, const void *const *off,
                   const int64_t *om, const int64_t *on, const int64_t *ridx, const int64_t *rptr,
                   const int64_t *cidx, const int64_t *cptr, int64_t nc_row,
                   const int64_t *crow_ptr, const int64_t *crow_blk, int64_t nc_col,
                   const int64_t *ccol_ptr, const int64_t *ccol_blk, int64_t nc_diag,
                   const int64_t *cdiag_ptr, const int64_t *cdiag_blk, const void *alpha,
                   const void *beta, int beta_is_false, const void *x, void *y, int64_t ny,
                   int nthreads) {
#define SBM_ARGS(TT)                                                                             \
    op, ndiag, (const TT *const *)diag, dsz, didx, dptr, noff, (const TT *const *)off, om, on,  \
        ridx, rptr, cidx, cptr, nc_row, crow_ptr, crow_blk, nc_col, ccol_ptr, ccol_blk, nc_diag, \
        cdiag_ptr, cdiag_blk, *(const TT *)alpha, *(const TT *)beta, beta_is_false,              \
        (const TT *)x, (TT *)y, ny, nthreads
    switch (dtype) {
    case ORACLE_F32: sbm_mul_f32(SBM_ARGS(float)); return 0;
    case ORACLE_F64: sbm_mul_f64(SBM_ARGS(double)); return 0;
    case ORACLE_C64: sbm_mul_c64(SBM_ARGS(double _Complex)); return 0;
    }
#undef SBM_ARGS
    return -1;
}

int oracle_vbcrs_mul(int dtype, int op, int64_t nbrows, const int64_t *rowptr,
                     const int64_t *colstart, const int64_t *rowstart, const void *const *blocks,
                     const int64_t *m, const int64_t *n, const void *alpha, const void *beta,
                     int beta_is_false, const void *x, void *y, int64_t ny, int nthreads) {
#define VB_ARGS(TT)                                                                          \
    op, nbrows, rowptr, colstart, rowstart, (const TT *const *)blocks, m, n,                 \
        *(const TT *)alpha, *(const TT *)beta, beta_is_false, (const TT *)x, (TT *)y, ny, nthreads
    switch (dtype) {
    case ORACLE_F32: vbcrs_mul_f32(VB_ARGS(float)); return 0;
    case ORACLE_F64: vbcrs_mul_f64(VB_ARGS(double)); return 0;
    case ORACLE_C64: vbcrs_mul_c64(VB_ARGS(double _Complex)); return 0;
    }
#undef VB_ARGS
    return -1;
}

/* First-fit greedy colouring of the conflict graph defined at
 * /root/reference/src/coloring.jl:45-61 (blocks are vertices; an edge joins two blocks whose
 * index vectors share an index). Stands in for GraphsColoring's WorkstreamDSATUR.
 * idx/ptr: concatenated 1-based index vectors; maxidx = largest index.
 * color_out[b] receives the 0-based colour; returns the number of colours, or -1 on failure.
 * Per index a growable bitset of the colours already used by a block touching it. */
int64_t oracle_greedy_color(int64_t nb, const int64_t *idx, const int64_t *ptr, int64_t maxidx,
                            int64_t *color_out) {
    int64_t words = 1; /* 64 colours per word */
    uint64_t *used = (uint64_t *)calloc((size_t)(maxidx + 1) * words, sizeof(uint64_t));
    uint64_t *acc = (uint64_t *)calloc((size_t)words, sizeof(uint64_t));
    if (!used || !acc) return -1;
    int64_t ncolors = 0;
    for (int64_t b = 0; b < nb; ++b) {
        for (;;) {
            memset(acc, 0, (size_t)words * sizeof(uint64_t));
            for (int64_t q = ptr[b]; q < ptr[b + 1]; ++q) {
                const uint64_t *u = used + (size_t)idx[q] * words;
                for (int64_t w = 0; w < words; ++w) acc[w] |= u[w];
            }
            int64_t c = -1;
            for (int64_t w = 0; w < words && c < 0; ++w) {
                if (~acc[w]) c = w * 64 + __builtin_ctzll(~acc[w]);
            }
            if (c >= 0) {
                for (int64_t q = ptr[b]; q < ptr[b + 1]; ++q)
                    used[(size_t)idx[q] * words + c / 64] |= (uint64_t)1 << (c % 64);
                color_out[b] = c;
                if (c + 1 > ncolors) ncolors = c + 1;
                break;
            }
            /* all colours of the current width are taken: widen the bitsets and retry */
            int64_t nwords = words * 2;
            uint64_t *nu = (uint64_t *)calloc((size_t)(maxidx + 1) * nwords, sizeof(uint64_t));
            uint64_t *na = (uint64_t *)calloc((size_t)nwords, sizeof(uint64_t));
            if (!nu || !na) return -1;
            for (int64_t i = 0; i <= maxidx; ++i)
                memcpy(nu + (size_t)i * nwords, used + (size_t)i * words, (size_t)words * 8);
            free(used);
            free(acc);
            used = nu;
            acc = na;
            words = nwords;
        }
    }
    free(used);
    free(acc);
    return ncolors;
}
