// persist.cuh — persistent form of the CTA-stream kernel (sym_fused_tma_kernel, lean plans: whole segments of <= 256
// rows whose T-form blocks are no taller than the rows the lanes hold — every SymmetricBlockMatrix plan, C2).
//
// Why: sym_fused_tma_kernel pays ~4 us of dead time per CTA (slice -> contribution -> index set -> x: four dependent
// global loads before the first FMA, and a reduction through the ring at the end) while its consumers, not HBM, are
// the limit — the other CTA of the SM cannot make up for it. On C2 that is 6 % of a multiply on one GPU (30 CTAs per
// slot) and 12–14 % on the 1.6 GB slabs of 8 GPUs (measured). Here a CTA stays resident and walks its
// share of the work items (longest first, dealt in snake order), and everything a consumer used to fetch from global
// memory is staged for it by two helper warps that run AHEAD across item boundaries:
//   warp 8  (one lane)   the TMA issuer: cp.async.bulk chunks of every block into the 4-stage ring, exactly as before,
//                        but never stopping at the end of an item;
//   warp 9  (32 lanes)   the x stager: per "window" (<= 512 columns of an N-form block, or the rows of a T-form block)
//                        it gathers the x entries into one of two shared-memory buffers and writes a 64-byte window
//                        descriptor (sizes, form, scratch offsets, and for the last window of an item where the
//                        outputs go); per item it also stages x at the segment's own rows (fused transposed partials);
//   warps 0-7            consumers: wait for the window (mbarrier), then for each of its chunks; they touch global
//                        memory only to write y / scratch.
// The ring is no longer reused for the cross-warp reduction (the issuer is already filling it for the next item): the
// eight warps add their row sums into one 256-entry vector in warp order (deterministic), ~0.3 us.
#pragma once
#include "kernels.cuh"

namespace bsm {

constexpr int kQThreads = kFThreads + 64;      // 8 consumer warps + TMA issuer warp + x stager warp
constexpr int kQWin = 512;                     // x entries per window buffer
constexpr int kQMaxCtasPerSm = 2;

struct alignas(16) WinDesc {
    int32_t m;            // rows of the block
    int32_t jw;           // first column of the window inside the block (N-form), 0 for T-form
    int32_t wcols;        // columns the window covers (N-form) / output columns of the T-form block inside the segment
    int32_t flags;        // bit0 T-form, bit1 fused transposed partial, bit2 first window of an item, bit3 last window of an
                          // item, bit4 the item writes y directly, bit5 end of this CTA's work
    int64_t toff;         // fused partial: scratch offset of the contribution's partial vector
    int32_t L;            // rows of the segment
    int32_t out_start;    // first output row if the segment is a contiguous range, else -1
    int64_t out_pool;     // pool offset of the segment's rows when out_start < 0
    int64_t scratch_off;  // item not direct: offset of its partial vector
    int64_t pad;
};
static_assert(sizeof(WinDesc) == 64, "window descriptor is 64 bytes");

template <class T>
struct PersistSmem {
    static constexpr size_t ring = (size_t)kPStages * kPStageBytes;
    static constexpr size_t xbuf = 2 * (size_t)kQWin * sizeof(T);
    static constexpr size_t xrs = 2 * (size_t)kFMaxRows * sizeof(T);
    static constexpr size_t acct = (size_t)kFMaxRows * sizeof(T);
    static constexpr size_t wdesc = 2 * sizeof(WinDesc);
    static constexpr size_t bars = 8 * (2 * kPStages + 4);
    static constexpr size_t total = ring + xbuf + xrs + acct + wdesc + bars;
};

// item k of this CTA (snake order over the longest-first slice list: balanced totals without a queue)
__device__ __forceinline__ int32_t persist_item(int32_t k, int32_t nslices) {
    const int32_t G = gridDim.x, c = blockIdx.x;
    const int32_t it = k * G + ((k & 1) ? (G - 1 - c) : c);
    return it < nslices ? it : -1;
}
__device__ __forceinline__ int32_t persist_item_count(int32_t nslices) {
    // number of k with persist_item(k) valid: tiers are full except possibly the last one
    const int32_t G = gridDim.x, full = nslices / G, rem = nslices - full * G, c = blockIdx.x;
    const int32_t pos = (full & 1) ? (G - 1 - c) : c;
    return full + (pos < rem ? 1 : 0);
}

template <class T>
__device__ __forceinline__ void persist_issuer(const MulArgs<T> &a, unsigned char *stages, uint64_t *full, uint64_t *empty) {
    const uint64_t policy = l2_evict_first_policy();
    uint32_t q = 0;
    const int32_t nit = persist_item_count(a.nslices);
    // metadata is fetched one step ahead (next slice, next contribution): the loads are in flight while the chunks of
    // the current block are issued, so the ring never drains at an item or block boundary
    bsm_slice sl_next = nit > 0 ? a.slices[persist_item(0, a.nslices)] : bsm_slice{};
    bsm_contrib cb_next = bsm_contrib{};
    int32_t ci_next = -1;          // which contribution cb_next holds
    if (nit > 0 && sl_next.c_begin < sl_next.c_end) {
        ci_next = sl_next.c_begin;
        cb_next = a.contrib[ci_next];
    }
    for (int32_t k = 0; k < nit; ++k) {
        const bsm_slice sl = sl_next;
        if (k + 1 < nit) sl_next = a.slices[persist_item(k + 1, a.nslices)];
        for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
            const bsm_contrib cb = (ci_next == ci) ? cb_next : a.contrib[ci];
            if (ci + 1 < sl.c_end)
                ci_next = ci + 1;
            else if (k + 1 < nit && sl_next.c_begin < sl_next.c_end)
                ci_next = sl_next.c_begin;
            else
                ci_next = -1;
            if (ci_next >= 0) cb_next = a.contrib[ci_next];
            const int32_t m = cb.m;
            const bool tform = (cb.form & 1) != 0;
            const int32_t jhi = tform ? min(sl.r1, cb.out_len) : cb.n;
            if (jhi <= 0 || m == 0) continue;
            const int32_t cc = chunk_cols<T>(m);
            const unsigned char *blk = reinterpret_cast<const unsigned char *>(a.arena + cb.off);
            const int32_t wstep = tform ? jhi : kQWin;       // windows never split a chunk sequence differently from the consumers
            for (int32_t jw = 0; jw < jhi; jw += wstep) {
                const int32_t wend = min(jhi, jw + wstep);
                for (int32_t j0 = jw; j0 < wend; j0 += cc, ++q) {
                    const int32_t ncols = min(cc, wend - j0);
                    const uint32_t stage = q % kPStages;
                    if (q >= kPStages) mbar_wait(&empty[stage], ((q / kPStages) - 1) & 1);
                    const int64_t boff = (int64_t)j0 * m * (int64_t)sizeof(T);
                    const uint32_t delta = (uint32_t)(boff & 15);
                    const uint32_t bytes = (delta + (uint32_t)ncols * (uint32_t)m * (uint32_t)sizeof(T) + 15u) & ~15u;
                    mbar_arrive_expect_tx(&full[stage], bytes);
                    bulk_g2s(stages + stage * kPStageBytes, blk + (boff - delta), bytes, &full[stage], policy);
                }
            }
        }
    }
}

// Gathers n (<= 512) entries dst[i] = i < nvalid ? x[idx(i0 + i)] : 0 with the 32 lanes of the stager warp: the loads of
// a batch are all issued before the first store, so a window costs one memory round trip, not n/32 of them.
template <class T>
__device__ __forceinline__ void stage_gather(const MulArgs<T> &a, const SetRef &set, int32_t i0, int32_t nvalid, int32_t n,
                                             T *dst, int lane) {
    constexpr int B = 8;
    for (int32_t base = 0; base < n; base += 32 * B) {
        T v[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int32_t i = base + u * 32 + lane;
            v[u] = (i < nvalid) ? a.x.at(set.at(i0 + i)) : El<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int32_t i = base + u * 32 + lane;
            if (i < n) dst[i] = v[u];
        }
    }
}

template <class T>
__device__ __forceinline__ void persist_stager(const MulArgs<T> &a, T *xbuf, T *xrs, WinDesc *wd, uint64_t *xfull, uint64_t *xempty) {
    const int lane = threadIdx.x & 31;
    const int32_t nit = persist_item_count(a.nslices);
    uint32_t w = 0;        // window counter
    for (int32_t k = 0; k < nit; ++k) {
        const bsm_slice sl = a.slices[persist_item(k, a.nslices)];
        const int32_t L = sl.r1;
        const SetRef out = set_ref(a, sl.out_set);
        // windows of this item that carry data
        int32_t clast = -1;
        for (int32_t ci = sl.c_begin; ci < sl.c_end; ++ci) {
            const bsm_contrib cb = a.contrib[ci];
            const int32_t jhi = (cb.form & 1) ? min(sl.r1, cb.out_len) : cb.n;
            if (jhi > 0 && cb.m > 0) clast = ci;
        }
        bool first = true;
        for (int32_t ci = sl.c_begin; ci <= clast || (clast < 0 && ci == sl.c_begin); ++ci) {
            // an item without data still needs ONE (empty) window so that its rows are written
            bsm_contrib cb;
            bool empty_item = clast < 0;
            if (!empty_item) cb = a.contrib[ci];
            const bool tform = !empty_item && (cb.form & 1) != 0;
            const int32_t m = empty_item ? 0 : cb.m;
            const int32_t jhi = empty_item ? 0 : (tform ? min(sl.r1, cb.out_len) : cb.n);
            if (!empty_item && (jhi <= 0 || m == 0)) continue;
            const SetRef in = empty_item ? SetRef{0, nullptr} : set_ref(a, cb.in_set);
            const int32_t wstep = (tform || empty_item) ? max(jhi, 1) : kQWin;
            for (int32_t jw = 0; jw < max(jhi, 1); jw += wstep, ++w) {
                const int32_t wend = min(jhi, jw + wstep);
                const uint32_t b = w & 1;
                if (w >= 2) mbar_wait(&xempty[b], ((w >> 1) - 1) & 1);
                T *xs = xbuf + b * kQWin;
                if (tform)          // x at the block's rows, zero-padded to the rows the lanes hold
                    stage_gather<T>(a, in, 0, m, kFMaxRows, xs, lane);
                else if (wend > jw)
                    stage_gather<T>(a, in, jw, wend - jw, wend - jw, xs, lane);
                if (first)          // x at the segment's own rows (read by the fused transposed partials of this item)
                    stage_gather<T>(a, out, 0, L, kFMaxRows, xrs + (k & 1) * kFMaxRows, lane);
                if (lane == 0) {
                    WinDesc d;
                    d.m = m;
                    d.jw = jw;
                    d.wcols = wend - jw;
                    const bool last = (empty_item || ci == clast) && wend >= jhi;
                    d.flags = (tform ? 1 : 0) | ((!empty_item && (cb.form & 2)) ? 2 : 0) | (first ? 4 : 0) | (last ? 8 : 0) |
                              ((sl.flags & 1) ? 16 : 0) | ((last && k == nit - 1) ? 32 : 0);
                    d.toff = (!empty_item && (cb.form & 2)) ? a.contrib_toff[ci] : 0;
                    d.L = L;
                    d.out_start = out.start;
                    d.out_pool = (int64_t)(out.pool - a.pool);
                    d.scratch_off = sl.scratch_off;
                    d.pad = 0;
                    wd[b] = d;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xfull[b]);        // release: buffer + descriptor are visible to the waiters
                first = false;
            }
            if (empty_item) break;
        }
    }
}

struct PersistCursor {      // what a consumer carries from item to item
    uint32_t q, pbase, w, item;
};

// One work item on the consumer side, compiled per rows-per-lane (RPL = 2 / 4 / 8 for segments of <= 64 / 128 / 256 rows):
// the row sums live only inside this function, so every instantiation gets the register allocation of the round-1
// per-CTA consumer (a single body with a run-time RPL kept eight row sums live everywhere and spilled in the hot loop:
// C2 1.97 -> 2.8 ms, measured). `d` is the item's first window descriptor, already waited for. Returns true when the
// CTA's work is finished.
template <class T, int RPL, bool CONJ>
__device__ __forceinline__ bool persist_item_body(const MulArgs<T> &a, unsigned char *stages, const T *xbuf, const T *xrs_all,
                                                  T *accT, const WinDesc *wd, uint64_t *full, uint64_t *empty, uint64_t *xfull,
                                                  uint64_t *xempty, WinDesc d, PersistCursor &cur) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    T accN[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) accN[k] = El<T>::zero();
    const T *xrs = xrs_all + (cur.item & 1) * kFMaxRows;
    for (;;) {
        const uint32_t b = cur.w & 1;
        const T *xs = xbuf + b * kQWin;
        const int32_t m = d.m;
        if (m > 0 && d.wcols > 0) {
            const int32_t cc = chunk_cols<T>(m);
            const bool tform = (d.flags & 1) != 0, fusedT = (d.flags & 2) != 0;
            T *tg = fusedT ? a.scratch + d.toff : nullptr;
            for (int32_t j0 = 0; j0 < d.wcols; j0 += cc, ++cur.q) {
                const int32_t ncols = min(cc, d.wcols - j0);
                const uint32_t stage = cur.q % kPStages;
                const uint32_t delta = (uint32_t)(((int64_t)(d.jw + j0) * m * (int64_t)sizeof(T)) & 15);
                mbar_wait(&full[stage], (cur.q / kPStages) & 1);
                const T *sm = reinterpret_cast<const T *>(stages + stage * kPStageBytes + delta);
                const int wrot = (warp - cur.pbase) & (kFWarps - 1);
                if (tform)
                    consume_chunk<T, RPL, CONJ>(sm, m, ncols, lane, wrot, false, true, nullptr, xs, accN, nullptr, accT + j0);
                else
                    consume_chunk<T, RPL, CONJ>(sm, m, ncols, lane, wrot, true, fusedT, xs + j0, xrs, accN,
                                                fusedT ? tg + d.jw + j0 : nullptr, nullptr);
                cur.pbase += (uint32_t)((ncols + 1) >> 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&xempty[b]);
        ++cur.w;
        if (d.flags & 8) break;
        mbar_wait(&xfull[cur.w & 1], (cur.w >> 1) & 1);
        d = wd[cur.w & 1];
    }
    // item complete: the eight warps add their row sums into accT in warp order, then the rows are written
    consumer_bar();
    for (int wi = 0; wi < kFWarps; ++wi) {
        if (warp == wi) {
#pragma unroll
            for (int k = 0; k < RPL; ++k) accT[k * 32 + lane] = El<T>::add(accT[k * 32 + lane], accN[k]);
        }
        consumer_bar();
    }
    if (t < d.L) {
        const T tot = accT[t];
        if (d.flags & 16) {
            const int32_t row = d.out_start >= 0 ? d.out_start + t : __ldg(a.pool + d.out_pool + t);
            T v = El<T>::mul(a.alpha, tot);
            if (!a.beta_false) v = El<T>::add(v, El<T>::mul(a.beta, a.y[row]));
            a.y[row] = v;
        } else {
            a.scratch[d.scratch_off + t] = tot;
        }
    }
    accT[t] = El<T>::zero();
    consumer_bar();
    ++cur.item;
    return (d.flags & 32) != 0;
}

template <class T, bool CONJ>
__device__ __forceinline__ void persist_consumer(const MulArgs<T> &a, unsigned char *stages, const T *xbuf, const T *xrs_all,
                                                 T *accT, const WinDesc *wd, uint64_t *full, uint64_t *empty, uint64_t *xfull,
                                                 uint64_t *xempty) {
    if (persist_item_count(a.nslices) == 0) return;
    PersistCursor cur{0, 0, 0, 0};
    accT[threadIdx.x] = El<T>::zero();
    consumer_bar();
    for (;;) {
        mbar_wait(&xfull[cur.w & 1], (cur.w >> 1) & 1);
        const WinDesc d = wd[cur.w & 1];
        bool end;
        if (d.L <= 64)
            end = persist_item_body<T, 2, CONJ>(a, stages, xbuf, xrs_all, accT, wd, full, empty, xfull, xempty, d, cur);
        else if (d.L <= 128)
            end = persist_item_body<T, 4, CONJ>(a, stages, xbuf, xrs_all, accT, wd, full, empty, xfull, xempty, d, cur);
        else
            end = persist_item_body<T, 8, CONJ>(a, stages, xbuf, xrs_all, accT, wd, full, empty, xfull, xempty, d, cur);
        if (end) break;
    }
}

template <class T>
__global__ void __launch_bounds__(kQThreads, kQMaxCtasPerSm) sym_persist_kernel(const MulArgs<T> a) {
    extern __shared__ __align__(128) unsigned char qsm[];
    using SM = PersistSmem<T>;
    unsigned char *stages = qsm;
    T *xbuf = reinterpret_cast<T *>(qsm + SM::ring);
    T *xrs = reinterpret_cast<T *>(qsm + SM::ring + SM::xbuf);
    T *accT = reinterpret_cast<T *>(qsm + SM::ring + SM::xbuf + SM::xrs);
    WinDesc *wd = reinterpret_cast<WinDesc *>(qsm + SM::ring + SM::xbuf + SM::xrs + SM::acct);
    uint64_t *full = reinterpret_cast<uint64_t *>(qsm + SM::ring + SM::xbuf + SM::xrs + SM::acct + SM::wdesc);
    uint64_t *empty = full + kPStages;
    uint64_t *xfull = empty + kPStages;
    uint64_t *xempty = xfull + 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kFWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&xfull[i], 1);
            mbar_init(&xempty[i], kFWarps);
        }
        mbar_fence_init();
        if (a.x.npeer) peer_entry(a.x.sync);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp == kFWarps) {
        if (threadIdx.x == kFThreads) persist_issuer<T>(a, stages, full, empty);
    } else if (warp == kFWarps + 1) {
        persist_stager<T>(a, xbuf, xrs, wd, xfull, xempty);
    } else {
        bool done = false;
        if constexpr (sizeof(T) == 16) {      // conj is the identity for real element types
            if (a.conj) {
                persist_consumer<T, true>(a, stages, xbuf, xrs, accT, wd, full, empty, xfull, xempty);
                done = true;
            }
        }
        if (!done) persist_consumer<T, false>(a, stages, xbuf, xrs, accT, wd, full, empty, xfull, xempty);
        // every x read of this CTA (the stager's) precedes the last window the consumers waited for
        if (a.x.npeer && threadIdx.x == 0) peer_exit(a.x.sync);
    }
}

}  // namespace bsm
