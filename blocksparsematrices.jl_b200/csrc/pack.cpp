// pack.cpp — host packer: index-set deduplication, arena layout, plan construction (slicing, ownership, static
// shared-memory schedules of the warp-stream kernel, local/remote phases of slab handles, colour-ordered plans).
// Everything here is deterministic (no hashing order leaks into the tables): the same input gives bit-identical
// tables with or without a device, which is what the CPU tests replay with the NumPy plan interpreter.
#include "plan.h"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace bsm {

// ---------------------------------------------------------------------------- index sets
static uint64_t fnv1a(const int32_t *p, int64_t n) {
    uint64_t h = 1469598103934665603ull;
    for (int64_t i = 0; i < n; ++i) {
        h ^= (uint32_t)p[i];
        h *= 1099511628211ull;
    }
    return h ^ (uint64_t)n;
}

int32_t IndexSets::add_range(int64_t start0, int64_t n) {
    // key space of ranges: hash of (start, len) with a tag so it cannot collide silently —
    // equality is always re-checked.
    uint64_t key = ((uint64_t)start0 * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)n << 1) ^ 1ull;
    auto &cands = by_hash[key];
    for (int32_t id : cands)
        if (start[id] == (int32_t)start0 && len[id] == (int32_t)n) return id;
    int32_t id = (int32_t)len.size();
    len.push_back((int32_t)n);
    start.push_back((int32_t)start0);
    pool_off.push_back(0);
    cands.push_back(id);
    return id;
}

int32_t IndexSets::add_vector(const int64_t *idx1, int64_t n, int64_t limit) {
    bool contiguous = n > 0;
    for (int64_t k = 0; k < n; ++k) {
        if (idx1[k] < 1 || idx1[k] > limit) return -1;
        if (idx1[k] != idx1[0] + k) contiguous = false;
    }
    if (contiguous) return add_range(idx1[0] - 1, n);
    std::vector<int32_t> tmp((size_t)n);
    for (int64_t k = 0; k < n; ++k) tmp[(size_t)k] = (int32_t)(idx1[k] - 1);
    uint64_t key = fnv1a(tmp.data(), n) << 1;  // even keys: vectors, odd keys: ranges
    auto &cands = by_hash[key];
    for (int32_t id : cands) {
        if (start[id] < 0 && len[id] == (int32_t)n &&
            (n == 0 || std::memcmp(pool.data() + pool_off[id], tmp.data(), (size_t)n * 4) == 0))
            return id;
    }
    int32_t id = (int32_t)len.size();
    len.push_back((int32_t)n);
    start.push_back(-1);
    pool_off.push_back((int64_t)pool.size());
    pool.insert(pool.end(), tmp.begin(), tmp.end());
    cands.push_back(id);
    return id;
}

int32_t IndexSets::add_subset(int32_t set, int64_t k0, int64_t n) {
    if (k0 == 0 && n == len[set]) return set;
    if (start[set] >= 0) return add_range((int64_t)start[set] + k0, n);
    const int32_t id = (int32_t)len.size();
    len.push_back((int32_t)n);
    start.push_back(-1);
    pool_off.push_back(pool_off[set] + k0);
    return id;
}

// ---------------------------------------------------------------------------- arena
void layout_arena(HostMatrix &M) {
    const int64_t s = dtype_size(M.dtype);
    const int64_t align = kArenaAlignBytes / s;
    M.block_off.resize(M.blocks.size());
    int64_t off = 0, stored = 0;
    for (size_t b = 0; b < M.blocks.size(); ++b) {
        M.block_off[b] = off;
        const int64_t e = (int64_t)M.blocks[b].m * M.blocks[b].n;
        stored += e;
        off += (e + align - 1) / align * align;
    }
    M.stored = stored;
    M.arena_elems = off + align;  // tail slack: vector loads may over-read up to 15 bytes
}

// ---------------------------------------------------------------------------- plan
// Columns per chunk of a warp-stream contribution (m x n block, element size s). The per-chunk instruction cost of
// stream_warp_kernel is fixed (descriptor, barrier, issue: ~1 us of a warp's time — measured on C3: time = a + b * chunks
// with b = 0.38 ns (op N) / 0.63 ns (op T) per chunk), so chunks are as large as the ring allows with TWO of them
// resident (one in flight while one is consumed): payload + its x values <= half the ring. T-form chunks prefer a
// multiple of 8 columns (whole tensor-core tiles).
static int64_t wchunk_cols(int64_t m, int64_t n, bool tform, int64_t s, int64_t tuned_bytes) {
    const int64_t colbytes = m * s;
    int64_t cmax;
    if (tuned_bytes > 0) {
        cmax = tuned_bytes / colbytes;
    } else {
        const int64_t half = kWRingBytes / 2 - 64;       // 16-byte alignment shifts and round-ups of both parts
        cmax = tform ? (half - ((m * s + 15) & ~(int64_t)15)) / colbytes : half / (colbytes + s);
    }
    cmax = std::max<int64_t>(1, std::min<int64_t>(cmax, kWMaxCols));
    const int64_t nch = (n + cmax - 1) / cmax;
    int64_t cc = (n + nch - 1) / nch;
    const int64_t c8 = (cc + 7) / 8 * 8, c4 = (cc + 3) / 4 * 4;
    // rounding UP to whole tiles never adds a chunk; when it would overflow the budget the even split stays as it is
    // (rounding down would: 36 columns with room for 18 per chunk are 18 + 18, not 16 + 16 + 4)
    if (tform && c8 <= cmax)
        cc = c8;
    else if (c4 <= cmax)
        cc = c4;
    return std::max<int64_t>(cc, 1);
}

std::string build_plan(HostMatrix &M, const std::vector<ContribIR> &ir, int64_t out_dim,
                       int64_t in_dim, const PlanParams &pp, HostPlan &P) {
    IndexSets &S = M.sets;
    const int64_t s = dtype_size(M.dtype);
    const int V = dtype_vec(M.dtype);
    const int64_t own_lo = pp.own_hi < 0 ? 0 : pp.own_lo;
    const int64_t own_hi = pp.own_hi < 0 ? out_dim : pp.own_hi;
    P = HostPlan();
    P.out_dim = out_dim;
    P.in_dim = in_dim;

    // 1. groups in order of first appearance of their output set; a group is dropped when neither
    //    its own rows nor the rows of a fused transposed partial of one of its blocks are owned
    std::vector<int8_t> owned_cache(S.len.size(), -1);
    auto has_owned = [&](int32_t set) -> bool {
        if (owned_cache[set] < 0) {
            bool any = false;
            for (int64_t k = 0; k < S.len[set] && !any; ++k) {
                const int64_t r = S.at(set, k);
                any = (r >= own_lo && r < own_hi);
            }
            owned_cache[set] = any ? 1 : 0;
        }
        return owned_cache[set] == 1;
    };
    std::vector<int32_t> group_of_set(S.len.size(), -1);
    std::vector<uint8_t> set_needed(S.len.size(), 0);
    for (size_t c = 0; c < ir.size(); ++c) {
        const int32_t os = ir[c].out_set;
        if (os < 0 || (size_t)os >= S.len.size()) return "bad output set";
        if (has_owned(os) || (ir[c].fuse_tset >= 0 && has_owned(ir[c].fuse_tset))) set_needed[os] = 1;
    }
    std::vector<int32_t> gset;
    std::vector<std::vector<int32_t>> members;
    for (size_t c = 0; c < ir.size(); ++c) {
        const int32_t os = ir[c].out_set;
        if (!set_needed[os]) continue;
        if (group_of_set[os] == -1) {
            group_of_set[os] = (int32_t)gset.size();
            gset.push_back(os);
            members.emplace_back();
        }
        members[group_of_set[os]].push_back((int32_t)c);
    }
    const size_t G = gset.size();

    // 2. contributions grouped (stable), CSR pointer = block-row pointer / transposed index
    std::vector<int32_t> contrib_tset;
    P.group_ptr.assign(G + 1, 0);
    P.group_set.assign(gset.begin(), gset.end());
    for (size_t g = 0; g < G; ++g) {
        for (int32_t c : members[g]) {
            const ContribIR &ci = ir[c];
            const BlockSrc &b = M.blocks[ci.block];
            bsm_contrib d;
            d.off = M.block_off[ci.block];
            d.m = b.m;
            d.n = b.n;
            d.in_set = ci.in_set;
            d.form = ci.form | (ci.fuse_tset >= 0 ? kFormFusedT : 0);
            d.out_len = ci.out_len;
            d.block = ci.block;
            if (ci.fuse_tset >= 0) {
                if (ci.form != 0 || !pp.fused || S.len[gset[g]] > kFusedMaxRows || S.len[ci.fuse_tset] < b.n)
                    return "bad fused contribution";
            }
            if (ci.out_len > S.len[gset[g]]) return "contribution longer than its output segment";
            if (S.len[ci.in_set] < (ci.form == 0 ? b.n : b.m)) return "input set shorter than block";
            P.applied_entries += (int64_t)b.m * b.n * (ci.fuse_tset >= 0 ? 2 : 1);
            // a wide N-form block of a CTA-kernel segment that alone exceeds the work-item budget is entered as
            // several column ranges (own input sub-set, own transposed partial): the slicer below can then
            // hand them to different CTAs
            const int64_t bytes = (int64_t)b.m * b.n * s;
            if (pp.fused && pp.split_bytes > 0 && ci.form == 0 && S.len[gset[g]] <= kFusedMaxRows &&
                S.len[gset[g]] > kWarpMaxRows && bytes > pp.split_bytes + pp.split_bytes / 2 && b.n >= 8) {
                const int64_t k = (bytes + pp.split_bytes - 1) / pp.split_bytes;
                const int64_t cn = std::max<int64_t>(4, ((b.n + k - 1) / k + 3) / 4 * 4);
                for (int64_t c0 = 0; c0 < b.n; c0 += cn) {
                    const int64_t w = std::min<int64_t>(cn, b.n - c0);
                    bsm_contrib e = d;
                    e.off = d.off + c0 * b.m;
                    e.n = (int32_t)w;
                    e.in_set = S.add_subset(ci.in_set, c0, w);
                    P.contrib.push_back(e);
                    contrib_tset.push_back(ci.fuse_tset >= 0 ? S.add_subset(ci.fuse_tset, c0, w) : -1);
                }
                continue;
            }
            contrib_tset.push_back(ci.fuse_tset);
            P.contrib.push_back(d);
        }
        P.group_ptr[g + 1] = (int64_t)P.contrib.size();
    }

    // 3. ownership: first come claims; a group is direct iff none of its rows is claimed, it has no
    //    repeated row, and it lies entirely inside the owned range
    std::vector<uint8_t> claimed((size_t)out_dim, 0);
    std::vector<int32_t> stamp((size_t)out_dim, -1);
    P.group_direct.assign(G, 0);
    for (size_t g = 0; g < G; ++g) {
        const int32_t os = gset[g];
        bool direct = true;
        for (int64_t k = 0; k < S.len[os]; ++k) {
            const int64_t r = S.at(os, k);
            if (r < own_lo || r >= own_hi || claimed[(size_t)r] || stamp[(size_t)r] == (int32_t)g) {
                direct = false;
                break;
            }
            stamp[(size_t)r] = (int32_t)g;
        }
        if (direct)
            for (int64_t k = 0; k < S.len[os]; ++k) claimed[(size_t)S.at(os, k)] = 1;
        P.group_direct[g] = direct;
    }

    // 4. slices: cut every output segment into pieces of <= kMaxSliceHeight outputs, more pieces if
    //    the segment is heavy (work_target_bytes), never thinner than 128 bytes of a column
    struct Tmp {
        bsm_slice s;
        int64_t work;
        int64_t order;
        bool remote = false;
    };
    std::vector<Tmp> tmp;
    std::vector<uint8_t> group_warp(G, 0);
    const int64_t hmin = std::max<int64_t>(1, 128 / s);
    // whether group g is streamed by the warp-stream kernel (whole segment of <= kWarpMaxRows rows, short blocks)
    auto warp_eligible = [&](size_t g) -> bool {
        const int64_t L = S.len[gset[g]];
        if (!(pp.fused && pp.warp_stream) || L == 0 || L > kWarpMaxRows || S.pool.size() >= (size_t)0x7fffffff) return false;
        int64_t entries = 0;
        for (int64_t c = P.group_ptr[g]; c < P.group_ptr[g + 1]; ++c) {
            const bsm_contrib &cb = P.contrib[c];
            entries += (int64_t)cb.m * cb.n;
            if (contrib_tset[(size_t)c] >= 0 || cb.m > kWarpMaxRows) return false;
        }
        return entries > 0;
    };
    // small (L2-resident, latency-bound) problems with few segments: CTA-part mode instead of per-block work items
    // with partial sums through scratch + a second launch
    if (pp.wsplit_bytes > 0 && pp.wcta && pp.in_hi < 0) {
        int64_t nw = 0;
        for (size_t g = 0; g < G; ++g) nw += warp_eligible(g) ? 1 : 0;
        P.wcta = nw > 0 && nw <= kWCtaMaxSegments;
    }
    for (size_t g = 0; g < G; ++g) {
        const int64_t L = S.len[gset[g]];
        if (L == 0) continue;
        int64_t W = 0;
        bool vec_ok = true;
        for (int64_t c = P.group_ptr[g]; c < P.group_ptr[g + 1]; ++c) {
            W += (int64_t)P.contrib[c].m * P.contrib[c].n * s;
            if (P.contrib[c].m % V != 0) vec_ok = false;
        }
        if (L % V != 0) vec_ok = false;
        // stream plans: classify the segment for the TMA-staged kernels
        bool any_fuse = false, all_t = true, t_ok = true, small_ok = true;
        int64_t entries = 0;
        for (int64_t c = P.group_ptr[g]; c < P.group_ptr[g + 1]; ++c) {
            const bsm_contrib &cb = P.contrib[c];
            entries += (int64_t)cb.m * cb.n;
            if (contrib_tset[(size_t)c] >= 0) any_fuse = true;
            if (cb.form & kFormT) {
                if (cb.m > kFusedMaxTRows) t_ok = false;
            } else {
                all_t = false;
            }
            if (cb.m > kWarpMaxRows) small_ok = false;
        }
        if (pp.fused && pp.warp_stream && L <= kWarpMaxRows && small_ok && !any_fuse && entries > 0 &&
            S.pool.size() < (size_t)0x7fffffff) {
            // whole segment streamed by ONE warp of stream_warp_kernel. Small problems (wsplit_bytes > 0) cut
            // a segment with several blocks into several work items: the first stays direct, the others
            // deliver partial vectors through the gather lists
            group_warp[g] = 1;
            const int64_t budget = (pp.wsplit_bytes > 0 && !P.wcta) ? pp.wsplit_bytes : (int64_t)1 << 62;
            int32_t cb0 = (int32_t)P.group_ptr[g];
            const int32_t cend = (int32_t)P.group_ptr[g + 1];
            bool first_item = true;
            while (cb0 < cend) {
                int32_t ce = cb0;
                int64_t wk = 0;
                while (ce < cend) {
                    const int64_t add = (int64_t)P.contrib[ce].m * P.contrib[ce].n * s;
                    if (wk > 0 && add > 0 && wk + add > budget + budget / 2 && W > budget + budget / 2) break;
                    wk += add;
                    ++ce;
                }
                // trailing empty blocks stay with the last item that carries data
                int64_t rest = 0;
                for (int32_t c = ce; c < cend; ++c) rest += (int64_t)P.contrib[c].m * P.contrib[c].n;
                if (rest == 0) ce = cend;
                Tmp t;
                t.s.out_set = gset[g];
                t.s.r0 = 0;
                t.s.r1 = (int32_t)L;
                t.s.c_begin = cb0;
                t.s.c_end = ce;
                t.s.flags = ((P.group_direct[g] && first_item) ? kSliceDirect : 0) | kSliceWarp;
                t.s.scratch_off = 0;
                t.work = wk;
                t.order = (int64_t)tmp.size();
                tmp.push_back(t);
                first_item = false;
                cb0 = ce;
            }
            continue;
        }
        if (pp.fused && L <= kFusedMaxRows && t_ok) {
            // whole segment in one CTA (all its rows in registers). A segment far heavier than the work-item
            // budget is cut along its contribution list: the first item stays direct, the others deliver
            // partial vectors through the gather lists (deterministic, fixed order)
            const int64_t budget = pp.split_bytes > 0 ? pp.split_bytes : (int64_t)1 << 62;
            int32_t cb0 = (int32_t)P.group_ptr[g];
            const int32_t cend = (int32_t)P.group_ptr[g + 1];
            bool first_item = true;
            while (cb0 < cend || first_item) {
                int32_t ce = cb0;
                int64_t wk = 0;
                while (ce < cend && (ce == cb0 || wk + (int64_t)P.contrib[ce].m * P.contrib[ce].n * s <= budget + budget / 2 ||
                                     W <= budget + budget / 2)) {
                    wk += (int64_t)P.contrib[ce].m * P.contrib[ce].n * s;
                    ++ce;
                }
                Tmp t;
                t.s.out_set = gset[g];
                t.s.r0 = 0;
                t.s.r1 = (int32_t)L;
                t.s.c_begin = cb0;
                t.s.c_end = ce;
                t.s.flags = ((P.group_direct[g] && first_item) ? kSliceDirect : 0) | kSliceFused;
                t.s.scratch_off = 0;
                t.work = wk;
                t.order = (int64_t)tmp.size();
                tmp.push_back(t);
                first_item = false;
                cb0 = ce;
            }
            continue;
        }
        if (pp.fused && all_t && t_ok && !any_fuse) {
            // long segment fed by T-form contributions only: a sub-range of outputs is a run of whole
            // (contiguous) block columns, so the CTA kernel streams it like a short segment
            int64_t pieces = (L + kFusedMaxRows - 1) / kFusedMaxRows;
            pieces = std::max(pieces, std::min((W + pp.work_target_bytes - 1) / pp.work_target_bytes,
                                               std::max<int64_t>(1, L / 16)));
            int64_t ps = (L + pieces - 1) / pieces;
            ps = std::min<int64_t>((ps + 3) / 4 * 4, kFusedMaxRows);
            for (int64_t r0 = 0; r0 < L; r0 += ps) {
                Tmp t;
                t.s.out_set = gset[g];
                t.s.r0 = (int32_t)r0;
                t.s.r1 = (int32_t)std::min(L, r0 + ps);
                t.s.c_begin = (int32_t)P.group_ptr[g];
                t.s.c_end = (int32_t)P.group_ptr[g + 1];
                t.s.flags = (P.group_direct[g] ? kSliceDirect : 0) | kSliceFused;
                t.s.scratch_off = 0;
                t.work = W * (t.s.r1 - t.s.r0) / L;
                t.order = (int64_t)tmp.size();
                tmp.push_back(t);
            }
            continue;
        }
        {
            // long segment fed by N-form contributions only (1024-row blocks, op N): 256-row pieces through the CTA
            // kernel — the producer copies the piece of every column with its own bulk copy, which needs every column of
            // every block to start 16-byte aligned (m * s % 16 == 0; pieces start at multiples of 256 rows)
            bool all_n = true, n_ok = true;
            for (int64_t c = P.group_ptr[g]; c < P.group_ptr[g + 1]; ++c) {
                if (P.contrib[c].form & kFormT) all_n = false;
                if (((int64_t)P.contrib[c].m * s) % 16 != 0) n_ok = false;
            }
            if (pp.fused && L > kFusedMaxRows && all_n && n_ok && !any_fuse && entries > 0) {
                // a piece far heavier than the work-item budget is cut along the contribution list as well (the first
                // item stays direct, the others deliver partial vectors through the gather lists)
                const int64_t budget = pp.split_bytes > 0 ? pp.split_bytes : (int64_t)1 << 62;
                for (int64_t r0 = 0; r0 < L; r0 += kFusedMaxRows) {
                    const int64_t r1 = std::min<int64_t>(L, r0 + kFusedMaxRows);
                    int32_t cb0 = (int32_t)P.group_ptr[g];
                    const int32_t cend = (int32_t)P.group_ptr[g + 1];
                    bool first_item = true;
                    while (cb0 < cend) {
                        int32_t ce = cb0;
                        int64_t wk = 0;
                        while (ce < cend) {
                            const int64_t wc = (int64_t)P.contrib[ce].n * (r1 - r0) * s;
                            if (ce > cb0 && wk + wc > budget + budget / 2) break;
                            wk += wc;
                            ++ce;
                        }
                        Tmp t;
                        t.s.out_set = gset[g];
                        t.s.r0 = (int32_t)r0;
                        t.s.r1 = (int32_t)r1;
                        t.s.c_begin = cb0;
                        t.s.c_end = ce;
                        t.s.flags = ((P.group_direct[g] && first_item) ? kSliceDirect : 0) | kSliceFused;
                        t.s.scratch_off = 0;
                        t.work = wk;
                        t.order = (int64_t)tmp.size();
                        tmp.push_back(t);
                        first_item = false;
                        cb0 = ce;
                    }
                }
                continue;
            }
        }
        int64_t pieces = (L + kMaxSliceHeight - 1) / kMaxSliceHeight;
        const int64_t by_work = (W + pp.work_target_bytes - 1) / pp.work_target_bytes;
        const int64_t max_pieces = std::max<int64_t>(1, L / hmin);
        pieces = std::max(pieces, std::min(by_work, max_pieces));
        int64_t ps = (L + pieces - 1) / pieces;
        ps = (ps + V - 1) / V * V;
        ps = std::min<int64_t>(ps, kMaxSliceHeight);
        for (int64_t r0 = 0; r0 < L; r0 += ps) {
            Tmp t;
            t.s.out_set = gset[g];
            t.s.r0 = (int32_t)r0;
            t.s.r1 = (int32_t)std::min(L, r0 + ps);
            t.s.c_begin = (int32_t)P.group_ptr[g];
            t.s.c_end = (int32_t)P.group_ptr[g + 1];
            t.s.flags = (P.group_direct[g] ? kSliceDirect : 0) | (vec_ok ? kSliceVecOk : 0);
            t.s.scratch_off = 0;
            t.work = W * (t.s.r1 - t.s.r0) / L;
            t.order = (int64_t)tmp.size();
            tmp.push_back(t);
        }
    }
    // scratch offsets are assigned in creation order (before the scheduling sort)
    int64_t scratch = 0;
    for (auto &t : tmp) {
        if (!(t.s.flags & kSliceDirect)) {
            t.s.scratch_off = scratch;
            scratch += t.s.r1 - t.s.r0;
        }
    }
    // fused transposed partials: one vector of n entries per fused contribution
    P.contrib_toff.assign(P.contrib.size(), -1);
    for (size_t c = 0; c < P.contrib.size(); ++c) {
        if (contrib_tset[c] >= 0) {
            P.contrib_toff[c] = scratch;
            scratch += P.contrib[c].n;
        }
    }
    P.scratch_elems = scratch;

    // 5. gather lists: every owned row that is not claimed by a direct group, or that receives
    //    partial sums; partials are listed in slice creation order (fixed reduction order)
    std::vector<int64_t> cnt((size_t)out_dim + 1, 0);
    for (const auto &t : tmp) {
        if (t.s.flags & kSliceDirect) continue;
        for (int64_t k = t.s.r0; k < t.s.r1; ++k) {
            const int64_t r = S.at(t.s.out_set, k);
            if (r >= own_lo && r < own_hi) cnt[(size_t)r + 1]++;
        }
    }
    for (size_t c = 0; c < P.contrib.size(); ++c) {
        if (contrib_tset[c] < 0) continue;
        for (int64_t k = 0; k < P.contrib[c].n; ++k) {
            const int64_t r = S.at(contrib_tset[c], k);
            if (r >= own_lo && r < own_hi) cnt[(size_t)r + 1]++;
        }
    }
    std::vector<int64_t> rowslot((size_t)out_dim, -1);
    P.gather_ptr.push_back(0);
    for (int64_t r = own_lo; r < own_hi; ++r) {
        if (cnt[(size_t)r + 1] > 0 || !claimed[(size_t)r]) {
            rowslot[(size_t)r] = (int64_t)P.gather_rows.size();
            P.gather_rows.push_back((int32_t)r | (claimed[(size_t)r] ? (int32_t)0x80000000 : 0));
            P.gather_ptr.push_back(P.gather_ptr.back() + cnt[(size_t)r + 1]);
        }
    }
    P.gather_pos.assign((size_t)P.gather_ptr.back(), 0);
    std::vector<int64_t> fill(P.gather_ptr.begin(), P.gather_ptr.end() - 1);
    for (const auto &t : tmp) {
        if (t.s.flags & kSliceDirect) continue;
        for (int64_t k = t.s.r0; k < t.s.r1; ++k) {
            const int64_t r = S.at(t.s.out_set, k);
            if (r >= own_lo && r < own_hi)
                P.gather_pos[(size_t)fill[(size_t)rowslot[(size_t)r]]++] = t.s.scratch_off + (k - t.s.r0);
        }
    }

    for (size_t c = 0; c < P.contrib.size(); ++c) {
        if (contrib_tset[c] < 0) continue;
        for (int64_t k = 0; k < P.contrib[c].n; ++k) {
            const int64_t r = S.at(contrib_tset[c], k);
            if (r >= own_lo && r < own_hi)
                P.gather_pos[(size_t)fill[(size_t)rowslot[(size_t)r]]++] = P.contrib_toff[c] + k;
        }
    }

    // 5b. slab handles: a slice is remote when one of its inputs is outside the x range this rank owns
    //     before the all-gather (the fused transposed partial also reads x at the segment's own rows)
    if (pp.in_hi >= 0) {
        std::vector<int8_t> set_local(S.len.size(), -1);
        auto is_local = [&](int32_t set) -> bool {
            if (set_local[set] < 0) {
                bool ok = true;
                if (S.start[set] >= 0) {
                    ok = S.start[set] >= pp.in_lo && (int64_t)S.start[set] + S.len[set] <= pp.in_hi;
                } else {
                    for (int64_t k = 0; k < S.len[set] && ok; ++k) {
                        const int64_t r = S.at(set, k);
                        ok = r >= pp.in_lo && r < pp.in_hi;
                    }
                }
                set_local[set] = ok ? 1 : 0;
            }
            return set_local[set] == 1;
        };
        for (auto &t : tmp) {
            for (int32_t c = t.s.c_begin; c < t.s.c_end && !t.remote; ++c) {
                if (!is_local(P.contrib[c].in_set)) t.remote = true;
                if (contrib_tset[(size_t)c] >= 0 && !is_local(t.s.out_set)) t.remote = true;
            }
            if (t.remote) {
                t.s.flags |= kSliceRemote;
                P.has_remote = true;
            }
        }
    }

    // 6. schedule: CTA-kernel slices first (heaviest first), then the warp-stream slices in creation
    //    order (= arena order, so every warp streams a contiguous run of HBM), then the gather slices
    //    (heaviest first); inside every class the rank-local slices come before the remote ones; stable
    auto cls = [](const Tmp &t) { return (t.s.flags & kSliceFused) ? 0 : (t.s.flags & kSliceWarp) ? 1 : 2; };
    std::stable_sort(tmp.begin(), tmp.end(), [&](const Tmp &a, const Tmp &b) {
        const int ca = cls(a), cb = cls(b);
        if (ca != cb) return ca < cb;
        if (a.remote != b.remote) return !a.remote;
        if (ca == 1) return false;
        return a.work > b.work;
    });
    // 6b. single-wave regime of the warp-stream kernel (the warp slices of this handle stream less than ~768 MB, e.g. a
    //     slab of C3 on 8 GPUs): ONE work item per resident warp slot. A slot's share is then only a few segments, and
    //     cutting the segment list into contiguous runs leaves the heaviest item at 1.7 x the mean (measured: 58 us against
    //     37 us at the roofline) — so the segments are PACKED into the items instead: longest processing time first onto
    //     the lightest item (cost = bytes + 1 KB per chunk, the chunk's fixed instruction cost). The items of the local
    //     and of the remote group are packed separately (the local ones run while x is gathered on the NCCL path).
    std::vector<int64_t> witem_cut_after;      // positions (in the final slice order) after which a work item ends
    if (!P.wcta && pp.witem_bytes <= 0) {
        size_t w0 = 0;
        while (w0 < tmp.size() && cls(tmp[w0]) != 1) ++w0;
        size_t w1 = w0;
        int64_t wbytes = 0;
        while (w1 < tmp.size() && cls(tmp[w1]) == 1) wbytes += tmp[w1++].work;
        const int64_t per_slot = pp.witems_per_slot > 0 ? pp.witems_per_slot : (wbytes < ((int64_t)768 << 20) ? 1 : 3);
        if (w1 > w0 && per_slot == 1) {
            auto cost_of = [&](const Tmp &t) {
                int64_t c = 0;
                for (int32_t k = t.s.c_begin; k < t.s.c_end; ++k) {
                    const int64_t bytes = (int64_t)P.contrib[k].m * P.contrib[k].n * s;
                    if (bytes == 0) continue;
                    const int64_t cc = wchunk_cols(P.contrib[k].m, P.contrib[k].n, (P.contrib[k].form & kFormT) != 0, s, pp.wchunk_bytes);
                    c += bytes + 1024 * ((P.contrib[k].n + cc - 1) / cc);
                }
                return c;
            };
            std::vector<Tmp> packed;
            packed.reserve(w1 - w0);
            const int64_t slots = 148 * 16;
            int64_t cost_grp[2] = {0, 0};
            for (size_t i = w0; i < w1; ++i) cost_grp[tmp[i].remote ? 1 : 0] += cost_of(tmp[i]);
            const int64_t allcost = std::max<int64_t>(cost_grp[0] + cost_grp[1], 1);
            // the slots are shared between the groups in proportion to their cost; together never more than one wave
            int64_t bins_grp[2];
            bins_grp[0] = cost_grp[1] == 0 ? slots : std::max<int64_t>(cost_grp[0] > 0 ? 1 : 0, slots * cost_grp[0] / allcost);
            bins_grp[1] = cost_grp[1] == 0 ? 0 : std::max<int64_t>(1, slots - bins_grp[0]);
            for (int grp = 0; grp < 2; ++grp) {       // 0: local, 1: remote
                std::vector<size_t> idx;
                for (size_t i = w0; i < w1; ++i)
                    if ((int)tmp[i].remote == grp) idx.push_back(i);
                if (idx.empty()) continue;
                const int64_t nbins = std::max<int64_t>(1, std::min<int64_t>((int64_t)idx.size(), bins_grp[grp]));
                std::vector<size_t> order(idx);
                std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost_of(tmp[a]) > cost_of(tmp[b]); });
                std::vector<int64_t> load((size_t)nbins, 0);
                std::vector<std::vector<size_t>> bins((size_t)nbins);
                // lightest bin first; ties broken by bin index, so the packing is deterministic
                std::vector<std::pair<int64_t, int64_t>> heap;
                for (int64_t b = 0; b < nbins; ++b) heap.emplace_back(0, b);
                auto cmp = [](const std::pair<int64_t, int64_t> &a, const std::pair<int64_t, int64_t> &b) { return a > b; };
                std::make_heap(heap.begin(), heap.end(), cmp);
                for (size_t i : order) {
                    std::pop_heap(heap.begin(), heap.end(), cmp);
                    auto &top = heap.back();
                    bins[(size_t)top.second].push_back(i);
                    top.first += cost_of(tmp[i]);
                    std::push_heap(heap.begin(), heap.end(), cmp);
                }
                for (auto &bin : bins) {
                    if (bin.empty()) continue;
                    std::sort(bin.begin(), bin.end());          // arena order inside an item
                    for (size_t i : bin) packed.push_back(tmp[i]);
                    witem_cut_after.push_back((int64_t)(w0 + packed.size() - 1));
                }
            }
            std::copy(packed.begin(), packed.end(), tmp.begin() + (long)w0);
        }
    }
    P.slices.reserve(tmp.size());
    for (const auto &t : tmp) {
        P.slices.push_back(t.s);
        if (t.s.flags & kSliceFused) {
            const int32_t Ls = t.s.r1 - t.s.r0;
            const int32_t lane_rows = Ls <= 64 ? 64 : Ls <= 128 ? 128 : 256;  // rows held by the lanes (32 * RPL)
            if (t.s.r0 > 0 || t.s.r1 < S.len[t.s.out_set]) P.fused_general = true;
            for (int32_t c = t.s.c_begin; c < t.s.c_end; ++c)
                if ((P.contrib[c].form & kFormT) && P.contrib[c].m > lane_rows) P.fused_general = true;
        }
        if (t.s.flags & kSliceFused) P.n_fused_slices++;
        if (t.s.flags & kSliceWarp) P.n_warp_slices++;
        if (!t.remote) {
            if (t.s.flags & kSliceFused)
                P.n_fused_local++;
            else if (!(t.s.flags & kSliceWarp))
                P.n_gather_local++;
        }
    }

    // 7. chunk stream of the warp slices: every chunk is one bulk copy of whole block columns and
    //    carries the resolved x / y positions, so the consuming warp never chases a table
    if (P.n_warp_slices > 0) {
        int64_t total = 0;
        for (int64_t i = P.n_fused_slices; i < P.n_fused_slices + P.n_warp_slices; ++i)
            for (int32_t c = P.slices[i].c_begin; c < P.slices[i].c_end; ++c)
                total += (int64_t)P.contrib[c].m * P.contrib[c].n * s;
        int64_t target = pp.witem_bytes;
        // items per warp slot (148 SMs x 16 resident warps): three for large problems (tail of the last wave), ONE when a
        // slot streams less than ~100 KB in total — every item start costs a ~3 us pipeline bubble (item pointer ->
        // descriptors -> first chunk), which a 34 KB item (5 us of streaming) cannot amortise (measured on the 240 MB
        // slabs of C3 on 8 GPUs: 60 us against 37 us at the roofline)
        if (target <= 0) {
            const int64_t per_slot = pp.witems_per_slot > 0 ? pp.witems_per_slot : (total < ((int64_t)768 << 20) ? 1 : 3);
            target = std::min<int64_t>(1 << 20, std::max<int64_t>(4 << 10, total / (148 * 16 * per_slot) + 1));
        }
        P.witem_ptr.push_back(0);
        int64_t item_bytes = 0;
        for (int64_t i = P.n_fused_slices; i < P.n_fused_slices + P.n_warp_slices; ++i) {
            const bsm_slice &sl = P.slices[i];
            const int64_t L = S.len[sl.out_set];
            const size_t first = P.wchunk.size();
            for (int32_t c = sl.c_begin; c < sl.c_end; ++c) {
                const bsm_contrib &cb = P.contrib[c];
                if (cb.m == 0 || cb.n == 0) continue;
                const bool tf = (cb.form & kFormT) != 0;
                const int64_t colbytes = (int64_t)cb.m * s;
                const int64_t cc = wchunk_cols(cb.m, cb.n, tf, s, pp.wchunk_bytes);
                for (int64_t j0 = 0; j0 < cb.n; j0 += cc) {
                    const int64_t nc = std::min<int64_t>(cc, cb.n - j0);
                    const int64_t b0 = cb.off * s + j0 * colbytes;      // first byte of the chunk
                    const int64_t a0 = b0 & ~(int64_t)15;
                    const int64_t bytes = ((b0 - a0) + nc * colbytes + 15) & ~(int64_t)15;
                    bsm_wchunk w;
                    std::memset(&w, 0, sizeof(w));
                    w.src16 = (uint32_t)((a0 >> 4) & 0xffffffffll);
                    w.src16_hi = (uint8_t)((a0 >> 4) >> 32);
                    w.bytes16 = (uint16_t)(bytes >> 4);
                    w.ncols = (uint16_t)nc;
                    w.m = (uint8_t)cb.m;
                    w.delta = (uint8_t)(b0 - a0);
                    w.seg_len = (uint8_t)L;
                    w.flags = (uint8_t)(tf ? kWcT : 0);
                    const bool xpool = S.start[cb.in_set] < 0;
                    if (xpool) w.flags |= kWcXPool;
                    const int64_t xbase = xpool ? S.pool_off[cb.in_set] : (int64_t)S.start[cb.in_set];
                    w.x_ref = (int32_t)(tf ? xbase : xbase + j0);
                    w.out_col = (uint8_t)(tf ? j0 : 0);
                    P.wchunk.push_back(w);
                    item_bytes += bytes;
                }
            }
            if (P.wchunk.size() == first) return "warp-stream segment without data";
            // output of the segment: the same record closes the segment (or every part of it)
            uint8_t out_flags = 0;
            int64_t out_ref;
            if (sl.flags & kSliceDirect) {
                out_flags |= kWcDirect;
                if (S.start[sl.out_set] < 0) {
                    out_flags |= kWcOutPool;
                    out_ref = S.pool_off[sl.out_set];
                } else {
                    out_ref = S.start[sl.out_set];
                }
            } else {
                out_ref = sl.scratch_off;
            }
            if (P.wcta) {
                // CTA-part mode: the chunk list of the segment is dealt to kWItemsPerCta consecutive work items (the
                // warps of one CTA), balanced by bytes; part 0 is never empty, later parts may be
                const size_t last = P.wchunk.size();
                int64_t seg_bytes = 0;
                for (size_t q = first; q < last; ++q) seg_bytes += (int64_t)P.wchunk[q].bytes16 * 16;
                size_t q = first;
                int64_t done = 0;
                for (int part = 0; part < kWItemsPerCta; ++part) {
                    const size_t q_begin = q;
                    const int64_t goal = seg_bytes * (part + 1) / kWItemsPerCta;
                    if (part == kWItemsPerCta - 1) {
                        q = last;
                    } else {
                        while (q < last && done < goal) {
                            done += (int64_t)P.wchunk[q].bytes16 * 16;
                            ++q;
                        }
                    }
                    if (q > q_begin) {
                        P.wchunk[q_begin].flags |= kWcSegBegin;
                        bsm_wchunk &pe = P.wchunk[q - 1];
                        pe.flags |= kWcSegEnd | kWcCtaPart | out_flags;
                        pe.out = out_ref;
                    }
                    P.witem_ptr.push_back((int32_t)q);
                }
                if (q != last) return "CTA-part split lost chunks";
                if (!(sl.flags & kSliceRemote)) P.n_warp_items_local = (int64_t)P.witem_ptr.size() - 1;
                continue;
            }
            bsm_wchunk &wb = P.wchunk[first];
            bsm_wchunk &we = P.wchunk.back();
            wb.flags |= kWcSegBegin;
            we.flags |= kWcSegEnd | out_flags;
            we.out = out_ref;
            // work items never mix local and remote segments: the local items run while x is gathered
            const bool last_local = !(sl.flags & kSliceRemote) && i + 1 < P.n_fused_slices + P.n_warp_slices &&
                                    (P.slices[i + 1].flags & kSliceRemote);
            bool cut_here = item_bytes >= target || last_local;
            if (!witem_cut_after.empty())      // packed items: the cuts were decided in step 6b
                cut_here = std::binary_search(witem_cut_after.begin(), witem_cut_after.end(), i);
            if (cut_here) {
                P.witem_ptr.push_back((int32_t)P.wchunk.size());
                item_bytes = 0;
            }
            if (!(sl.flags & kSliceRemote)) P.n_warp_items_local = (int64_t)P.witem_ptr.size() - 1;
        }
        if (P.witem_ptr.back() != (int32_t)P.wchunk.size()) P.witem_ptr.push_back((int32_t)P.wchunk.size());
        if (!(P.slices[P.n_fused_slices + P.n_warp_slices - 1].flags & kSliceRemote))
            P.n_warp_items_local = (int64_t)P.witem_ptr.size() - 1;   // no remote warp segment at all
        {
            bool any_n = false, any_t = false;
            for (const bsm_wchunk &w : P.wchunk) ((w.flags & kWcT) ? any_t : any_n) = true;
            P.wform = (any_n && any_t) ? 2 : any_t ? 1 : 0;
        }
        // static shared-memory schedule of every work item: chunks are placed in a circular byte
        // buffer in issue order; a chunk that does not fit waits for the oldest live chunks
        for (size_t it = 0; it + 1 < P.witem_ptr.size(); ++it) {
            const int32_t q0 = P.witem_ptr[it], q1 = P.witem_ptr[it + 1];
            int64_t head = 0;            // next free byte
            int32_t oldest = q0;         // oldest live (issued, not yet consumed) chunk
            for (int32_t q = q0; q < q1; ++q) {
                bsm_wchunk &w = P.wchunk[(size_t)q];
                const int64_t cnt = (w.flags & kWcT) ? w.m : w.ncols;
                // x values behind the chunk: a bulk copy fetches the enclosing 16-byte aligned range (one more group)
                const int64_t foot = (int64_t)w.bytes16 * 16 + ((cnt * s + 15) & ~(int64_t)15) + (s < 16 ? 16 : 0);
                if (foot > kWRingBytes) return "warp-stream chunk larger than the ring";
                for (;;) {
                    // live region: from the oldest live chunk's offset to head (circular)
                    bool fits;
                    if (oldest == q) {
                        head = 0;
                        fits = true;
                    } else {
                        const int64_t tail = (int64_t)P.wchunk[(size_t)oldest].smem16 * 16;
                        if (head > tail)
                            fits = (head + foot <= kWRingBytes) || (foot <= tail);
                        else
                            fits = head + foot <= tail;
                        if (fits && head > tail && head + foot > kWRingBytes) head = 0;
                    }
                    if (fits && q - oldest < kWSlots - 1) break;
                    ++oldest;            // wait for one more chunk to be consumed
                }
                w.smem16 = (uint16_t)(head >> 4);
                w.lag = (uint8_t)(q - oldest);
                head += foot;
            }
        }
    }
    // 8. multi-RHS work items: the SpMM kernel walks slices/contributions itself (its producer warp runs
    //    ahead of the tensor-core consumers), so an item is just a range of slices
    P.spmm_ok = G > 0;
    for (size_t g = 0; g < G; ++g)
        if (S.len[gset[g]] > 0 && !(group_warp[g] && P.group_direct[g])) P.spmm_ok = false;
    if (P.spmm_ok) {
        int64_t total = 0;
        P.spmm_small = true;
        for (const auto &c : P.contrib) {
            total += (int64_t)c.m * c.n * s;
            if (c.m > 32 || ((c.form & kFormT) && c.n > 32)) P.spmm_small = false;
        }
        P.spmm_tma = true;
        for (const auto &c : P.contrib) {
            if (c.m == 0 || c.n == 0) continue;
            if (c.m > 32 || ((c.form & kFormT) && c.n > 32) || S.start[c.in_set] < 0) P.spmm_tma = false;
        }
        for (size_t g = 0; g < G; ++g)
            if (S.len[gset[g]] > 32) P.spmm_tma = false;
        for (size_t g = 0; g < G; ++g) {
            if (S.len[gset[g]] == 0) continue;
            bsm_slice sl;
            sl.out_set = gset[g];
            sl.r0 = 0;
            sl.r1 = S.len[gset[g]];
            sl.c_begin = (int32_t)P.group_ptr[g];
            sl.c_end = (int32_t)P.group_ptr[g + 1];
            sl.flags = kSliceDirect | kSliceWarp;
            sl.scratch_off = 0;
            P.mslices.push_back(sl);
        }
        for (int64_t r = own_lo; r < own_hi; ++r)
            if (!claimed[(size_t)r]) P.muncovered.push_back((int32_t)r);
        const int64_t target = std::max<int64_t>(128 << 10, total / (148 * 2 * 8));
        P.mitem_ptr.push_back(0);
        int64_t acc = 0;
        for (size_t i = 0; i < P.mslices.size(); ++i) {
            for (int32_t c = P.mslices[i].c_begin; c < P.mslices[i].c_end; ++c)
                acc += (int64_t)P.contrib[c].m * P.contrib[c].n * s;
            if (acc >= target) {
                P.mitem_ptr.push_back((int32_t)(i + 1));
                acc = 0;
            }
        }
        if (P.mitem_ptr.back() != (int32_t)P.mslices.size()) P.mitem_ptr.push_back((int32_t)P.mslices.size());
    }
    return std::string();
}

// ---------------------------------------------------------------------------- colour-ordered plan
std::string build_color_plan(const HostMatrix &M, const std::vector<ContribIR> &ir, int64_t out_dim,
                             int64_t in_dim, const PlanParams &pp, HostPlan &P) {
    const IndexSets &S = M.sets;
    const int V = dtype_vec(M.dtype);
    P = HostPlan();
    P.out_dim = out_dim;
    P.in_dim = in_dim;
    if (pp.own_hi >= 0) return std::string();          // slabs use the atomic-free plans
    struct Item {
        int32_t c, sweep, color;
    };
    std::vector<Item> items;
    std::vector<uint64_t> rowmask((size_t)out_dim, 0);
    std::vector<int32_t> stamp((size_t)out_dim, -1);
    int32_t cur_sweep = -1;
    for (size_t c = 0; c < ir.size(); ++c) {
        const ContribIR &ci = ir[c];
        if (ci.sweep != cur_sweep) {                   // a new sweep starts with a clean conflict graph
            if (ci.sweep < cur_sweep) return std::string();   // sweeps must arrive in order
            cur_sweep = ci.sweep;
            std::fill(rowmask.begin(), rowmask.end(), 0);
        }
        uint64_t used = 0;
        for (int64_t k = 0; k < ci.out_len; ++k) {
            const int64_t r = S.at(ci.out_set, k);
            if (stamp[(size_t)r] == (int32_t)c) return std::string();   // repeated index inside one block
            stamp[(size_t)r] = (int32_t)c;
            used |= rowmask[(size_t)r];
        }
        if (~used == 0) return std::string();          // more than 64 colours
        const int32_t color = __builtin_ctzll(~used);
        for (int64_t k = 0; k < ci.out_len; ++k) rowmask[(size_t)S.at(ci.out_set, k)] |= (1ull << color);
        items.push_back(Item{(int32_t)c, ci.sweep, color});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) {
        return a.sweep != b.sweep ? a.sweep < b.sweep : a.color < b.color;
    });
    // contributions stay in IR order (one group each); slices follow the launch order
    P.group_ptr.assign(ir.size() + 1, 0);
    for (size_t c = 0; c < ir.size(); ++c) {
        const ContribIR &ci = ir[c];
        const BlockSrc &b = M.blocks[ci.block];
        bsm_contrib d;
        d.off = M.block_off[ci.block];
        d.m = b.m;
        d.n = b.n;
        d.in_set = ci.in_set;
        d.form = ci.form;
        d.out_len = ci.out_len;
        d.block = ci.block;
        P.contrib.push_back(d);
        P.contrib_toff.push_back(-1);
        P.group_set.push_back(ci.out_set);
        P.group_ptr[c + 1] = (int64_t)c + 1;
        P.applied_entries += (int64_t)b.m * b.n;
    }
    P.group_direct.assign(ir.size(), 1);
    int32_t last_sweep = -1, last_color = -1;
    for (const Item &it : items) {
        const ContribIR &ci = ir[it.c];
        if (it.sweep != last_sweep || it.color != last_color) {
            P.color_ptr.push_back((int32_t)P.slices.size());
            last_sweep = it.sweep;
            last_color = it.color;
        }
        const int64_t L = ci.out_len;
        const bool vec_ok = (P.contrib[it.c].m % V == 0) && (L % V == 0);
        for (int64_t r0 = 0; r0 < L; r0 += kMaxSliceHeight) {
            bsm_slice sl;
            sl.out_set = ci.out_set;
            sl.r0 = (int32_t)r0;
            sl.r1 = (int32_t)std::min<int64_t>(L, r0 + kMaxSliceHeight);
            sl.c_begin = it.c;
            sl.c_end = it.c + 1;
            sl.flags = kSliceDirect | (vec_ok ? kSliceVecOk : 0);
            sl.scratch_off = 0;
            P.slices.push_back(sl);
        }
    }
    P.color_ptr.push_back((int32_t)P.slices.size());
    P.color_ok = true;
    return std::string();
}

}  // namespace bsm
