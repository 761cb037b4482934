// CPU replay of spmm_tma_kernel's shared-memory layout (TEST INFRASTRUCTURE): places a block slab and an X tile in
// byte arrays exactly as the TMA engine does under CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk index XOR 128-byte row
// index mod 8), then walks the consumers' fragment addressing (spmm_layout.h, the very functions the kernel uses),
// applies the DMMA m8n8k4 fragment semantics (a = A[lane/4][lane%4], b = B[lane%4][lane/4], c = C[lane/4][2*(lane%4)+e])
// and the epilogue's column mapping, and compares with a plain matrix product. Also reports the worst bank-conflict
// degree of every fragment load for full 32 x 32 slabs. Prints "OK" or the first mismatch.
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define BSM_HD inline
#include "../blocksparsematrices.jl_b200/csrc/spmm_layout.h"

using namespace bsm;

template <int S> struct Elem;
template <> struct Elem<4> { using T = float; using W = double; };
template <> struct Elem<8> { using T = double; using W = double; };
template <> struct Elem<16> { using T = std::complex<double>; using W = std::complex<double>; };

static double frand() { return (double)rand() / RAND_MAX - 0.5; }
template <class T> static T rnd();
template <> float rnd<float>() { return (float)frand(); }
template <> double rnd<double>() { return frand(); }
template <> std::complex<double> rnd<std::complex<double>>() { return {frand(), frand()}; }

// TMA placement of `bytes` linear bytes (128-byte rows) at a 1 KB-aligned tile base
static void tma_place_linear(std::vector<unsigned char> &smem, size_t base, const unsigned char *src, size_t bytes) {
    for (size_t L = 0; L < bytes; ++L) {
        const size_t row = L >> 7, chunk = (L >> 4) & 7;
        smem[base + (row << 7) + (((chunk ^ (row & 7))) << 4) + (L & 15)] = src[L];
    }
}

// worst conflict degree of one load instruction: lanes are served in phases of 128 bytes
template <int S>
static int conflict_degree(const int (&byteaddr)[32]) {
    const int per_phase = 128 / S;      // lanes per phase: 32 (4 B), 16 (8 B), 8 (16 B)
    int worst = 1;
    for (int p0 = 0; p0 < 32; p0 += per_phase) {
        for (int bank = 0; bank < 32; ++bank) {
            std::vector<int> words;
            for (int l = p0; l < p0 + per_phase; ++l)
                for (int w = 0; w < S / 4; ++w) {
                    const int a = byteaddr[l] + 4 * w;
                    if ((a / 4) % 32 != bank) continue;
                    bool seen = false;
                    for (int q : words) seen |= (q == a);
                    if (!seen) words.push_back(a);
                }
            if ((int)words.size() > worst) worst = (int)words.size();
        }
    }
    return worst;
}

// xs0: first row of X the block multiplies (its alignment decides the shifted fetch of the X tile)
template <int S, int NB>
static int run_case(int m, int n, bool tform, int L, bool conj, int xs0, int *worstA, int *worstB) {
    using T = typename Elem<S>::T;
    using W = typename Elem<S>::W;
    constexpr int NT = NB >= 16 ? 2 : 1, WN = NB >= 16 ? NB / 16 : 1, WM = 4 / WN, MT = (4 + WM - 1) / WM;
    constexpr int KB = 128 / S;
    const int K = tform ? m : n;             // contraction length (T-form: m <= 32)
    const int mo = tform ? n : m;            // outputs of the block
    constexpr int XBoxes = S == 4 ? 2 : 32 / KB;
    std::vector<T> B((size_t)m * n), X((size_t)K * NB), Xglob((size_t)(xs0 + K + 64) * NB);
    for (auto &v : B) v = rnd<T>();
    for (auto &v : Xglob) v = rnd<T>();                      // rows around the block's range hold other data
    for (int j = 0; j < NB; ++j)
        for (int k = 0; k < K; ++k) X[(size_t)j * K + k] = Xglob[(size_t)j * (xs0 + K + 64) + xs0 + k];
    const int xd = S == 16 ? 0 : (S == 8 ? (xs0 & 1) : (xs0 & 3));
    const int kstep = (S == 8 && xd) ? 16 : 32;
    std::vector<W> ref((size_t)L * NB, W(0)), out((size_t)L * NB, W(0));
    for (int o = 0; o < mo; ++o)
        for (int j = 0; j < NB; ++j) {
            W s(0);
            for (int k = 0; k < K; ++k) {
                W a = tform ? (W)B[(size_t)o * m + k] : (W)B[(size_t)k * m + o];
                if constexpr (S == 16) { if (conj) a = std::conj(a); }
                s += a * (W)X[(size_t)j * K + k];
            }
            ref[(size_t)j * L + o] = s;
        }
    // stages of 32 contraction entries
    std::vector<W> accs((size_t)4 * MT * NT * 64, W(0));   // [warp][i][u][g][n]
    for (int k0 = 0; k0 < K; k0 += kstep) {
        const int kcv = std::min(kstep, K - k0);
        const int kbase = tform ? k0 : 0;
        std::vector<unsigned char> As(32 * 32 * S + 1024, 0xA5), Xs((size_t)XBoxes * NB * 128, 0x5A);
        const size_t slab_bytes = (size_t)(tform ? m * n : m * kcv) * S;
        tma_place_linear(As, 0, reinterpret_cast<const unsigned char *>(B.data() + (tform ? 0 : (size_t)k0 * m)), slab_bytes);
        // X boxes: KB contraction entries x NB columns each, fetched from the 16-byte aligned row xs0 + k0 - xd; only
        // the boxes that hold valid entries are fetched (the others keep stale bytes)
        const int nbox = (kcv + xd + KB - 1) / KB;
        if (nbox > XBoxes) { std::printf("X area too small\n"); return 1; }
        if (((size_t)(xs0 + k0 - xd) * S) % 16 != 0) { std::printf("misaligned X fetch\n"); return 1; }
        if (((size_t)(tform ? 0 : k0) * m * S) % 128 != 0) { std::printf("misaligned slab\n"); return 1; }
        for (int b = 0; b < nbox; ++b)
            for (int j = 0; j < NB; ++j)
                for (int kk = 0; kk < KB; ++kk) {
                    const int grow = xs0 + k0 - xd + b * KB + kk;
                    const size_t bytepos = (size_t)kk * S;
                    const size_t chunk = bytepos >> 4;
                    const size_t dst = (size_t)b * (NB * 128) + (size_t)j * 128 + ((chunk ^ (j & 7)) << 4) + (bytepos & 15);
                    std::memcpy(&Xs[dst], &Xglob[(size_t)j * (xs0 + K + 64) + grow], S);
                }
        const bool full = kcv == 32 && mo == 32 && L == 32 && m == 32 && xd == 0;
        const int Mt = (L + 7) >> 3;
        for (int warp = 0; warp < 4; ++warp) {
            const int wn = warp % WN, wm = warp / WN;
            const int nk4 = (kcv + 3) >> 2;
            for (int k4 = 0; k4 < nk4; ++k4) {
                W bfrag[NT][32], afrag[MT][32];
                int baddr[NT][32], aaddr[MT][32];
                bool tile_on[MT];
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tg = lane & 3;
                    const int k = 4 * k4 + tg;
                    const bool kv = k < kcv;
                    for (int u = 0; u < NT; ++u) {
                        const int j = NT * 8 * wn + 8 * u + ntile_col<S>(g);
                        const int idx = xtile_index<S, NB>(k + xd, j);
                        baddr[u][lane] = idx * S;
                        T v;
                        std::memcpy(&v, &Xs[(size_t)idx * S], S);
                        bfrag[u][lane] = kv ? (W)v : W(0);
                    }
                    for (int i = 0; i < MT; ++i) {
                        const int t = wm + WM * i;
                        tile_on[i] = t < Mt;
                        const int o = 8 * t + g;
                        const bool ok = kv && o < mo && t < Mt;
                        const int idx = ok ? swz128<S>(tform ? o * m + kbase + k : k * m + o) : 0;
                        aaddr[i][lane] = idx * S;
                        T v;
                        std::memcpy(&v, &As[(size_t)idx * S], S);
                        W a = ok ? (W)v : W(0);
                        if constexpr (S == 16) { if (conj) a = std::conj(a); }
                        afrag[i][lane] = a;
                    }
                }
                if (full) {
                    for (int u = 0; u < NT; ++u) *worstB = std::max(*worstB, conflict_degree<S>(baddr[u]));
                    for (int i = 0; i < MT; ++i) *worstA = std::max(*worstA, conflict_degree<S>(aaddr[i]));
                }
                for (int i = 0; i < MT; ++i) {
                    if (!tile_on[i]) continue;
                    for (int u = 0; u < NT; ++u)
                        for (int g = 0; g < 8; ++g)
                            for (int nn = 0; nn < 8; ++nn) {
                                W s(0);
                                for (int tg = 0; tg < 4; ++tg) s += afrag[i][4 * g + tg] * bfrag[u][4 * nn + tg];
                                accs[(((size_t)warp * MT + i) * NT + u) * 64 + g * 8 + nn] += s;
                            }
                }
            }
        }
    }
    // epilogue mapping
    const int Mt = (L + 7) >> 3;
    for (int warp = 0; warp < 4; ++warp) {
        const int wn = warp % WN, wm = warp / WN;
        for (int i = 0; i < MT; ++i) {
            const int t = wm + WM * i;
            if (t >= Mt) continue;
            for (int u = 0; u < NT; ++u)
                for (int g = 0; g < 8; ++g)
                    for (int nn = 0; nn < 8; ++nn) {
                        const int o = 8 * t + g, j = NT * 8 * wn + 8 * u + ntile_col<S>(nn);
                        if (o < L) out[(size_t)j * L + o] = accs[(((size_t)warp * MT + i) * NT + u) * 64 + g * 8 + nn];
                    }
        }
    }
    double err = 0, nrm = 0;
    for (size_t q = 0; q < ref.size(); ++q) {
        err += std::norm(std::complex<double>(out[q]) - std::complex<double>(ref[q]));
        nrm += std::norm(std::complex<double>(ref[q]));
    }
    if (std::sqrt(err) > 1e-12 * std::sqrt(nrm) + 1e-300 && !(S == 4 && std::sqrt(err) <= 1e-6 * std::sqrt(nrm))) {
        std::printf("MISMATCH S=%d NB=%d m=%d n=%d tform=%d L=%d rel=%g\n", S, NB, m, n, (int)tform, L, std::sqrt(err / nrm));
        return 1;
    }
    return 0;
}

template <int S, int NB>
static int run_all() {
    int bad = 0, wa = 1, wb = 1;
    const int shapes[][3] = {{32, 32, 32}, {32, 70, 32}, {24, 17, 24}, {8, 8, 8}, {5, 3, 7}, {32, 1, 32}, {17, 32, 32}, {1, 1, 1}, {31, 33, 31}};
    for (auto &sh : shapes)
        for (int xs0 = 0; xs0 < 4; ++xs0) {
            bad += run_case<S, NB>(sh[0], sh[1], false, sh[2], false, xs0, &wa, &wb);
            if (sh[1] <= 32) bad += run_case<S, NB>(sh[0], sh[1], true, std::max(sh[1], 1), S == 16, xs0, &wa, &wb);
        }
    std::printf("S=%d NB=%d worst conflict degree on full slabs: A %d, B %d\n", S, NB, wa, wb);
    return bad;
}

int main() {
    srand(12345);
    int bad = 0;
    bad += run_all<8, 64>() + run_all<8, 32>() + run_all<8, 16>() + run_all<8, 8>();
    bad += run_all<16, 32>() + run_all<16, 16>() + run_all<16, 8>();
    bad += run_all<4, 64>() + run_all<4, 32>() + run_all<4, 16>() + run_all<4, 8>();
    std::printf(bad ? "FAILED %d\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
