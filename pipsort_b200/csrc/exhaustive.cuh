// Exhaustive path: device code in exhaustive_dev.cuh, host-side work decomposition + launch below.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "exhaustive_dev.cuh"

// ---------------------------------------------------------------------------------------------------------
// host side: work decomposition + launch
// ---------------------------------------------------------------------------------------------------------
namespace pipsort {

struct ExhScratch {           // owned by the engine, reused across launches
    u64* d_prefix = nullptr;
    size_t cap_prefix = 0;
    unsigned* d_counter = nullptr;
    int occ = 0;              // resident blocks per SM of exhaustive_all_kernel
    // key of the prefix table currently on the device
    int k_U = -1, k_bw = 0, k_xch = 0, k_alo = 0, k_ahi = 0;
};

inline u64 exh_binom(int n, int k) {
    if (k < 0 || k > n) return 0;
    unsigned __int128 v = 1;
    for (int i = 1; i <= k; i++) v = v * (unsigned)(n - k + i) / (unsigned)i;
    return (u64)v;
}

inline void exh_unrank(u64 r, int U, int j, int* g) {
    int x0 = 0;
    for (int i = 0; i < j; i++) {
        const int jj = j - i;
        const u64 base = exh_binom(U - x0, jj);
        int lo = x0, hi = U - jj;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (base - exh_binom(U - mid, jj) <= r) lo = mid; else hi = mid - 1;
        }
        r -= base - exh_binom(U - lo, jj);
        g[i] = lo;
        x0 = lo + 1;
    }
}

// number of items of one `a` (J == 3) or of the whole class (J == 2, a = -1)
inline u64 exh_items_of(int U, int a, int bw, int xch) {
    const int bfirst = a + 1, blast = U - 2, t1 = (U - 1) >> 5;
    u64 n = 0;
    for (int wb0 = bfirst; wb0 <= blast; wb0 += bw) {
        const int t0 = (wb0 + 1) >> 5;
        n += (u64)((t1 - t0 + 1 + xch - 1) / xch);
    }
    return n;
}

// Fills the rank-range part of P for size class j; returns false when the class does not intersect.
inline bool exh_class_params(ExhParams& P, int U, int j, u64 rb, u64 re) {
    memset(&P, 0, sizeof P);
    P.J = j;
    if (U < j || rb >= re) return false;
    const u64 total = exh_binom(U, j);
    P.r_begin = rb; P.r_end = re;
    P.partial = !(rb == 0 && re == total);
    int glo[3] = {0, 0, 0}, ghi[3] = {U, U, U};
    exh_unrank(rb, U, j, glo);
    if (re < total) exh_unrank(re, U, j, ghi);
    for (int i = 0; i < 3; i++) { P.lo[i] = glo[i]; P.hi[i] = ghi[i]; }
    if (j == 3) {
        P.a_lo = glo[0];
        P.a_hi = std::min(re < total ? ghi[0] : U - 3, U - 3);
    } else {
        P.a_lo = P.a_hi = -1;
    }
    return true;
}

// Launches ONE kernel for the size classes 0..min(c,3) restricted to the global union-subset rank range
// [rank_begin, rank_end) (size-then-lexicographic order over the internal SNP order).  Returns a cudaError_t.
inline int exhaustive_launch_all(const LocusDev& L, const LocusDev* Lg, int c, u64 rank_begin, u64 rank_end, int sm_count,
                                 cudaStream_t stream, unsigned long long* launches, ExhScratch* sc) {
    const int U = L.U;
    cudaError_t err;
    if (!sc->d_counter) {
        if ((err = cudaMallocAsync(&sc->d_counter, sizeof(unsigned), stream)) != cudaSuccess) return (int)err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sc->occ, exhaustive_all_kernel, EXH_WARPS * 32, 0)) != cudaSuccess) return (int)err;
    }
    ExhAll A;
    memset(&A, 0, sizeof A);
    u64 off = 0;
    bool have3 = false, have2 = false;
    for (int j = 0; j <= std::min(std::min(c, 3), U); j++) {
        const u64 cnt = exh_binom(U, j);
        const u64 lo = std::max<u64>(rank_begin, off), hi = std::min<u64>(rank_end, off + cnt);
        if (lo < hi) {
            const u64 rb = lo - off, re = hi - off;
            if (j == 0) A.do_null = 1;
            else if (j == 1) { A.x1_lo = (int)rb; A.x1_hi = (int)re; A.tile1_0 = A.x1_lo >> 5; A.n1_tiles = ((A.x1_hi - 1) >> 5) - A.tile1_0 + 1; }
            else if (j == 2) have2 = exh_class_params(A.p2, U, 2, rb, re);
            else have3 = exh_class_params(A.p3, U, 3, rb, re);
        }
        off += cnt;
    }
    const int occ = std::max(1, sc->occ);
    const u64 slots = (u64)sm_count * occ * EXH_WARPS;
    static const int per_slot = [] {   // items per resident warp (tuning knob, default 6)
        const char* v = getenv("PIPSORT_EXH_ITEMS_PER_SLOT");
        const int k = v ? atoi(v) : 0;
        return k > 0 ? k : 6;
    }();
    // granularity: large items amortise the per-item setup, but there must be enough of them to balance
    const int cand[][2] = {{32, 1 << 20}, {32, 16}, {32, 8}, {32, 4}, {32, 2}, {32, 1}, {16, 1}, {8, 1}};
    std::vector<u64> prefix;
    if (have3) {
        ExhParams& P = A.p3;
        for (const auto& cd : cand) {
            P.bw = cd[0]; P.xch = cd[1];
            u64 n = 0;
            for (int a = P.a_lo; a <= P.a_hi; a++) n += exh_items_of(U, a, P.bw, P.xch);
            A.n3 = n;
            if (n >= (u64)per_slot * slots) break;
        }
        if (A.n3 >= 0xfff00000ull) return (int)cudaErrorInvalidValue;   // 32-bit work queue (never in practice)
        if (!(sc->k_U == U && sc->k_bw == P.bw && sc->k_xch == P.xch && sc->k_alo == P.a_lo && sc->k_ahi == P.a_hi)) {
            prefix.push_back(0);
            for (int a = P.a_lo; a <= P.a_hi; a++) prefix.push_back(prefix.back() + exh_items_of(U, a, P.bw, P.xch));
            if (sc->cap_prefix < prefix.size()) {
                if (sc->d_prefix) cudaFreeAsync(sc->d_prefix, stream);
                sc->cap_prefix = prefix.size() * 2;
                if ((err = cudaMallocAsync(&sc->d_prefix, sc->cap_prefix * sizeof(u64), stream)) != cudaSuccess) return (int)err;
            }
            if ((err = cudaMemcpyAsync(sc->d_prefix, prefix.data(), prefix.size() * sizeof(u64), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)err;
            sc->k_U = U; sc->k_bw = P.bw; sc->k_xch = P.xch; sc->k_alo = P.a_lo; sc->k_ahi = P.a_hi;
        }
        P.item_prefix = sc->d_prefix;
        P.n_items = A.n3;
    }
    if (have2) {
        ExhParams& P = A.p2;
        for (const auto& cd : cand) {
            P.bw = cd[0]; P.xch = cd[1];
            A.n2 = exh_items_of(U, -1, P.bw, P.xch);
            if (A.n2 + A.n3 >= (u64)per_slot * slots || (have3 && A.n2 >= slots / 4)) break;
        }
        P.n_items = A.n2;
    }
    A.n_total = A.n3 + A.n2 + (u64)A.n1_tiles + (u64)A.do_null;
    if (A.n_total == 0) return 0;
    if ((err = cudaMemsetAsync(sc->d_counter, 0, sizeof(unsigned), stream)) != cudaSuccess) return (int)err;
    A.counter = sc->d_counter;
    const int blocks = (int)std::min<u64>((A.n_total + EXH_WARPS - 1) / EXH_WARPS, (u64)sm_count * occ);
    exhaustive_all_kernel<<<blocks, EXH_WARPS * 32, 0, stream>>>(L, A, Lg);
    (*launches)++;
    return (int)cudaGetLastError();
}

}  // namespace pipsort
