// Exhaustive path: device code in exhaustive_dev.cuh, host-side work decomposition + launch below.
#pragma once
#include <algorithm>
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
    int occ2 = 0, occ3 = 0;   // resident blocks per SM of the two instantiations
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

// Launches the register kernel for the in-class rank range [rb, re) of size class j.  *done = false when the
// class is not covered (j not in {2,3} or a degenerate locus) and the caller must use the generic kernel.
// Returns a cudaError_t as int.
inline int exhaustive_launch(const LocusDev& L, const LocusDev* Lg, int j, u64 rb, u64 re, int sm_count, cudaStream_t stream, bool* done,
                             unsigned long long* launches, ExhScratch* sc) {
    *done = false;
    const int U = L.U;
    if ((j != 2 && j != 3) || U < j || rb >= re) return 0;
    cudaError_t err;
    if (!sc->d_counter) {
        if ((err = cudaMalloc(&sc->d_counter, sizeof(unsigned))) != cudaSuccess) return (int)err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sc->occ2, exhaustive_reg_kernel<2>, EXH_WARPS * 32, 0)) != cudaSuccess) return (int)err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sc->occ3, exhaustive_reg_kernel<3>, EXH_WARPS * 32, 0)) != cudaSuccess) return (int)err;
    }
    ExhParams P;
    memset(&P, 0, sizeof P);
    P.J = j;
    P.r_begin = rb; P.r_end = re;
    const u64 total = exh_binom(U, j);
    P.partial = !(rb == 0 && re == total);
    int glo[3] = {0, 0, 0}, ghi[3] = {U, U, U};
    exh_unrank(rb, U, j, glo);
    if (re < total) exh_unrank(re, U, j, ghi);
    for (int i = 0; i < 3; i++) { P.lo[i] = glo[i]; P.hi[i] = ghi[i]; }
    const int occ = std::max(1, j == 3 ? sc->occ3 : sc->occ2);
    const u64 slots = (u64)sm_count * occ * EXH_WARPS;
    // granularity: large items amortise the per-item setup, but there must be enough of them to balance
    const int cand[][2] = {{32, 1 << 20}, {32, 16}, {32, 8}, {32, 4}, {32, 2}, {32, 1}, {16, 1}, {8, 1}};
    std::vector<u64> prefix;
    u64 n_items = 0;
    for (const auto& cd : cand) {
        P.bw = cd[0]; P.xch = cd[1];
        prefix.clear();
        if (j == 3) {
            P.a_lo = glo[0];
            P.a_hi = std::min(re < total ? ghi[0] : U - 3, U - 3);
            prefix.push_back(0);
            for (int a = P.a_lo; a <= P.a_hi; a++) prefix.push_back(prefix.back() + exh_items_of(U, a, P.bw, P.xch));
            n_items = prefix.back();
        } else {
            P.a_lo = P.a_hi = -1;
            n_items = exh_items_of(U, -1, P.bw, P.xch);
            prefix = {0, n_items};
        }
        if (n_items >= 6 * slots) break;
    }
    if (n_items == 0) { *done = true; return 0; }
    if (n_items >= 0xffffff00ull) return 0;   // work queue is 32 bit: let the generic kernel take it (never in practice)
    if (sc->cap_prefix < prefix.size()) {
        if (sc->d_prefix) cudaFree(sc->d_prefix);
        sc->cap_prefix = prefix.size() * 2;
        if ((err = cudaMalloc(&sc->d_prefix, sc->cap_prefix * sizeof(u64))) != cudaSuccess) return (int)err;
    }
    if ((err = cudaMemcpyAsync(sc->d_prefix, prefix.data(), prefix.size() * sizeof(u64), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)err;
    if ((err = cudaMemsetAsync(sc->d_counter, 0, sizeof(unsigned), stream)) != cudaSuccess) return (int)err;
    P.item_prefix = sc->d_prefix;
    P.n_items = n_items;
    P.counter = sc->d_counter;
    const int blocks = (int)std::min<u64>((n_items + EXH_WARPS - 1) / EXH_WARPS, (u64)sm_count * occ);
    if (j == 3) exhaustive_reg_kernel<3><<<blocks, EXH_WARPS * 32, 0, stream>>>(L, P, Lg);
    else exhaustive_reg_kernel<2><<<blocks, EXH_WARPS * 32, 0, stream>>>(L, P, Lg);
    (*launches)++;
    if ((err = cudaGetLastError()) != cudaSuccess) return (int)err;
    *done = true;
    return 0;
}

}  // namespace pipsort
