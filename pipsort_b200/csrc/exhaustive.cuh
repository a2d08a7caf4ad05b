// Exhaustive path: device code in exhaustive_dev.cuh, host-side work decomposition + launch below.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
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
    // the work decomposition of the last launch (a pass is usually repeated with the same c and rank range)
    char* pin_prefix = nullptr;   // pinned staging slice for the FIRST prefix upload (later ones may overlap an in-flight copy)
    size_t pin_prefix_cap = 0;
    bool plan_valid = false;
    int plan_c = 0;
    unsigned long long plan_rb = 0, plan_re = 0;
    unsigned char plan[512];
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
// ev0 / ev1 (optional) are recorded immediately around the kernel launch: the host-side planning below and the reset of
// the work-queue head are NOT part of the kernel's device time.
inline int exhaustive_launch_all(const LocusDev& L, const LocusDev* Lg, int c, u64 rank_begin, u64 rank_end, int sm_count,
                                 cudaStream_t stream, unsigned long long* launches, ExhScratch* sc, cudaEvent_t ev0 = nullptr,
                                 cudaEvent_t ev1 = nullptr) {
    const int U = L.U;
    cudaError_t err;
    static_assert(sizeof(ExhAll) <= sizeof(sc->plan), "plan cache too small");
    if (!sc->d_counter && (err = cudaMallocAsync(&sc->d_counter, sizeof(unsigned), stream)) != cudaSuccess) return (int)err;
    if (!sc->occ) {
        static int occ_cache = 0;    // a property of the kernel and the device generation: query once per process
        static std::mutex occ_mutex;
        std::lock_guard<std::mutex> g(occ_mutex);
        if (!occ_cache && (err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, exhaustive_all_kernel, EXH_WARPS * 32, 0)) != cudaSuccess) return (int)err;
        sc->occ = occ_cache;
    }
    ExhAll A;
    const bool knobs = getenv("PIPSORT_EXH_BW") || getenv("PIPSORT_EXH_ITEMS_PER_SLOT");   // experiments: plan afresh
    if (sc->plan_valid && !knobs && sc->plan_c == c && sc->plan_rb == rank_begin && sc->plan_re == rank_end) {
        memcpy(&A, sc->plan, sizeof A);
        if (A.n_total == 0) return 0;
        if ((err = cudaMemsetAsync(sc->d_counter, 0, sizeof(unsigned), stream)) != cudaSuccess) return (int)err;
        const int blocks = (int)std::min<u64>((A.n_total + EXH_WARPS - 1) / EXH_WARPS, (u64)sm_count * std::max(1, sc->occ));
        if (ev0 && (err = cudaEventRecord(ev0, stream)) != cudaSuccess) return (int)err;
        exhaustive_all_kernel<<<blocks, EXH_WARPS * 32, 0, stream>>>(L, A, Lg);
        if (ev1 && (err = cudaEventRecord(ev1, stream)) != cudaSuccess) return (int)err;
        (*launches)++;
        return (int)cudaGetLastError();
    }
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
    // Granularity (measured on B200, scripts/sweep_bw.sh): the full 32-wide b-window amortises the window table and the
    // per-tile x values best at every locus size; the number of x tiles per item is the largest one that still leaves
    // about a dozen items per resident warp for the work queue to balance (150 SNPs/study: (32,1), 0.068 ms instead of
    // 0.079 with 16-wide windows; 300: (32,1), 0.29 instead of 0.48 ms; 600: (32,3)).  Only loci with fewer items than
    // resident warps fall back to narrower windows.
    const int per_slot = [] {   // items per resident warp to aim for (tuning knob; read per launch so one process can sweep it)
        const char* v = getenv("PIPSORT_EXH_ITEMS_PER_SLOT");
        const int k = v ? atoi(v) : 0;
        return k > 0 ? k : 12;
    }();
    const int force_bw = [] { const char* v = getenv("PIPSORT_EXH_BW"); return v ? atoi(v) : 0; }();   // experiments
    const int force_xch = [] { const char* v = getenv("PIPSORT_EXH_XCH"); return v && atoi(v) > 0 ? atoi(v) : 1; }();
    auto choose = [&](auto count_items, u64 already, int& bw, int& xch) -> u64 {
        if (force_bw > 0) { bw = force_bw; xch = force_xch; return count_items(bw, xch); }
        const int xchs[] = {1 << 20, 32, 16, 12, 8, 6, 4, 3, 2, 1};
        u64 n = 0;
        bw = 32;
        xch = 1;
        const u64 n1 = count_items(bw, 1);                     // the finest 32-wide decomposition: an upper bound
        if (n1 + already >= (u64)per_slot * slots) {
            for (int x : xchs) {
                xch = x;
                n = x == 1 ? n1 : count_items(bw, xch);
                if (n + already >= (u64)per_slot * slots) return n;
            }
        }
        if ((n1 + already) * 2 >= slots) { bw = 32; xch = 1; return n1; }
        for (int w : {16, 8}) {       // small locus / small shard: narrower windows only when full ones leave most warps idle
            bw = w; xch = 1;          // (150 SNPs/study: 1678 items of (32,1) on 1776 warps beat 3020 items of (16,1))
            n = count_items(bw, xch);
            if ((n + already) * 2 >= slots) break;
        }
        return n;
    };
    std::vector<u64> prefix;
    if (have3) {
        ExhParams& P = A.p3;
        A.n3 = choose([&](int bw, int xch) { u64 n = 0; for (int a = P.a_lo; a <= P.a_hi; a++) n += exh_items_of(U, a, bw, xch); return n; },
                      0, P.bw, P.xch);
        if (A.n3 >= 0xfff00000ull) return (int)cudaErrorInvalidValue;   // 32-bit work queue (never in practice)
        if (!(sc->k_U == U && sc->k_bw == P.bw && sc->k_xch == P.xch && sc->k_alo == P.a_lo && sc->k_ahi == P.a_hi)) {
            prefix.push_back(0);
            for (int a = P.a_lo; a <= P.a_hi; a++) prefix.push_back(prefix.back() + exh_items_of(U, a, P.bw, P.xch));
            if (sc->cap_prefix < prefix.size()) {     // (never for a buffer that came from the engine's arena: U + 2 entries)
                if (sc->d_prefix) cudaFreeAsync(sc->d_prefix, stream);
                sc->cap_prefix = prefix.size() * 2;
                if ((err = cudaMallocAsync(&sc->d_prefix, sc->cap_prefix * sizeof(u64), stream)) != cudaSuccess) return (int)err;
            }
            const void* src = prefix.data();
            if (sc->pin_prefix && prefix.size() * sizeof(u64) <= sc->pin_prefix_cap) {
                memcpy(sc->pin_prefix, prefix.data(), prefix.size() * sizeof(u64));
                src = sc->pin_prefix;
                sc->pin_prefix = nullptr;      // one use
            }
            if ((err = cudaMemcpyAsync(sc->d_prefix, src, prefix.size() * sizeof(u64), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)err;
            sc->k_U = U; sc->k_bw = P.bw; sc->k_xch = P.xch; sc->k_alo = P.a_lo; sc->k_ahi = P.a_hi;
        }
        P.item_prefix = sc->d_prefix;
        P.n_items = A.n3;
    }
    if (have2) {
        ExhParams& P = A.p2;
        if (have3) { P.bw = 32; P.xch = 1; A.n2 = exh_items_of(U, -1, P.bw, P.xch); }   // a sliver next to the triples: finest items
        else A.n2 = choose([&](int bw, int xch) { return exh_items_of(U, -1, bw, xch); }, 0, P.bw, P.xch);
        P.n_items = A.n2;
    }
    A.n_total = A.n3 + A.n2 + (u64)A.n1_tiles + (u64)A.do_null;
    if (A.n_total == 0) return 0;
    static const bool debug = getenv("PIPSORT_EXH_DEBUG") != nullptr;
    if (debug) {
        const unsigned char* q = reinterpret_cast<const unsigned char*>(&A);
        unsigned long long h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof A; i++) { h ^= q[i]; h *= 1099511628211ull; }
        fprintf(stderr, "[exhaustive] params hash %016llx blocks=%d\n", h, (int)std::min<u64>((A.n_total + EXH_WARPS - 1) / EXH_WARPS, (u64)sm_count * occ));
    }
    if (debug)
        fprintf(stderr, "[exhaustive] U=%d slots=%llu  triples: bw=%d xch=%d items=%llu  pairs: bw=%d xch=%d items=%llu\n", U,
                (unsigned long long)slots, A.p3.bw, A.p3.xch, (unsigned long long)A.n3, A.p2.bw, A.p2.xch, (unsigned long long)A.n2);
    if ((err = cudaMemsetAsync(sc->d_counter, 0, sizeof(unsigned), stream)) != cudaSuccess) return (int)err;
    A.counter = sc->d_counter;
    memcpy(sc->plan, &A, sizeof A);
    sc->plan_valid = true; sc->plan_c = c; sc->plan_rb = rank_begin; sc->plan_re = rank_end;
    const int blocks = (int)std::min<u64>((A.n_total + EXH_WARPS - 1) / EXH_WARPS, (u64)sm_count * occ);
    if (ev0 && (err = cudaEventRecord(ev0, stream)) != cudaSuccess) return (int)err;
    exhaustive_all_kernel<<<blocks, EXH_WARPS * 32, 0, stream>>>(L, A, Lg);
    if (ev1 && (err = cudaEventRecord(ev1, stream)) != cudaSuccess) return (int)err;
    (*launches)++;
    return (int)cudaGetLastError();
}

}  // namespace pipsort
