// Stochastic shotgun search with a device-resident search state (sss_postcal.cpp:102-380).
//
// The reference keeps the neighbourhood as vector<vector<int>>, the explored configurations in an
// unordered_map<vector<int>, double> (postcal.h:43-56,98) and scores the unseen neighbours in an OpenMP loop
// (sss_postcal.cpp:223-255).  Here, per iteration, nothing but the current configuration (<= 8 ints) goes to the
// device and nothing but the neighbours' log-likelihoods (n doubles) comes back:
//   sss_lookup_kernel   thread i UNRANKS neighbour i of the current configuration directly (zero ++ minus ++ plus in the
//                       reference's order, sss_postcal.cpp:20-99,166-186), looks it up in an open-addressing hash table
//                       in HBM (key = the sorted union indices packed into 128 bits) and either copies the stored value
//                       or appends the configuration to the compact list of unseen ones;
//   score_batch_kernel  (score.cuh) expands + scores + accumulates the unseen list: one launch per neighbourhood;
//   sss_insert_kernel   stores the new values in the table (the reference also inserts after its loop, :280-284) and
//                       scatters them to the neighbours' slots.
// The sampling step (three within-group std::discrete_distribution draws and one across groups from
// std::mt19937(12345), sss_postcal.cpp:289-343) stays on the host: it must reproduce libstdc++'s sequence bit for bit,
// and it needs only the n doubles.
#pragma once
#include "common.cuh"

namespace pipsort {

typedef unsigned long long u64;

constexpr u64 SSS_OCC = 1ull << 63;      // "slot occupied" bit of the high key word
constexpr int SSS_IDX_BITS = 15;         // per union index (stored as idx + 1): U <= 32766
constexpr int SSS_MAX_U = (1 << SSS_IDX_BITS) - 2;

struct SssTable {
    u64* klo;
    u64* khi;
    double* val;
    u64 mask;                            // capacity - 1 (capacity is a power of two)
};

struct SssCur {                          // the current configuration, sorted ascending (snp_map order)
    int k;
    int g[KMAX];
};

__host__ __device__ inline void sss_key(const int* cfg, int k, u64& lo, u64& hi) {
    lo = 0; hi = SSS_OCC;
    for (int i = 0; i < k; i++) {
        const u64 v = (u64)(cfg[i] + 1);
        if (i < 4) lo |= v << (SSS_IDX_BITS * i);
        else hi |= v << (SSS_IDX_BITS * (i - 4));
    }
}

__host__ __device__ inline u64 sss_hash(u64 lo, u64 hi) {
    u64 x = lo ^ (hi * 0x9e3779b97f4a7c15ull);
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__device__ inline bool sss_find(const SssTable& t, u64 lo, u64 hi, double& v) {
    u64 s = sss_hash(lo, hi) & t.mask;
    for (;;) {
        const u64 h = t.khi[s];
        if (!(h & SSS_OCC)) return false;
        if (h == hi && t.klo[s] == lo) { v = t.val[s]; return true; }
        s = (s + 1) & t.mask;
    }
}

// distinct keys only (every configuration is inserted once); concurrent with other inserts, not with finds
__device__ inline void sss_put(const SssTable& t, u64 lo, u64 hi, double v) {
    u64 s = sss_hash(lo, hi) & t.mask;
    for (;;) {
        const u64 old = atomicCAS(t.khi + s, 0ull, hi);
        if (old == 0ull) { t.klo[s] = lo; t.val[s] = v; return; }
        s = (s + 1) & t.mask;
    }
}

// Sizes of the three neighbourhood groups of a configuration with k of U SNPs (sss_postcal.cpp:20-99)
__host__ __device__ inline void sss_nbd_sizes(int U, int k, int c, long long& nz, long long& nm, long long& np) {
    nz = (long long)(U - k) * k;
    nm = k;
    np = k < c ? (U - k) : 0;
}

// Neighbour i (in zero ++ minus ++ plus order) of cur -> sorted configuration out[0..kk); returns kk.
//   zero : added SNP outer (ascending over the non-members), dropped position inner (ascending)   (:72-99)
//   minus: dropped position ascending                                                            (:50-69)
//   plus : added SNP ascending over the non-members, empty when k >= c                           (:20-48)
__host__ __device__ inline int sss_neighbour(const SssCur& cur, int U, int c, long long i, int* out) {
    long long nz, nm, np;
    sss_nbd_sizes(U, cur.k, c, nz, nm, np);
    int drop = -1, add = -1;
    if (i < nz) { add = (int)(i / cur.k); drop = (int)(i % cur.k); }
    else if (i < nz + nm) drop = (int)(i - nz);
    else add = (int)(i - nz - nm);
    int g = -1;
    if (add >= 0) {                      // the add-th SNP that is not in cur
        g = add;
        for (int j = 0; j < cur.k; j++) if (cur.g[j] <= g) g++;
    }
    int kk = 0;
    bool placed = g < 0;
    for (int j = 0; j < cur.k; j++) {
        if (j == drop) continue;
        if (!placed && g < cur.g[j]) { out[kk++] = g; placed = true; }
        out[kk++] = cur.g[j];
    }
    if (!placed) out[kk++] = g;
    return kk;
}

// thread i < n: neighbour i; thread n: the current configuration itself (scored with updates only when unseen, :190-202)
__global__ void __launch_bounds__(256)
sss_lookup_kernel(SssTable tab, SssCur cur, int U, int c, long long n, int kmax, double* __restrict__ out_l,
                  int* __restrict__ batch, unsigned char* __restrict__ upd, int* __restrict__ unseen_idx, int* __restrict__ counter) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int cfg[KMAX];
    int kk;
    if (i == n) { kk = cur.k; for (int j = 0; j < kk; j++) cfg[j] = cur.g[j]; }
    else kk = sss_neighbour(cur, U, c, i, cfg);
    u64 lo, hi;
    sss_key(cfg, kk, lo, hi);
    double v = 0.0;
    const bool found = sss_find(tab, lo, hi, v);
    int row;
    if (i == n) {
        row = 0;
        upd[0] = found ? 0 : 1;
    } else {
        if (found) { out_l[i] = v; return; }
        const int pos = atomicAdd(counter, 1);
        unseen_idx[pos] = (int)i;
        row = 1 + pos;
        upd[row] = 1;
    }
    for (int j = 0; j < kmax; j++) batch[(size_t)row * kmax + j] = j < kk ? cfg[j] : -1;
}

__global__ void __launch_bounds__(256)
sss_insert_kernel(SssTable tab, const int* __restrict__ batch, int kmax, const double* __restrict__ scored,
                  const int* __restrict__ unseen_idx, const int* __restrict__ counter, double* __restrict__ out_l) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= *counter) return;
    const int* cfg = batch + (size_t)(1 + pos) * kmax;
    int kk = 0;
    int tmp[KMAX];
    for (int j = 0; j < kmax; j++) if (cfg[j] >= 0) tmp[kk++] = cfg[j];
    u64 lo, hi;
    sss_key(tmp, kk, lo, hi);
    const double v = scored[1 + pos];
    out_l[unseen_idx[pos]] = v;
    sss_put(tab, lo, hi, v);
}

__global__ void __launch_bounds__(256) sss_rehash_kernel(SssTable from, SssTable to) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > from.mask) return;
    const u64 h = from.khi[s];
    if (h & SSS_OCC) sss_put(to, from.klo[s], h, from.val[s]);
}

}  // namespace pipsort
