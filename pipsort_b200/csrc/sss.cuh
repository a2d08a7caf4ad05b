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

// ---- the same search with the neighbourhood split over several GPUs (one process per GPU, sss_postcal.cpp:223-255 is the
// loop being split) -------------------------------------------------------------------------------------------------------
// Every rank runs the whole search in lockstep on a REPLICATED explored-configuration table (same seed, same values, hence
// the same trajectory); of every round's unseen neighbours rank r expands + scores + accumulates those with
// neighbour index i = r (mod world); the max-|l| values travel to every peer through the mailboxes of p2p.cuh (each double
// as two self-validating words, written straight into the peer's memory over NVLink); the accumulators stay rank-partial
// until they are combined (pipsort_p2p_reduce_to_root / all-reduce) at the end.
struct SssShard {
    int world, rank;
    unsigned flag;                       // this round's epoch (low 32 bits), never 0
    ulonglong2* peer_slot[16];           // [p]: where THIS rank's values go in peer p's mailbox (nullptr for p == rank)
    const ulonglong2* my_slot[16];       // [p]: where peer p's values arrive in this rank's mailbox
};

// thread i < n: neighbour i; thread n: the current configuration (scored by rank 0 when unseen).  out_l[i] gets the stored
// value of a seen neighbour; unseen ones are flagged (state[i] = 1) on every rank and appended to THIS rank's batch when
// i = rank (mod world) (pos_of[i] = row of the batch).
__global__ void __launch_bounds__(256)
sss_lookup_shard_kernel(SssTable tab, SssCur cur, int U, int c, long long n, int kmax, int world, int rank, double* __restrict__ out_l,
                        int* __restrict__ batch, unsigned char* __restrict__ upd, int* __restrict__ pos_of,
                        unsigned char* __restrict__ state, int* __restrict__ counter /* [0] own batch rows, [1] unseen in total */) {
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
    int row = -1;
    if (i == n) {
        // the current configuration is re-scored only when unseen, by rank 0; the other ranks keep row 0 as a no-op
        row = 0;
        upd[0] = (!found && rank == 0) ? 1 : 0;
        if (rank != 0) kk = -1;                                  // all -1: scored as the null configuration WITHOUT updates
    } else {
        state[i] = found ? 0 : 1;
        if (found) { out_l[i] = v; return; }
        atomicAdd(counter + 1, 1);
        if ((int)(i % world) != rank) return;
        const int pos = atomicAdd(counter, 1);
        pos_of[i] = 1 + pos;
        row = 1 + pos;
        upd[row] = 1;
    }
    for (int j = 0; j < kmax; j++) batch[(size_t)row * kmax + j] = j < kk ? cfg[j] : -1;
}

// thread i < n: an unseen neighbour scored by this rank sends its value to every peer
__global__ void __launch_bounds__(256)
sss_push_shard_kernel(SssShard sh, long long n, const unsigned char* __restrict__ state, const int* __restrict__ pos_of,
                      const double* __restrict__ scored) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !state[i] || (int)(i % sh.world) != sh.rank) return;
    const u64 b = (u64)__double_as_longlong(scored[pos_of[i]]);
    const u64 f = (u64)sh.flag << 32;
    const ulonglong2 w = make_ulonglong2((b & 0xffffffffull) | f, (b >> 32) | f);
    for (int p = 0; p < sh.world; p++)
        if (p != sh.rank) sh.peer_slot[p][i] = w;
}

// thread i < n: every unseen neighbour enters the (replicated) table on every rank, with the value its owner computed
__global__ void __launch_bounds__(256)
sss_insert_shard_kernel(SssTable tab, SssCur cur, int U, int c, long long n, SssShard sh, const unsigned char* __restrict__ state,
                        const int* __restrict__ pos_of, const double* __restrict__ scored, double* __restrict__ out_l,
                        double* __restrict__ err_flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !state[i]) return;
    const int owner = (int)(i % sh.world);
    double v;
    if (owner == sh.rank) {
        v = scored[pos_of[i]];
    } else {
        const volatile ulonglong2* src = sh.my_slot[owner] + i;
        u64 w0, w1;
        const long long t0 = clock64();
        bool ok = true;
        for (;;) {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src));
            if ((unsigned)(w0 >> 32) == sh.flag && (unsigned)(w1 >> 32) == sh.flag) break;
            if (clock64() - t0 > 120000000000ll) { ok = false; break; }
            __nanosleep(64);
        }
        if (!ok) { atomicAdd(err_flag, 1.0); return; }
        v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
    int cfg[KMAX];
    const int kk = sss_neighbour(cur, U, c, i, cfg);
    u64 lo, hi;
    sss_key(cfg, kk, lo, hi);
    out_l[i] = v;
    sss_put(tab, lo, hi, v);
}

// The running total (sss_sum_lkl, consulted by the convergence rule from round 100 on, sss_postcal.cpp:265-270) is the sum of
// the ranks' partial totals: every rank sends the bins of its total to every peer and adds all of them in rank order, so
// that all ranks see bit-identical sums and take the same branch.  One block.
__global__ void __launch_bounds__(64)
sss_total_shard_kernel(AccDev acc, SssShard sh, long long off, double cx, double* __restrict__ out_total, double* __restrict__ err_flag) {
    const int b = threadIdx.x;
    const int NB = acc.NB;
    const u64 f = (u64)sh.flag << 32;
    for (int t = b; t < NB; t += 64) {
        const u64 bits = (u64)__double_as_longlong(bin_ptr(acc, SCAL, S_TOTAL)[(size_t)t * acc.Upad]);
        const ulonglong2 w = make_ulonglong2((bits & 0xffffffffull) | f, (bits >> 32) | f);
        for (int p = 0; p < sh.world; p++)
            if (p != sh.rank) sh.peer_slot[p][off + t] = w;
    }
    // top three non-empty bins of the summed total (NB <= 64 * k handled by a serial tail in thread 0: NB is ~10-100)
    __shared__ double tot[512];
    for (int t = b; t < NB && t < 512; t += 64) {
        double v = 0.0;
        for (int p = 0; p < sh.world; p++) {
            double x;
            if (p == sh.rank) {
                x = bin_ptr(acc, SCAL, S_TOTAL)[(size_t)t * acc.Upad];
            } else {
                const volatile ulonglong2* src = sh.my_slot[p] + off + t;
                u64 w0, w1;
                const long long t0 = clock64();
                for (;;) {
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src));
                    if ((unsigned)(w0 >> 32) == sh.flag && (unsigned)(w1 >> 32) == sh.flag) break;
                    if (clock64() - t0 > 120000000000ll) { atomicAdd(err_flag, 1.0); break; }
                    __nanosleep(64);
                }
                x = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            }
            v += x;
        }
        tot[t] = v;
    }
    __syncthreads();
    if (b == 0) {
        int top = -1;
        for (int t = 0; t < NB && t < 512; t++) if (tot[t] > 0.0) top = t;
        double r = 0.0;
        if (top >= 0) {
            double M = tot[top];
            if (top > 0) M += tot[top - 1] * 0x1p-512;
            if (top > 1) M += (tot[top - 2] * 0x1p-512) * 0x1p-512;
            const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
            const double nn = (double)(512 * top - acc.bias);
            r = (nn * LN2_HI + (log(M) + nn * LN2_LO)) + cx;
        }
        *out_total = r;
    }
}

__global__ void __launch_bounds__(256) sss_rehash_kernel(SssTable from, SssTable to) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > from.mask) return;
    const u64 h = from.khi[s];
    if (h & SSS_OCC) sss_put(to, from.klo[s], h, from.val[s]);
}

}  // namespace pipsort
