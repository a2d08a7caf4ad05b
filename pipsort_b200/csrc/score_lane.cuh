// Scoring path for batches of union configurations of up to LANE_KMAX SNPs: ONE LANE per configuration.
//
// Device restatement of a batch of expand_and_compute_lkl calls (sss_postcal.cpp:447-685, the body of the OpenMP loop
// sss_postcal.cpp:223-255): the stochastic shotgun search scores every neighbourhood with one launch of this kernel.
// The warp-per-configuration kernel of score.cuh left half of its lanes idle for k <= 4, factorised every one of the
// 2^k causal sub-blocks from scratch and walked all 3^k expansions per (SNP, state) cell.  Here, per lane:
//
//   1. gather   the k(k-1)/2 + 2k entries of W = d Sigma~, A, z of both studies (neighbours of one search state share all
//               but one SNP, so the rows are L1/L2 hits: HBM sees the LD rows of the current state once per launch);
//   2. tables   E_s(mask) for all 2^k sub-masks by BORDERING along a depth-first walk of the subset lattice: every mask
//               costs one row of forward substitution against its parent's Cholesky factor + one rsqrt + one exp
//               (fully unrolled per k: the factor lives in registers); the tables go to shared memory [mask][lane];
//   3. cells    the sum over the 3^k expansions (mask pairs with m0 | m1 = all, postcal.cpp:903-958) factorises because
//               the prior does: pi'(k, a) = pi'(k, 0) rho^a with a = |m0| + |m1| - k, rho = p / ((1 - p)/2).  With
//               g[m1] = rho^|m1| E_1[m1] and its superset sums Z[c] = sum_{m1 >= c} g[m1] (a zeta transform, k 2^(k-1)
//               additions) the total is sum_m0 rho^|m0| E_0[m0] Z[~m0] -- 2^k terms instead of 3^k -- and the per-SNP
//               cells (study 0 only / study 1 only / both) are the same sums with the transform left out along that
//               SNP's axis.  Only additions of non-negative terms: no cancellation;
//   4. max |l|  (sss_postcal.cpp:624-626) is attained at the largest or the smallest expansion weight: superset max / min
//               transforms of the same tables, one log each;
//   5. update   the cells of one SNP are summed over the warp's lanes (a neighbourhood adds ~10^5 terms to the cells of
//               the 4-5 SNPs of the current state) and leave it as one set of native fp64 atomics per SNP and warp.
//
// Numeric range: everything above is ordinary doubles, valid while every E_s(mask) < 2^450 (DESIGN.md "range").  A lane
// that meets a larger value hands its configuration to the mantissa/exponent code of score.cuh (whole warp, rare).
// Loci with p == 1 (rho infinite) and batches with more than LANE_KMAX SNPs per configuration use score.cuh throughout.
#pragma once
#include "exhaustive_dev.cuh"
#include "score.cuh"

namespace pipsort {

constexpr int LANE_KMAX = 5;
constexpr int LANE_WARPS = 4;
constexpr int LANE_THREADS = LANE_WARPS * 32;
constexpr int LANE_HOT_KEYS = 16;               // SNPs per warp-step whose cells are summed over the warp before they leave it
constexpr int LANE_TAB = 1 << LANE_KMAX;        // masks per table
#ifndef LANE_MINBLOCKS
#define LANE_MINBLOCKS 2
#endif

struct LaneBlockShared {
    double scal[3];                             // total, noCausal[0], noCausal[1] of the block (one lane per warp adds)
    double cnt;
};
constexpr size_t LANE_TAB_BYTES = (size_t)LANE_WARPS * 2 * LANE_TAB * 32 * sizeof(double);   // [warp][study][mask][lane]
constexpr size_t LANE_SMEM_BYTES = LANE_TAB_BYTES + sizeof(LaneBlockShared);

template <int K>
struct LaneStudy {                              // causal sub-block of one study, positions 0..K-1 (virtual: W = 0, A = 1, z = 0)
    double w[K > 1 ? K * (K - 1) / 2 : 1];      // W[i][j], i < j, at j (j - 1) / 2 + i
    double A[K], z[K];
};

__host__ __device__ constexpr int lane_tri(int i, int j) { return j * (j - 1) / 2 + i; }
__host__ __device__ constexpr int lane_nth_bit(int mask, int j) {
    for (int b = 0; b < 31; b++)
        if (mask >> b & 1) { if (j == 0) return b; j--; }
    return 0;
}

template <int K>
struct LaneDfsCtx {
    const LaneStudy<K>& S;
    double hd;
    double* tab;            // &table[study][0][lane]; stride 32 doubles per mask
    int absent;             // positions the study does not have: masks touching them hold 0
    double Lr[K][K];        // chain rows: Lr[d][j], j < d
    double inv[K], y[K];
    double emax;
    int bad;
};

// children MASK | 1 << h, h = H .. K-1, of the node MASK (D = popcount(MASK) elements on the chain, all below H)
template <int K, int MASK, int D, int H>
struct LaneVisit {
    static __device__ __forceinline__ void run(LaneDfsCtx<K>& c, const double Eparent) {
        if constexpr (H < K) {
            constexpr int CH = MASK | (1 << H);
            double s = c.S.A[H], r = c.S.z[H];
            double v[D > 0 ? D : 1];
#pragma unroll
            for (int j = 0; j < D; j++) {
                double t = c.S.w[lane_tri(lane_nth_bit(MASK, j), H)];
#pragma unroll
                for (int i = 0; i < j; i++) t = fma(-v[i], c.Lr[j][i], t);
                t *= c.inv[j];
                v[j] = t;
                s = fma(-t, t, s);
                r = fma(-t, c.y[j], r);
            }
            c.bad |= !(s > 0.25);                 // A >= I: every Schur complement is >= 1 (postcal.cpp:291-294)
            const double rs = rsqrt_fast(s);
            const double u = r * rs;
            const double E = Eparent * (exp_pos(c.hd * (u * u)) * rs);
            c.emax = fmax(c.emax, E);
            c.tab[CH * 32] = (CH & c.absent) ? 0.0 : E;
            if constexpr (H + 1 < K) {            // descend: the new element becomes chain row D
#pragma unroll
                for (int j = 0; j < D; j++) c.Lr[D][j] = v[j];
                c.inv[D] = rs;
                c.y[D] = u;
                LaneVisit<K, CH, D + 1, H + 1>::run(c, E);
            }
            LaneVisit<K, MASK, D, H + 1>::run(c, Eparent);
        }
    }
};

// superset sums / max / min along every axis of a 2^K array except axis SKIP (SKIP = -1: all axes)
template <int K, int SKIP, class Op>
__device__ __forceinline__ void lane_zeta(double (&z)[1 << K], Op op) {
#pragma unroll
    for (int a = 0; a < K; a++) {
        if (a == SKIP) continue;
#pragma unroll
        for (int c = 0; c < (1 << K); c++)
            if (!(c >> a & 1)) z[c] = op(z[c], z[c | (1 << a)]);
    }
}

// per-SNP cells of one lane from its two tables: out[i] = {study 0 only, study 1 only, both}
template <int K>
__device__ __forceinline__ void lane_cells(const double* __restrict__ t0, const double* __restrict__ t1, double (&out)[K][3]) {
    constexpr int NM = 1 << K, FULL = NM - 1;
#pragma unroll
    for (int i = 0; i < K; i++) {
        const int b = 1 << i, R = FULL ^ b;
        double z[NM];
#pragma unroll
        for (int m = 0; m < NM; m++) z[m] = t1[m * 32];
        // z[c] = sum over m1 >= c that agree with c on SNP i
        switch (i) {                              // the skipped axis must be a compile-time constant
            case 0: lane_zeta<K, 0>(z, [](double a, double q) { return a + q; }); break;
            case 1: lane_zeta<K, 1>(z, [](double a, double q) { return a + q; }); break;
            case 2: lane_zeta<K, 2>(z, [](double a, double q) { return a + q; }); break;
            case 3: lane_zeta<K, 3>(z, [](double a, double q) { return a + q; }); break;
            default: lane_zeta<K, 4>(z, [](double a, double q) { return a + q; }); break;
        }
        double x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
        for (int r0 = 0; r0 < NM; r0++) {
            if (r0 & b) continue;
            const int cmp = R ^ r0;               // the other SNPs that study 1 must carry
            const double ew = t0[(r0 | b) * 32], eo = t0[r0 * 32];
            x1 = fma(ew, z[cmp], x1);             // i in study 0 only
            x2 = fma(eo, z[cmp | b], x2);         // i in study 1 only
            x3 = fma(ew, z[cmp | b], x3);         // i in both
        }
        out[i][0] = x1; out[i][1] = x2; out[i][2] = x3;
    }
}

// Cells of the warp's 32 configurations -> accumulator store.  The configurations of a neighbourhood share all but one SNP
// (sss_postcal.cpp:20-99), so ~10^5 terms of one launch go to the cells of the 4-5 SNPs of the current state: the warp
// sums the cells of one SNP over its lanes (wherever the SNP sits in each lane's row) and issues ONE set of atomics for
// it -- for the first LANE_HOT_KEYS distinct SNPs in lane order; what is left after that (unrelated rows: every SNP
// different) goes out lane by lane, to distinct addresses.  Warp-collective; pending = positions of this lane to flush.
template <int K>
__device__ __forceinline__ void lane_flush_cells(const AccDev& acc, const int lane, const int (&g)[LANE_KMAX], int pending,
                                                 const double (&cell)[K][5]) {
    for (int it = 0; it < LANE_HOT_KEYS; it++) {
        const unsigned have = __ballot_sync(0xffffffffu, pending != 0);
        if (!have) return;
        const int src = __ffs(have) - 1;
        int mykey = -1;
        {
            const int p0 = __ffs(pending) - 1;
#pragma unroll
            for (int i = 0; i < K; i++) if (i == p0) mykey = g[i];
        }
        const int key = __shfl_sync(0xffffffffu, mykey, src);
        double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < K; i++) {
            if ((pending >> i & 1) && g[i] == key) {
#pragma unroll
                for (int q = 0; q < 5; q++) v[q] = cell[i][q];
                pending &= ~(1 << i);
            }
        }
        double mine = 0.0;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            double r = v[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == q) mine = r;
        }
        if (lane < 5) bin_add(acc, lane, key, mine, 0);
    }
#pragma unroll
    for (int i = 0; i < K; i++)
        if (pending >> i & 1) {
#pragma unroll
            for (int q = 0; q < 5; q++) bin_add(acc, q, g[i], cell[i][q], 0);
        }
}

// One warp-step: 32 configurations, all with at most K SNPs (K = the largest k of the warp; shorter ones are padded with
// virtual SNPs that exist in study 0 only and carry no weight).  g[i] = internal union index of position i (i < k).
template <int K>
__device__ __forceinline__ void lane_score(const LocusDev& L, LaneBlockShared& B, double* __restrict__ tabw, const int lane,
                                           const int (&g)[LANE_KMAX], const int k, const bool live, const bool upd,
                                           double& best, bool& slow) {
    constexpr int NM = 1 << K, FULL = NM - 1;
    const AccDev& acc = L.acc;
    double* t0 = tabw + lane;
    double* t1 = tabw + LANE_TAB * 32 + lane;
    int pres[2] = {0, 0};
    const int padmask = FULL & ~((1 << k) - 1);
    double emax = 0.0;
    int bad = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const StudyDev& S = L.st[s];
        LaneStudy<K> T;
        int loc[K];
#pragma unroll
        for (int i = 0; i < K; i++) {
            loc[i] = i < k ? L.loc[s][g[i]] : -1;
            const bool h = loc[i] >= 0;
            pres[s] |= h ? (1 << i) : 0;
            T.A[i] = h ? S.A[loc[i]] : 1.0;
            T.z[i] = h ? S.z[loc[i]] : 0.0;
        }
#pragma unroll
        for (int j = 1; j < K; j++)
#pragma unroll
            for (int i = 0; i < j; i++) {
                const bool h = loc[i] >= 0 && loc[j] >= 0;
                const int lo = min(loc[i], loc[j]), hi = max(loc[i], loc[j]);
                T.w[lane_tri(i, j)] = h ? S.W[(size_t)lo * S.ldw + hi] : 0.0;
            }
        LaneDfsCtx<K> c{T, S.hd, s == 0 ? t0 : t1, FULL & ~(pres[s] | (s == 0 ? padmask : 0)), {}, {}, {}, 0.0, 0};
        c.tab[0] = 1.0;
        LaneVisit<K, 0, 0, 0>::run(c, 1.0);
        emax = fmax(emax, c.emax);
        bad |= c.bad;
    }
    slow = live && (bad || !(emax < FAST_LIMIT));
    best = 0.0;
    int nvalid = 1;
#pragma unroll
    for (int i = 0; i < K; i++)
        if (i < k) nvalid *= ((pres[0] >> i & 1) && (pres[1] >> i & 1)) ? 3 : (((pres[0] | pres[1]) >> i & 1) ? 1 : 0);
    // a SNP that exists in neither study: no expansion passes checkOR -> nothing to score
    const bool go = live && !slow && nvalid > 0;
    const bool acc_on = go && upd;
    double cell[K][5];                            // X1 X2 X3 YS YN per position
    double tot = 0.0, nc0 = 0.0, nc1 = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
        for (int q = 0; q < 5; q++) cell[i][q] = 0.0;
    if (go) {
        // ---- likelihood-only cells (sharedLL / notSharedLL, postcal.cpp:1003-1016) from the plain tables ------------
        if (upd) {
            double ycell[K][3];
            lane_cells<K>(t0, t1, ycell);
#pragma unroll
            for (int i = 0; i < K; i++) { cell[i][YS] = ycell[i][2]; cell[i][YN] = ycell[i][0] + ycell[i][1]; }
        }
        // ---- prior weights: t0[m] *= pi'(k,0) rho^(|m| - K), t1[m] *= rho^|m|  (pads count in |m0| only) ------------
        {
            double rp[K + 1];
            rp[0] = 1.0;
#pragma unroll
            for (int j = 1; j <= K; j++) rp[j] = rp[j - 1] * L.rho;
            const double cx = L.pi[k][0] / rp[K];
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const int pc = __popc(m);         // compile-time after unrolling
                t0[m * 32] *= cx * rp[pc];
                t1[m * 32] *= rp[pc];
            }
        }
        // ---- max |l| over the expansions: the largest or the smallest weight ---------------------------------------
        {
            const double inf = __longlong_as_double(0x7ff0000000000000ll);
            double vhi = 0.0, vlo = inf;
            {
                double z[NM];
#pragma unroll
                for (int m = 0; m < NM; m++) z[m] = t1[m * 32];
                lane_zeta<K, -1>(z, [](double a, double q) { return fmax(a, q); });
#pragma unroll
                for (int m0 = 0; m0 < NM; m0++) vhi = fmax(vhi, t0[m0 * 32] * z[FULL ^ m0]);
            }
            {
                double z[NM];
#pragma unroll
                for (int m = 0; m < NM; m++) { const double v = t1[m * 32]; z[m] = v > 0.0 ? v : inf; }
                lane_zeta<K, -1>(z, [](double a, double q) { return fmin(a, q); });
#pragma unroll
                for (int m0 = 0; m0 < NM; m0++) {
                    const double e = t0[m0 * 32];
                    vlo = fmin(vlo, e > 0.0 ? e * z[FULL ^ m0] : inf);
                }
            }
            const double lhi = L.cx + log(vhi), llo = L.cx + log(vlo);
            best = fabs(lhi) > fabs(llo) ? lhi : llo;
        }
        // ---- prior-weighted cells (postValues / sharedPips, postcal.cpp:1003-1030) and the scalars --------------------
        if (upd) {
            double xcell[K][3];
            lane_cells<K>(t0, t1, xcell);
#pragma unroll
            for (int i = 0; i < K; i++) { cell[i][X1] = xcell[i][0]; cell[i][X2] = xcell[i][1]; cell[i][X3] = xcell[i][2]; }
            // every expansion has SNP 0 in exactly one state: the total; noCausal: the other study carries every SNP (:988-1000)
            tot = (xcell[0][0] + xcell[0][1]) + xcell[0][2];
            nc1 = t0[FULL * 32] * t1[0];
            nc0 = t0[padmask * 32] * t1[(FULL ^ padmask) * 32];
        }
    }
    // ---- out of the warp: cells per SNP, scalars per warp -----------------------------------------------------------
    if (__any_sync(0xffffffffu, acc_on)) {
        lane_flush_cells<K>(acc, lane, g, acc_on ? ((1 << k) - 1) : 0, cell);
        double r[4] = {tot, nc0, nc1, acc_on ? (double)nvalid : 0.0};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < 4; q++) r[q] += __shfl_xor_sync(0xffffffffu, r[q], o);
        if (lane == 0) {
            atomicAdd(&B.scal[S_TOTAL], r[0]); atomicAdd(&B.scal[S_NC0], r[1]); atomicAdd(&B.scal[S_NC1], r[2]);
            atomicAdd(&B.cnt, r[3]);
        }
    }
}

// Batch of union configurations in snp_map (user) order, -1 padded, kmax <= LANE_KMAX columns.
__global__ void __launch_bounds__(LANE_THREADS, LANE_MINBLOCKS)
score_lane_kernel(LocusDev L, const int* __restrict__ idx, long long n, int kmax, const unsigned char* __restrict__ make_updates,
                  double* __restrict__ out, const int* __restrict__ n_extra) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* tabs = reinterpret_cast<double*>(smem);
    LaneBlockShared& B = *reinterpret_cast<LaneBlockShared*>(smem + LANE_TAB_BYTES);
    if (n_extra) n += *n_extra;          // batch length decided on the device (sss.cuh: 1 + number of unseen neighbours)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* tabw = tabs + (size_t)wib * 2 * LANE_TAB * 32;
    if (threadIdx.x < 3) B.scal[threadIdx.x] = 0.0;
    if (threadIdx.x == 3) B.cnt = 0.0;
    __syncthreads();
    const long long nchunk = (n + LANE_THREADS - 1) / LANE_THREADS;
    for (long long ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
        const long long c = ch * LANE_THREADS + threadIdx.x;
        bool live = c < n;
        int g[LANE_KMAX];
        int k = 0;
        bool badrow = false;
        if (live) {
            int prev = -1;
#pragma unroll
            for (int i = 0; i < LANE_KMAX; i++) {
                g[i] = 0;
                const int v = i < kmax ? idx[c * kmax + i] : -1;
                if (v >= 0) {
                    // a row lists distinct union SNPs in increasing order, all below U (sss_postcal.cpp:20-99 builds them so)
                    if (v >= L.U || v <= prev) badrow = true;
                    else {
#pragma unroll
                        for (int q = 0; q < LANE_KMAX; q++) if (q == k) g[q] = L.u2i[v];
                        k++;
                    }
                    prev = v;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < LANE_KMAX; i++) g[i] = 0;
        }
        const bool upd = live && (make_updates ? make_updates[c] != 0 : true);
        if (badrow) { flag_set(L.acc, ERR_BAD_CONFIG); if (out) out[c] = 0.0; live = false; k = 0; }
        if (live && k == 0) {                       // the null configuration (sss_postcal.cpp:463-499)
            if (upd) {
                const double einv = 0.36787944117144233;
                atomicAdd(&B.scal[S_TOTAL], einv); atomicAdd(&B.scal[S_NC0], einv); atomicAdd(&B.scal[S_NC1], einv);
                atomicAdd(&B.cnt, 1.0);
            }
            if (out) out[c] = L.null_l;
            live = false;
        }
        const int kw = __reduce_max_sync(0xffffffffu, live ? k : 0);
        double best = 0.0;
        bool slow = false;
        switch (kw) {
            case 0: break;
            case 1: lane_score<1>(L, B, tabw, lane, g, k, live, upd, best, slow); break;
            case 2: lane_score<2>(L, B, tabw, lane, g, k, live, upd, best, slow); break;
            case 3: lane_score<3>(L, B, tabw, lane, g, k, live, upd, best, slow); break;
            case 4: lane_score<4>(L, B, tabw, lane, g, k, live, upd, best, slow); break;
            default: lane_score<5>(L, B, tabw, lane, g, k, live, upd, best, slow); break;
        }
        if (live && !slow && out) out[c] = best;
        // configurations outside the plain-double range: the whole warp scores them one by one with the mantissa/exponent
        // code of score.cuh (its workspace aliases this warp's tables, which are dead by now)
        unsigned todo = __ballot_sync(0xffffffffu, slow);
        if (todo) {
            __syncwarp();
            WarpWS ws = warp_ws(reinterpret_cast<unsigned char*>(tabw), LANE_KMAX);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int ks = __shfl_sync(0xffffffffu, k, src);
                const int us = __shfl_sync(0xffffffffu, (int)upd, src);
#pragma unroll
                for (int i = 0; i < LANE_KMAX; i++) {
                    const int gi = __shfl_sync(0xffffffffu, g[i], src);
                    if (lane == 0 && i < ks) ws.g[i] = gi;
                }
                __syncwarp();
                const double b = score_config(L, ws, ks, us != 0, lane);
                if (lane == src && out) out[c] = b;
                __syncwarp();
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // the block's scalars: plain doubles relative to exponent 0
    if (threadIdx.x < 3) bin_add(L.acc, SCAL, threadIdx.x, B.scal[threadIdx.x], 0);
    if (threadIdx.x == 3 && B.cnt > 0.0) atomicAdd(L.acc.counters, B.cnt);
}

}  // namespace pipsort
