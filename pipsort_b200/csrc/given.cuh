// Explicit-configuration path: device restatement of PostCal::computeTotalLikelihoodGivenConfigs
// (postcal.cpp:400-714, the -b/-d/-e flags).  Every row of the int16 matrix is ONE configuration (no mask
// expansion): global SNP indices offset_s + i in increasing order, negative = unused group
// (utils/construct_configs_all_studies.py:104-158 writes that layout).
//
// One THREAD per row: rows are independent, a row has at most KMAX causal SNPs per study, and the matrices
// this path is fed are large (the Cartesian product of per-group choices), so row-level parallelism fills
// the machine; the <= 8x8 Cholesky factor lives in the thread's local arrays and the causal LD entries are
// read straight from the L2-resident W.  Contributions go to the same exponent-binned accumulator store as
// every other path (one RED.ADD.F64 each).
//
// Rows the reference cannot process end the whole run there ("This did not work as expected", exit 1,
// postcal.cpp:593-596): entries that are not strictly increasing (the aux_idx walk :551-592 consumes the
// entries in (study, union position) order) -> ERR_BAD_CONFIG.  Entries >= N index past num_snps_all in the
// reference (undefined behaviour) and more than KMAX causal SNPs in one study / in the union exceed the
// engine's tables -> ERR_BAD_CONFIG as well.
#pragma once
#include "common.cuh"

namespace pipsort {

// f_s(C) for the k causal SNPs loc[0..k) of study st:  E = exp(f) = em * 2^en   (postcal.cpp:214-304 closed form)
__device__ inline void chol_block(const StudyDev& st, const int* loc, int k, double& em, int& en, bool& notpd) {
    em = 1.0;
    en = 0;
    if (k == 0) return;
    double Lp[KMAX * (KMAX + 1) / 2], y[KMAX];
    double q = 0.0, prodL = 1.0;
    for (int a = 0; a < k; a++) {
        const int ra = a * (a + 1) / 2;
        const double* Wrow = st.W + (size_t)loc[a] * st.ldw;
        for (int b = 0; b <= a; b++) {
            const int rb = b * (b + 1) / 2;
            double s = Wrow[loc[b]] + (a == b ? 1.0 : 0.0);
            for (int c = 0; c < b; c++) s -= Lp[ra + c] * Lp[rb + c];
            if (a == b) {
                if (!(s > 0.0)) { notpd = true; s = 1.0; }
                const double l = sqrt(s);
                Lp[ra + a] = l;
                prodL *= l;
            } else {
                Lp[ra + b] = s / Lp[rb + b];
            }
        }
        double ya = st.z[loc[a]];
        for (int c = 0; c < a; c++) ya -= Lp[ra + c] * y[c];
        ya /= Lp[ra + a];
        y[a] = ya;
        q += ya * ya;
    }
    xexp(st.hd * q, em, en);
    em /= prodL;
}

constexpr int GIVEN_THREADS = 128;

__global__ void __launch_bounds__(GIVEN_THREADS)
given_configs_kernel(LocusDev L, const short* __restrict__ configs, long long num_configs, int num_groups) {
    const AccDev& acc = L.acc;
    const long long row = (long long)blockIdx.x * GIVEN_THREADS + threadIdx.x;
    unsigned long long counted = 0;
    // every row adds to the three scalars (total, noCausal[0], noCausal[1]): summed over the warp first (common exponent =
    // the warp's largest), ONE atomic per warp instead of 32 on the same address -- 4 M rows hammering one L2 line were
    // 90 % of this kernel's time
    XAcc sTot = xacc_empty(), sNC0 = xacc_empty(), sNC1 = xacc_empty();
    if (row < num_configs) {
        const short* in = configs + row * num_groups;
        const int n0 = L.n_raw[0], N = L.n_raw[0] + L.n_raw[1];
        int loc[2][KMAX], un[2][KMAX], k[2] = {0, 0};
        bool bad = false;
        int prev = -1;
        for (int i = 0; i < num_groups; i++) {
            const int gi = in[i];
            if (gi < 0) continue;                                   // postcal.cpp:451 ">= 0" marks a causal entry
            if (gi <= prev || gi >= N) { bad = true; break; }
            prev = gi;
            const int s = gi >= n0 ? 1 : 0;
            const int l = L.raw2loc[s][gi - (s ? n0 : 0)];          // idx_to_union_pos_map (model.h:134-144)
            if (l < 0 || k[s] == KMAX) { bad = true; break; }
            loc[s][k[s]] = l;
            un[s][k[s]] = L.loc2u[s][l];
            k[s]++;
        }
        if (bad) {
            flag_set(acc, ERR_BAD_CONFIG);
        } else if (k[0] + k[1] == 0) {                              // postcal.cpp:461-492
            const double einv = 0.36787944117144233;                // exp(-1): "- sqrt(|1|)"
            sTot = XAcc{einv, 0}; sNC0 = XAcc{einv, 0}; sNC1 = XAcc{einv, 0};
            counted = 1;
        } else {
            // state of every causal union SNP: in study 0 only / study 1 only / both (:656-672)
            int both0 = 0, both1 = 0, a = 0;
            for (int i = 0; i < k[0]; i++)
                for (int t = 0; t < k[1]; t++)
                    if (un[0][i] == un[1][t]) { both0 |= 1 << i; both1 |= 1 << t; a++; }
            const int j = k[0] + k[1] - a;                          // unique union SNPs = numCausal (:516)
            if (j > KMAX) {
                flag_set(acc, ERR_BAD_CONFIG);
            } else {
                bool notpd = false;
                double m0, m1;
                int e0, e1;
                chol_block(L.st[0], loc[0], k[0], m0, e0, notpd);
                chol_block(L.st[1], loc[1], k[1], m1, e1, notpd);
                if (notpd) flag_set(acc, ERR_NOT_PD);
                const double y = m0 * m1, x = y * L.pi[j][a];
                const int ne = e0 + e1;
                if (x > 0.0) {
                    sTot = XAcc{x, ne};
                    if (k[0] == 0) sNC0 = XAcc{x, ne};               // :641-653
                    if (k[1] == 0) sNC1 = XAcc{x, ne};
                }
                for (int i = 0; i < k[0]; i++) {
                    const bool sh = both0 >> i & 1;
                    bin_add(acc, sh ? X3 : X1, un[0][i], x, ne);
                    bin_add(acc, sh ? YS : YN, un[0][i], y, ne);
                }
                for (int t = 0; t < k[1]; t++) {
                    if (both1 >> t & 1) continue;                   // counted once, from study 0
                    bin_add(acc, X2, un[1][t], x, ne);
                    bin_add(acc, YN, un[1][t], y, ne);
                }
                counted = 1;
            }
        }
    }
    {
        const XAcc t = xwarp_sum(sTot), n0 = xwarp_sum(sNC0), n1 = xwarp_sum(sNC1);
        if ((threadIdx.x & 31) == 0) {
            bin_add(acc, SCAL, S_TOTAL, t);
            bin_add(acc, SCAL, S_NC0, n0);
            bin_add(acc, SCAL, S_NC1, n1);
        }
    }
    // mycount (:475,619): one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, counted != 0);
    if ((threadIdx.x & 31) == 0 && m) count_add(acc, (unsigned long long)__popc(m));
}

}  // namespace pipsort
