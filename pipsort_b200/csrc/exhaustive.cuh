// Register kernel of the exhaustive path: one LANE per union subset (pair / triple), FP64 in registers.
//
// Replaces the OpenMP loop of PostCal::computeTotalLikelihood (postcal.cpp:769-1044) for the subset-size
// classes j = 2 and j = 3 (the reference's exhaustive limit, postcal.cpp:760-762).  For a triple (a < b < x)
// of internal union SNPs:
//   * a and b are WARP-UNIFORM (a fixed per work item, b walks a window of up to 32 consecutive SNPs),
//     x is the LANE's SNP (a 32-wide tile): LD rows a and b are read coalesced, everything that depends on
//     (a), (a,b) or (a,x) only is computed once per item / window / tile and reused;
//   * per study the kernel needs E_s(C) = exp(f_s(C)) for the 8 sub-masks of {a,b,x}; 6 of them are loop
//     invariant, the other two ({b,x} and {a,b,x}) cost one bordered Cholesky step + rsqrt + exp each;
//   * the <= 27 expansions (postcal.cpp:903-958) are products E_0[m0] E_1[m1]; they are summed per
//     (SNP, state) CELL relative to the cell's structurally largest term (all other SNPs causal in both
//     studies), so a plain double holds every cell without overflow and underflow is harmless
//     (DESIGN.md "cells");  SNPs a study does not have are "virtual" (their LD row is 0: E unchanged) and
//     the expansions that would use them are switched off by an exponent penalty -- no divergent code;
//   * x-cells accumulate in lane registers over the b window, a-cells in lane registers over the whole
//     item, b-cells are warp-reduced per step into a per-warp shared-memory window; everything leaves the
//     SM as native fp64 atomic adds into the exponent-binned accumulator store (common.cuh).
//
// Work decomposition: item = (a, b-window, chunk of x tiles), handed out by an atomic counter; the union-
// subset rank range [r_begin, r_end) of the C-ABI is honoured by a lexicographic predicate per lane.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace pipsort {

typedef unsigned long long u64;

constexpr int EXH_WARPS = 4;          // warps per block
constexpr int EXH_BW = 32;            // max b-window
constexpr int PEN = -4096;            // exponent penalty that switches an expansion off

struct E2 { double m; int n; };       // m * 2^n, m a positive normal double

// m * 2^e for a positive normal m; 0 when the result would leave the normal range downwards
__device__ __forceinline__ double scale2(double m, int e) {
    e = max(min(e, 900), -2040);
    const int hi = __double2hiint(m);
    const int f = (hi >> 20) + e;
    const double r = __hiloint2double(hi + (e << 20), __double2loint(m));
    return f > 0 ? r : 0.0;
}

// 1/sqrt(s), s >= 1 (Schur complements of A >= I): MUFU seed + 2 Newton steps (full double accuracy)
__device__ __forceinline__ double rsqrt_pos(double s) { return rsqrt(s); }

// E{C + y} = E{C} * exp(hd r^2 / s) / sqrt(s)   (bordered Cholesky step: s = Schur complement, r = residual)
__device__ __forceinline__ E2 extend(const E2& base, double hd, double r, double s, int& bad) {
    bad |= !(s > 0.25);                          // A >= I  =>  every Schur complement is >= 1 (postcal.cpp:291-294)
    const double rs = rsqrt_pos(s);
    const double u = r * rs;
    double em;
    int en;
    xexp(hd * (u * u), em, en);
    return E2{base.m * (em * rs), base.n + en};
}

// per-warp shared-memory window: everything that depends on (a, b) or on b alone, for <= 32 values of b
struct WinStudy {
    double Wab[EXH_BW];    // d Sigma[a][b]                      (0 when a or b is absent from the study)
    double inv22[EXH_BW];  // 1 / Schur(b | a)
    double c2[EXH_BW];     // residual(b | a) / Schur(b | a)
    double invAb[EXH_BW];  // 1 / A_b
    double ub[EXH_BW];     // z_b / A_b
    double m2[EXH_BW];     // E{b}   mantissa
    double m3[EXH_BW];     // E{a,b} mantissa
    double u3_0[EXH_BW];   // family relative to mask {a,b}:  2^-n3, E{a} 2^(n1-n3), E{b} 2^(n2-n3), with penalties
    double u3_1[EXH_BW];
    double u3_2[EXH_BW];
    double m3eff[EXH_BW];  // E{a,b} mantissa or 0 when a or b absent
    int n2[EXH_BW];        // exponent of E{b} (virtual semantics, no penalty)
    int n3[EXH_BW];        // exponent of E{a,b}
    int locb[EXH_BW];      // study-local index of b or -1
};
struct WarpWin {
    WinStudy st[2];
    double accM[EXH_BW][5];   // b-cell accumulators of the window (XAcc), flushed at the end of the item
    int accN[EXH_BW][5];
};

struct ExhParams {
    int J;                 // 2 or 3
    int bw;                // b-window size (<= 32)
    int xch;               // x tiles per item
    u64 r_begin, r_end;    // in-class rank range (lexicographic over internal order)
    int a_lo, a_hi;        // J == 3: range of a that intersects the rank range
    const u64* item_prefix;  // [a_hi - a_lo + 2] cumulative number of items (J == 3), or [2] for J == 2
    u64 n_items;
    unsigned* counter;     // work-queue head
    // lexicographic bounds of the rank range: first subset in range, first subset past it (or U,U,U)
    int lo[3], hi[3];
    bool partial;          // false: the whole class is in range (no per-lane predicate)
};

template <int J>
__device__ __forceinline__ bool lex_in_range(const ExhParams& P, int a, int b, int x) {
    // (a,b,x) >= lo && (a,b,x) < hi ; for J == 2 the tuple is (b,x) and a is ignored
    if (J == 3) {
        const bool ge = (a > P.lo[0]) || (a == P.lo[0] && (b > P.lo[1] || (b == P.lo[1] && x >= P.lo[2])));
        const bool lt = (a < P.hi[0]) || (a == P.hi[0] && (b < P.hi[1] || (b == P.hi[1] && x < P.hi[2])));
        return ge && lt;
    } else {
        const bool ge = (b > P.lo[0]) || (b == P.lo[0] && x >= P.lo[1]);
        const bool lt = (b < P.hi[0]) || (b == P.hi[0] && x < P.hi[1]);
        return ge && lt;
    }
}

constexpr __host__ __device__ bool in0(int t) { return t != 1; }   // state 0: study 0 only, 1: study 1 only, 2: both
constexpr __host__ __device__ bool in1(int t) { return t != 0; }

template <int J>
__global__ void __launch_bounds__(EXH_WARPS * 32, 2)
exhaustive_reg_kernel(LocusDev L, ExhParams P) {
    constexpr bool HAS_A = (J == 3);
    constexpr int FULL = HAS_A ? 7 : 6;     // bit 0 = a, bit 1 = b, bit 2 = x
    __shared__ WarpWin wins[EXH_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpWin& win = wins[wib];
    const AccDev& acc = L.acc;
    const int U = L.U;
    const double pi0 = L.pi[J][0], pi1 = L.pi[J][1], pi2 = L.pi[J][2], pi3 = HAS_A ? L.pi[J][3] : 0.0;

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(P.counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if ((u64)item >= P.n_items) break;

        // ---- decode the item: a, first b of the window, first x tile, number of x tiles ---------------
        int a = -1, b0, nb, xt0, nxt;
        {
            u64 rem = item;
            if (HAS_A) {
                int lo = 0, hi = P.a_hi - P.a_lo;            // largest i with prefix[i] <= item
                while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (P.item_prefix[mid] <= (u64)item) lo = mid; else hi = mid - 1; }
                a = P.a_lo + lo;
                rem = item - P.item_prefix[lo];
            }
            const int bfirst = a + 1, blast = U - 2;         // b in [bfirst, blast], x in (b, U-1]
            int w = 0;
            for (;; w++) {                                   // windows of this a, each split into x-tile chunks
                const int wb0 = bfirst + w * P.bw;
                const int t0 = (wb0 + 1) >> 5, t1 = (U - 1) >> 5;
                const u64 chunks = (u64)((t1 - t0 + 1 + P.xch - 1) / P.xch);
                if (rem < chunks) { b0 = wb0; xt0 = t0 + (int)rem * P.xch; nxt = min(P.xch, t1 - xt0 + 1); break; }
                rem -= chunks;
            }
            nb = min(P.bw, blast - b0 + 1);
        }

        // ---- per-item uniform values of a ---------------------------------------------------------------
        int ha[2] = {0, 0}, la[2] = {-1, -1};
        double invAa[2] = {1.0, 1.0}, ua[2] = {0.0, 0.0};
        E2 Ea[2] = {{1.0, 0}, {1.0, 0}};
        if (HAS_A) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                la[s] = L.loc[s][a];
                ha[s] = la[s] >= 0;
                if (ha[s]) {
                    invAa[s] = L.st[s].invA[la[s]];
                    ua[s] = L.st[s].u[la[s]];
                    Ea[s] = E2{L.st[s].e1m[la[s]], L.st[s].e1n[la[s]]};
                }
            }
        }
        const int pen_a[2] = {HAS_A ? (ha[0] ? 0 : PEN) : 0, HAS_A ? (ha[1] ? 0 : PEN) : 0};
        const int nstates_a = HAS_A ? (ha[0] && ha[1] ? 3 : (ha[0] || ha[1] ? 1 : 0)) : 1;

        int bad = 0;
        // ---- window table: lane t prepares b = b0 + t -----------------------------------------------------
        __syncwarp();
        {
            const int b = b0 + lane;
            const bool bv = lane < nb;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const StudyDev& S = L.st[s];
                WinStudy& w = win.st[s];
                const int lb = bv ? L.loc[s][b] : -1;
                const bool hb = lb >= 0;
                double Wab = 0.0, invAb = 1.0, ub = 0.0, Ab = 1.0, zb = 0.0;
                E2 Eb{1.0, 0};
                if (hb) {
                    invAb = S.invA[lb]; ub = S.u[lb]; Ab = S.A[lb]; zb = S.z[lb];
                    Eb = E2{S.e1m[lb], S.e1n[lb]};
                    if (HAS_A && ha[s]) Wab = S.W[(size_t)la[s] * S.ldw + lb];
                }
                // bordered step a -> {a,b}:  Schur = A_b - W_ab^2 / A_a,  residual = z_b - W_ab z_a / A_a
                const double s22 = fma(-Wab * Wab, invAa[s], Ab);
                const double r2 = fma(-Wab, ua[s], zb);
                const double inv22 = 1.0 / s22;
                E2 Eab = hb ? extend(Ea[s], S.hd, r2, s22, bad) : Ea[s];
                if (!(HAS_A && ha[s])) Eab = Eb;         // virtual a: E{a,b} = E{b}
                const int pen_b = hb ? 0 : PEN;
                w.Wab[lane] = Wab; w.inv22[lane] = inv22; w.c2[lane] = r2 * inv22;
                w.invAb[lane] = invAb; w.ub[lane] = ub;
                w.m2[lane] = Eb.m; w.n2[lane] = Eb.n;
                w.m3[lane] = Eab.m; w.n3[lane] = Eab.n;
                w.u3_0[lane] = scale2(1.0, -Eab.n);
                w.u3_1[lane] = scale2(Ea[s].m, Ea[s].n - Eab.n + pen_a[s]);
                w.u3_2[lane] = scale2(Eb.m, Eb.n - Eab.n + pen_b);
                w.m3eff[lane] = (pen_a[s] + pen_b) == 0 ? Eab.m : 0.0;
                w.locb[lane] = lb;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) { win.accM[lane][k] = 0.0; win.accN[lane][k] = XNEG; }
        }
        __syncwarp();

        // ---- item-lifetime accumulators (lane private): a-cells, total, noCausal, configuration count ------
        XAcc accA[5], accT = xacc_empty(), accNC0 = xacc_empty(), accNC1 = xacc_empty();
#pragma unroll
        for (int k = 0; k < 5; k++) accA[k] = xacc_empty();
        unsigned nconf = 0;

        for (int xt = xt0; xt < xt0 + nxt; xt++) {
            const int x = xt * 32 + lane;
            const bool xin = x < U;
            // ---- per-tile lane values of x: masks {x} (4) and {a,x} (5) ----------------------------------
            int hx[2], lx[2];
            double px[2], cx[2], rx[2], Axv[2], zxv[2];
            E2 E4[2], E5[2];
            double u5_0[2], u5_1[2], u5_4[2], m5eff[2];
            int n5[2], n4p[2];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const StudyDev& S = L.st[s];
                lx[s] = xin ? L.loc[s][x] : -1;
                hx[s] = lx[s] >= 0;
                double Wax = 0.0, Ax = 1.0, zx = 0.0;
                E4[s] = E2{1.0, 0};
                if (hx[s]) {
                    Ax = S.A[lx[s]]; zx = S.z[lx[s]];
                    E4[s] = E2{S.e1m[lx[s]], S.e1n[lx[s]]};
                    if (HAS_A && ha[s]) Wax = S.W[(size_t)la[s] * S.ldw + lx[s]];
                }
                Axv[s] = Ax; zxv[s] = zx;
                px[s] = Wax * invAa[s];
                cx[s] = fma(-Wax, px[s], Ax);
                rx[s] = fma(-Wax, ua[s], zx);
                if (HAS_A && ha[s] && hx[s]) E5[s] = extend(Ea[s], S.hd, rx[s], cx[s], bad);
                else E5[s] = hx[s] ? E4[s] : Ea[s];
                const int pen_x = hx[s] ? 0 : PEN;
                n5[s] = E5[s].n;
                n4p[s] = E4[s].n + pen_x;
                u5_0[s] = scale2(1.0, -n5[s]);
                u5_1[s] = scale2(Ea[s].m, Ea[s].n - n5[s] + pen_a[s]);
                u5_4[s] = scale2(E4[s].m, n4p[s] - n5[s]);
                m5eff[s] = (pen_a[s] + pen_x) == 0 ? E5[s].m : 0.0;
            }
            const int nstates_x = hx[0] && hx[1] ? 3 : (hx[0] || hx[1] ? 1 : 0);
            XAcc accX[5];
#pragma unroll
            for (int k = 0; k < 5; k++) accX[k] = xacc_empty();

            for (int t = 0; t < nb; t++) {
                const int b = b0 + t;
                if (b >= xt * 32 + 31) break;                       // no x of this tile is beyond b
                bool active = xin && x > b;
                if (P.partial) active = active && lex_in_range<J>(P, a, b, x);
                const int pen_act = active ? 0 : PEN;
                E2 E6[2], E7[2];
                double u7[2][8], u6[2][8], u3[2][4];
                int n7[2], n6[2], n3v[2];
                int hb[2];
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const StudyDev& S = L.st[s];
                    const WinStudy& w = win.st[s];
                    const int lb = w.locb[t];
                    hb[s] = lb >= 0;
                    const E2 E2b{w.m2[t], w.n2[t]}, E3{w.m3[t], w.n3[t]};
                    double Wbx = 0.0;
                    if (hb[s] && hx[s]) Wbx = S.W[(size_t)lb * S.ldw + lx[s]];
                    // mask {b,x}
                    if (hb[s] && hx[s]) {
                        const double s6 = fma(-Wbx * Wbx, w.invAb[t], Axv[s]);
                        const double r6 = fma(-Wbx, w.ub[t], zxv[s]);
                        E6[s] = extend(E2b, S.hd, r6, s6, bad);
                    } else {
                        E6[s] = hb[s] ? E2b : E4[s];
                    }
                    // mask {a,b,x}
                    if (HAS_A && ha[s] && hb[s] && hx[s]) {
                        const double tt = fma(-w.Wab[t], px[s], Wbx);
                        const double s7 = fma(-tt * tt, w.inv22[t], cx[s]);
                        const double r7 = fma(-tt, w.c2[t], rx[s]);
                        E7[s] = extend(E3, S.hd, r7, s7, bad);
                    } else {
                        E7[s] = !hx[s] ? E3 : (!hb[s] ? E5[s] : E6[s]);
                    }
                    const int pen_b = hb[s] ? 0 : PEN;
                    const int pen_x = (hx[s] ? 0 : PEN) + pen_act;
                    n7[s] = E7[s].n; n6[s] = E6[s].n; n3v[s] = E3.n;
                    // family relative to FULL (mask 7): numerators 1..6 with their penalties
                    u7[s][0] = 0.0;
                    u7[s][1] = HAS_A ? scale2(Ea[s].m, Ea[s].n + pen_a[s] - n7[s]) : 0.0;
                    u7[s][2] = scale2(E2b.m, E2b.n + pen_b - n7[s]);
                    u7[s][3] = HAS_A ? scale2(E3.m, E3.n + pen_a[s] + pen_b - n7[s]) : 0.0;
                    u7[s][4] = scale2(E4[s].m, n4p[s] + pen_act - n7[s]);
                    u7[s][5] = HAS_A ? scale2(E5[s].m, n5[s] + pen_a[s] + pen_x - n7[s]) : 0.0;
                    u7[s][6] = scale2(E6[s].m, n6[s] + pen_b + pen_x - n7[s]);
                    u7[s][7] = (pen_a[s] + pen_b + pen_x) == 0 ? E7[s].m : 0.0;
                    if (!HAS_A) {   // J == 2: FULL = 6; masks 6/7 coincide (a is virtual everywhere)
                        u7[s][6] = (pen_b + pen_x) == 0 ? E6[s].m : 0.0;
                    }
                    // family relative to mask 6 = FULL ^ a  (cells of a that miss this study)
                    u6[s][0] = scale2(1.0, -n6[s]);
                    u6[s][2] = scale2(E2b.m, E2b.n + pen_b - n6[s]);
                    u6[s][4] = scale2(E4[s].m, n4p[s] + pen_act - n6[s]);
                    u6[s][6] = (pen_b + pen_x) == 0 ? E6[s].m : 0.0;
                    // family relative to mask 3 = FULL ^ x: uniform, from the window table (J == 2: mask 2)
                    u3[s][0] = w.u3_0[t]; u3[s][1] = w.u3_1[t]; u3[s][2] = w.u3_2[t]; u3[s][3] = w.m3eff[t];
                }
                const double u54a[2] = {active ? u5_4[0] : 0.0, active ? u5_4[1] : 0.0};
                const double m5a[2] = {active ? m5eff[0] : 0.0, active ? m5eff[1] : 0.0};
                const int nstates_b = hb[0] && hb[1] ? 3 : (hb[0] || hb[1] ? 1 : 0);
                if (active) nconf += (unsigned)(nstates_a * nstates_b * nstates_x);

                // ---- cells: G[snp][state][a'] = sum over the expansions with that SNP in that state ----------
                // snp 0 = a, 1 = b, 2 = x ; a' = number of OTHER SNPs causal in both studies
                double G[3][3][3];
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int q = 0; q < 3; q++)
#pragma unroll
                        for (int r = 0; r < 3; r++) G[i][q][r] = 0.0;
#pragma unroll
                for (int tx = 0; tx < 3; tx++)
#pragma unroll
                    for (int tb = 0; tb < 3; tb++)
#pragma unroll
                        for (int ta = 0; ta < (HAS_A ? 3 : 1); ta++) {
                            const int m0 = (HAS_A && in0(ta) ? 1 : 0) | (in0(tb) ? 2 : 0) | (in0(tx) ? 4 : 0);
                            const int m1 = (HAS_A && in1(ta) ? 1 : 0) | (in1(tb) ? 2 : 0) | (in1(tx) ? 4 : 0);
                            const int ac = (HAS_A && ta == 2 ? 1 : 0) + (tb == 2 ? 1 : 0) + (tx == 2 ? 1 : 0);
                            // x cell: other-study family is relative to FULL ^ x (3, uniform)
                            {
                                const double f0 = in0(tx) ? u7[0][m0] : u3[0][m0];
                                const double f1 = in1(tx) ? u7[1][m1] : u3[1][m1];
                                G[2][tx][ac - (tx == 2 ? 1 : 0)] = fma(f0, f1, G[2][tx][ac - (tx == 2 ? 1 : 0)]);
                            }
                            // b cell: relative to FULL ^ b (5): lane constants of the tile (x part switched off when inactive)
                            {
                                const double f0 = in0(tb) ? u7[0][m0] : (m0 == 0 ? u5_0[0] : (m0 == 1 ? u5_1[0] : (m0 == 4 ? u54a[0] : m5a[0])));
                                const double f1 = in1(tb) ? u7[1][m1] : (m1 == 0 ? u5_0[1] : (m1 == 1 ? u5_1[1] : (m1 == 4 ? u54a[1] : m5a[1])));
                                G[1][tb][ac - (tb == 2 ? 1 : 0)] = fma(f0, f1, G[1][tb][ac - (tb == 2 ? 1 : 0)]);
                            }
                            if (HAS_A) {  // a cell: relative to FULL ^ a (6)
                                const double f0 = in0(ta) ? u7[0][m0] : u6[0][m0];
                                const double f1 = in1(ta) ? u7[1][m1] : u6[1][m1];
                                G[0][ta][ac - (ta == 2 ? 1 : 0)] = fma(f0, f1, G[0][ta][ac - (ta == 2 ? 1 : 0)]);
                            }
                        }
                // exponents of the cells' reference terms
                const int nfull = n7[0] + n7[1];
                // J == 2 note: with a virtual, masks 7/5/3 carry the values of 6/4/2, so the same code is right.
                const int nx_s0 = n7[0] + n3v[1], nx_s1 = n3v[0] + n7[1];       // x in study 0 only / study 1 only
                const int nb_s0 = n7[0] + n5[1], nb_s1 = n5[0] + n7[1];
                const int na_s0 = n7[0] + n6[1], na_s1 = n6[0] + n7[1];

                auto wsumX = [&](const double (&g)[3], bool both) -> double {   // prior-weighted cell value
                    return both ? fma(g[0], pi1, fma(g[1], pi2, g[2] * pi3)) : fma(g[0], pi0, fma(g[1], pi1, g[2] * pi2));
                };
                auto sumY = [&](const double (&g)[3]) -> double { return g[0] + g[1] + g[2]; };

                // x cells -> lane registers (flushed after the window)
                {
                    const double x1 = wsumX(G[2][0], false), x2 = wsumX(G[2][1], false), x3 = wsumX(G[2][2], true);
                    xadd(accX[X1], x1, nx_s0); xadd(accX[X2], x2, nx_s1); xadd(accX[X3], x3, nfull);
                    xadd(accX[YS], sumY(G[2][2]), nfull);
                    xadd(accX[YN], sumY(G[2][0]), nx_s0); xadd(accX[YN], sumY(G[2][1]), nx_s1);
                    xadd(accT, x1, nx_s0); xadd(accT, x2, nx_s1); xadd(accT, x3, nfull);   // every expansion once
                }
                if (HAS_A) {
                    xadd(accA[X1], wsumX(G[0][0], false), na_s0); xadd(accA[X2], wsumX(G[0][1], false), na_s1);
                    xadd(accA[X3], wsumX(G[0][2], true), nfull);
                    xadd(accA[YS], sumY(G[0][2]), nfull);
                    xadd(accA[YN], sumY(G[0][0]), na_s0); xadd(accA[YN], sumY(G[0][1]), na_s1);
                }
                // no causal SNP in study 1 (0): every chosen SNP causal in study 0 (1) only  (postcal.cpp:988-1000)
                xadd(accNC1, pi0 * u7[0][FULL], n7[0]);
                xadd(accNC0, pi0 * u7[1][FULL], n7[1]);
                // b cells -> warp reduction -> shared-memory window accumulators
                {
                    XAcc vb[5];
                    vb[X1] = XAcc{wsumX(G[1][0], false), nb_s0};
                    vb[X2] = XAcc{wsumX(G[1][1], false), nb_s1};
                    vb[X3] = XAcc{wsumX(G[1][2], true), nfull};
                    vb[YS] = XAcc{sumY(G[1][2]), nfull};
                    vb[YN] = XAcc{sumY(G[1][0]), nb_s0};
                    xadd(vb[YN], sumY(G[1][1]), nb_s1);
                    double rm = 0.0;
                    int rn = XNEG;
#pragma unroll
                    for (int k = 0; k < 5; k++) {
                        const XAcc r = xwarp_sum(vb[k]);
                        if (lane == k) { rm = r.M; rn = r.N; }
                    }
                    if (lane < 5) {
                        XAcc cur{win.accM[t][lane], win.accN[t][lane]};
                        xadd(cur, rm, rn);
                        win.accM[t][lane] = cur.M; win.accN[t][lane] = cur.N;
                    }
                }
            }  // b window

            // flush the x cells of this tile
            if (xin) {
#pragma unroll
                for (int k = 0; k < 5; k++) bin_add(acc, k, x, accX[k]);
            }
        }  // x tiles

        // ---- flush the item: b window, a cells, scalars -----------------------------------------------------
        __syncwarp();
        if (lane < nb) {
#pragma unroll
            for (int k = 0; k < 5; k++) bin_add(acc, k, b0 + lane, win.accM[lane][k], win.accN[lane][k]);
        }
        if (HAS_A) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const XAcc r = xwarp_sum(accA[k]);
                if (lane == 0) bin_add(acc, k, a, r);
            }
        }
        {
            const XAcc rt = xwarp_sum(accT), r0 = xwarp_sum(accNC0), r1 = xwarp_sum(accNC1);
            unsigned cnt = nconf;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) {
                bin_add(acc, SCAL, S_TOTAL, rt);
                bin_add(acc, SCAL, S_NC0, r0);
                bin_add(acc, SCAL, S_NC1, r1);
                atomicAdd(acc.counters, (u64)cnt);
            }
            if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(acc.counters + 1, (unsigned long long)ERR_NOT_PD);
        }
        __syncwarp();
    }
}

}  // namespace pipsort

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
inline int exhaustive_launch(const LocusDev& L, int j, u64 rb, u64 re, int sm_count, cudaStream_t stream, bool* done,
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
    if (j == 3) exhaustive_reg_kernel<3><<<blocks, EXH_WARPS * 32, 0, stream>>>(L, P);
    else exhaustive_reg_kernel<2><<<blocks, EXH_WARPS * 32, 0, stream>>>(L, P);
    (*launches)++;
    if ((err = cudaGetLastError()) != cudaSuccess) return (int)err;
    *done = true;
    return 0;
}

}  // namespace pipsort
