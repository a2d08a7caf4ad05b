// Register kernel of the exhaustive path: one LANE per union subset (pair / triple), FP64 in registers.
//
// Replaces the OpenMP loop of PostCal::computeTotalLikelihood (postcal.cpp:769-1044) for the subset-size
// classes j = 2 and j = 3 (the reference's exhaustive limit, postcal.cpp:760-762).  For a triple (a < b < x)
// of internal union SNPs:
//   * a and b are WARP-UNIFORM (a fixed per work item, b walks a window of up to 32 consecutive SNPs),
//     x is the LANE's SNP (a 32-wide tile): LD rows a and b are read coalesced, everything that depends on
//     (a), (a,b) or (a,x) only is computed once per item / window / tile and reused;
//   * per study the kernel needs E_s(C) = exp(f_s(C)) for the 8 sub-masks of {a,b,x}; 6 of them are loop
//     invariant, E{b,x} comes with W[b][x] from the interleaved table WP (one 16-byte load per study and step) and
//     E{a,b,x} costs one bordered Cholesky step + rsqrt + exp;
//   * the <= 27 expansions (postcal.cpp:903-958) are products E_0[m0] E_1[m1], summed per (SNP, state) cell: the cells of
//     x and of a accumulate UNWEIGHTED over the steps of the tile / of a (the prior weights are applied once at the end),
//     the cells of b get theirs per step and are summed over the lanes through a shared-memory staging area;
//   * the step loop exists in versions specialised on what a (window, tile) segment cannot contain (no x of a study: its
//     loads, chain and two thirds of the products are not compiled in); the x tiles end at the boundaries between SNP
//     types, so most segments run a specialised version (run_steps, ExhTiles in exh_plan.h).
//
// Two numeric regimes (DESIGN.md "range"):
//   FAST  every E_s(mask) of the triple is below 2^450: all E's are ordinary doubles, every cell is a plain
//         double sum, accumulators are plain doubles in lane registers (x cells per tile, a cells per item) or
//         in the per-warp shared-memory window (b cells).  No exponent bookkeeping.
//   SLOW  otherwise (a very strongly associated SNP is involved): the lane notes the step in a mask and, after the
//         segment's steps, re-evaluates the triple with mantissa/exponent arithmetic, sums each cell relative to its
//         structurally largest term and adds the results straight to the binned store.  Per-lane, divergent, rare.
// Both regimes leave the SM as native fp64 atomic adds into the exponent-binned accumulator store (common.cuh).
//
// Work decomposition (exh_plan.h): the warp-steps of a size class form one fixed sequence (a, then 32-wide windows of b,
// then 32-wide tiles of x, then the b's of the window); a CHUNK is a contiguous run of it, handed out by an atomic counter.
// The union-subset rank range [r_begin, r_end) of the C-ABI is honoured per lane: for fixed a and x the admissible b's are an
// interval, computed once per segment from the lexicographic bounds.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "exh_plan.h"

namespace pipsort {

typedef unsigned long long u64;

#ifndef EXH_MINBLOCKS
#define EXH_MINBLOCKS 3
#endif
#ifndef EXH_WARPS_PER_BLOCK
#define EXH_WARPS_PER_BLOCK 4
#endif
constexpr int EXH_WARPS = EXH_WARPS_PER_BLOCK;   // warps per block
constexpr int EXH_BW = 32;            // max b-window
constexpr int EXH_STG = 6;            // steps whose b-cell values are staged in shared memory before their rows are summed (5 x 6 rows <= 32 lanes)
constexpr int EXH_STG_LD = 34;        // row length of the staging area in doubles (32 lanes + padding: 16-byte loads of a row stay conflict free)
constexpr int PEN = -4096;            // exponent penalty that switches an expansion off (slow path)
constexpr double FAST_LIMIT = 0x1p+450;

struct E2 { double m; int n; };       // m * 2^n, m a positive normal double

// m * 2^e for a positive normal m; 0 when the result would leave the normal range downwards
__device__ __forceinline__ double scale2(double m, int e) {
    e = max(min(e, 900), -2040);
    const int hi = __double2hiint(m);
    const int f = (hi >> 20) + e;
    const double r = __hiloint2double(hi + (e << 20), __double2loint(m));
    return f > 0 ? r : 0.0;
}

// exp(t) for t >= 0 as an ordinary double; +inf-like huge values when t is beyond the double range
__device__ __forceinline__ double exp_pos(double t) {
    double m;
    int n;
    xexp(t, m, n);
    n = min(n, 1000);
    return __hiloint2double(__double2hiint(m) + (n << 20), __double2loint(m));
}

// E{C + y} = E{C} * exp(hd r^2 / s) / sqrt(s)   (bordered Cholesky step: s = Schur complement, r = residual)
__device__ __forceinline__ E2 extend(const E2& base, double hd, double r, double s, int& bad) {
    bad |= !(s > 0.25);                          // A >= I  =>  every Schur complement is >= 1 (postcal.cpp:291-294)
    const double rs = rsqrt(s);
    const double u = r * rs;
    double em;
    int en;
    xexp(hd * (u * u), em, en);
    return E2{base.m * (em * rs), base.n + en};
}
// 1/sqrt(s) for s of ordinary magnitude (here: Schur complements >= 1 up to rounding): the MUFU.RSQ64H seed (2^-22)
// followed by one third-order correction  y0 (1 + e/2 + 3 e^2/8),  e = 1 - s y0^2  (error 5 e^3/16 < 2^-66).  The library
// rsqrt() wraps the same sequence in a branch + call for denormal / infinite arguments; that branch ends the basic block
// and keeps the compiler from interleaving the two studies' dependency chains.
__device__ __forceinline__ double rsqrt_fast(double s) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
    const double e = fma(s, -(y0 * y0), 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}
__device__ __forceinline__ double extend_fast(double base, double hd, double r, double s, int& bad) {
    bad |= !(s > 0.25);
    const double rs = rsqrt_fast(s);
    const double u = r * rs;
    return base * (exp_pos(hd * (u * u)) * rs);
}

// Pair table of one study: P[i][j] = E{i,j} = E{i} * exp(hd r^2 / s) / sqrt(s),  s = A_j - W_ij^2 / A_i,  r = z_j - W_ij z_i / A_i.
// E{b,x} does not depend on the third SNP of a triple, so the exhaustive kernel would otherwise recompute each entry once
// per `a` (U times): n^2 exponentials here replace n^3/3 there.  Values outside the fast range (or a lost positive
// definiteness) are stored as +inf: the kernel then sends the subset down its mantissa/exponent path, which also
// raises the error flag where the reference would stop.
// (from the values themselves: the preparation kernel builds the table in the same launch that computes them)
__device__ __forceinline__ double pair_entry(double hd, double e1m_i, int e1n_i, double invA_i, double u_i, double W, double A_j, double z_j) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double s = fma(-W * W, invA_i, A_j);
    const double r = fma(-W, u_i, z_j);
    double v = inf;
    if (e1n_i < 440 && s > 0.25) {
        int bad = 0;
        v = extend_fast(scale2(e1m_i, e1n_i), hd, r, s, bad);
        if (!(v < FAST_LIMIT)) v = inf;
    }
    return v;
}
__device__ __forceinline__ double pair_table_entry(const StudyDev& S, int i, int j) {
    if (!(j < S.n && j != i)) return 0.0;
    return pair_entry(S.hd, S.e1m[i], S.e1n[i], S.invA[i], S.u[i], S.W[(size_t)i * S.ldw + j], S.A[j], S.z[j]);
}

constexpr __host__ __device__ bool in0(int t) { return t != 1; }   // state 0: study 0 only, 1: study 1 only, 2: both
constexpr __host__ __device__ bool in1(int t) { return t != 0; }

struct ExhParams {
    int J;                 // 2 or 3
    u64 r_begin, r_end;    // in-class rank range (lexicographic over internal order)
    int a_lo, a_hi;        // J == 3: range of a that intersects the rank range
    int lo[3], hi[3];      // lexicographic bounds of the rank range: first subset in range, first subset past it
    bool partial;          // false: the whole class is in range (no per-lane predicate)
};

template <int J>
__device__ __forceinline__ bool lex_in_range(const ExhParams& P, int a, int b, int x) {
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

// ---------------------------------------------------------------------------------------------------------
// SLOW path: one subset, one lane, mantissa/exponent arithmetic, results straight to the bins.
// ---------------------------------------------------------------------------------------------------------
template <int J>
__device__ __noinline__ void slow_subset(const LocusDev& L, int a, int b, int x) {
    constexpr bool HAS_A = (J == 3);
    constexpr int FULL = HAS_A ? 7 : 6;
    const AccDev& acc = L.acc;
    int bad = 0;
    E2 E[2][8];
    int pen[2][3];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const StudyDev& S = L.st[s];
        const int la = HAS_A ? L.loc[s][a] : -1, lb = L.loc[s][b], lx = L.loc[s][x];
        const bool ha = la >= 0, hb = lb >= 0, hx = lx >= 0;
        pen[s][0] = (!HAS_A || ha) ? 0 : PEN; pen[s][1] = hb ? 0 : PEN; pen[s][2] = hx ? 0 : PEN;
        // virtual SNPs: A = 1, z = 0, W = 0  ->  E unchanged
        const double Aa = ha ? S.A[la] : 1.0, za = ha ? S.z[la] : 0.0;
        const double Ab = hb ? S.A[lb] : 1.0, zb = hb ? S.z[lb] : 0.0;
        const double Ax = hx ? S.A[lx] : 1.0, zx = hx ? S.z[lx] : 0.0;
        const double Wab = (ha && hb) ? S.W[(size_t)la * S.ldw + lb] : 0.0;
        const double Wax = (ha && hx) ? S.W[(size_t)la * S.ldw + lx] : 0.0;
        const double Wbx = (hb && hx) ? S.W[(size_t)lb * S.ldw + lx] : 0.0;
        const E2 one{1.0, 0};
        E[s][0] = one;
        E[s][1] = ha ? E2{S.e1m[la], S.e1n[la]} : one;
        E[s][2] = hb ? E2{S.e1m[lb], S.e1n[lb]} : one;
        E[s][4] = hx ? E2{S.e1m[lx], S.e1n[lx]} : one;
        const double iAa = 1.0 / Aa, iAb = 1.0 / Ab;
        const double s22 = fma(-Wab * Wab, iAa, Ab), r2 = fma(-Wab * za, iAa, zb);
        E[s][3] = extend(E[s][1], S.hd, r2, s22, bad);
        const double cx = fma(-Wax * Wax, iAa, Ax), rx = fma(-Wax * za, iAa, zx);
        E[s][5] = extend(E[s][1], S.hd, rx, cx, bad);
        const double s6 = fma(-Wbx * Wbx, iAb, Ax), r6 = fma(-Wbx * zb, iAb, zx);
        E[s][6] = extend(E[s][2], S.hd, r6, s6, bad);
        const double i22 = 1.0 / s22;
        const double tt = fma(-Wab * Wax, iAa, Wbx);
        const double s7 = fma(-tt * tt, i22, cx), r7 = fma(-tt * r2, i22, rx);
        E[s][7] = extend(E[s][3], S.hd, r7, s7, bad);
    }
    if (bad) flag_set(acc, ERR_NOT_PD);
    // numerator exponents carry the penalties of the SNPs the study lacks; reference exponents do not
    auto fam = [&](int s, int m, int ref) -> double {
        int p = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) if (m >> i & 1) p += pen[s][i];
        return scale2(E[s][m].m, E[s][m].n + p - E[s][ref].n);
    };
    const int snp[3] = {a, b, x};
    XAcc tot = xacc_empty();
#pragma unroll
    for (int i = HAS_A ? 0 : 1; i < 3; i++) {
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int bit = 1 << i;
            const int r0 = in0(t) ? FULL : FULL ^ bit, r1 = in1(t) ? FULL : FULL ^ bit;
            double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < (HAS_A ? 9 : 3); q++) {   // states of the other SNPs
                int st[3];
                st[i] = t;
                int qq = q;
#pragma unroll
                for (int k = HAS_A ? 0 : 1; k < 3; k++) if (k != i) { st[k] = qq % 3; qq /= 3; }
                int m0 = 0, m1 = 0, ac = 0;
#pragma unroll
                for (int k = HAS_A ? 0 : 1; k < 3; k++) {
                    if (in0(st[k])) m0 |= 1 << k;
                    if (in1(st[k])) m1 |= 1 << k;
                    if (k != i && st[k] == 2) ac++;
                }
                g[ac] = fma(fam(0, m0, r0), fam(1, m1, r1), g[ac]);
            }
            const int nref = E[0][r0].n + E[1][r1].n;
            const double pa = L.pi[J][t == 2 ? 1 : 0], pb = L.pi[J][t == 2 ? 2 : 1], pc = (HAS_A ? L.pi[J][t == 2 ? 3 : 2] : 0.0);
            const double xv = fma(g[0], pa, fma(g[1], pb, g[2] * pc));
            const double yv = g[0] + g[1] + g[2];
            bin_add(acc, t == 0 ? X1 : (t == 1 ? X2 : X3), snp[i], xv, nref);
            bin_add(acc, t == 2 ? YS : YN, snp[i], yv, nref);
            if (i == 2) xadd(tot, xv, nref);
        }
    }
    bin_add(acc, SCAL, S_TOTAL, tot);
    {   // no causal SNP in study 1 / study 0  (postcal.cpp:988-1000)
        int p0 = 0, p1 = 0;
#pragma unroll
        for (int i = HAS_A ? 0 : 1; i < 3; i++) { p0 += pen[0][i]; p1 += pen[1][i]; }
        if (p0 == 0) bin_add(acc, SCAL, S_NC1, L.pi[J][0] * E[0][FULL].m, E[0][FULL].n);
        if (p1 == 0) bin_add(acc, SCAL, S_NC0, L.pi[J][0] * E[1][FULL].m, E[1][FULL].n);
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-warp shared-memory window: everything that depends on (a, b) or on b alone, for <= 32 values of b
// ---------------------------------------------------------------------------------------------------------
struct alignas(16) WinRec {   // one b of one study: what a step reads, as three 16-byte (warp-uniform) loads
    double Wab;            // d Sigma[a][b]                      (0 when a or b is absent from the study)
    double inv22;          // 1 / Schur(b | a)
    double c2;             // sqrt(d/2) residual(b | a) / Schur(b | a)
    double v3;             // E{a,b} (0 when a or b is absent)
    double v2;             // E{b}   (0 when b is absent from the study)
    int row1;              // row of the NEXT b in the study's WP table (the step prefetches it)
    int ok;                // E{b}, E{a,b} within the fast range in BOTH studies
};
struct WinStudy {
    WinRec rec[EXH_BW];
    int row[EXH_BW + 4];   // row of b in the study's WP table; the all-zero row n when b is absent or past the window
};
__device__ __forceinline__ void win_load(const WinRec& r, double& Wab, double& inv22, double& c2, double& v3, double& v2, int& row1, int& ok) {
    const double2* p = reinterpret_cast<const double2*>(&r);
    const double2 q0 = p[0], q1 = p[1], q2 = p[2];
    Wab = q0.x; inv22 = q0.y; c2 = q1.x; v3 = q1.y; v2 = q2.x;
    row1 = __double2loint(q2.y); ok = __double2hiint(q2.y);
}
struct alignas(16) WarpWin {
    // staging area of the b-cell values: a step writes its five values lane by lane into rows (slot, cell); every EXH_STG
    // steps lane 5 slot + cell sums its row (16 loads of 16 bytes) into `part` -- no shuffles, and one addition per value
    // instead of a butterfly per step.  (First member: the rows are read as 16-byte words.)
    double stage[EXH_STG * 5][EXH_STG_LD];
    WinStudy st[2];
    // b-cell accumulators of the window, flushed when the window is left
    double part[EXH_BW][5];
    int cum[EXH_BW + 4];     // cum[t] = number of states of the b's before step t (prefix sums: configuration count per segment)
    // a cells, unweighted (GA[state][a'] per lane): they live as long as a does but are only touched once per tile
    double ga[9][32];
};

// One chunk of size class J (2 or 3): `remaining` warp-steps starting at step t_lo of the segment (a, window at b0,
// x tile xt), continuing through the following tiles, windows and a's.  Warp-collective.
template <int J>
__device__ __forceinline__ void exh_chunk(const LocusDev& L, const ExhParams& P, int a, int b0, int xt, int t_lo, int remaining,
                                          WarpWin& win, const LocusDev* __restrict__ Lg, const int lane) {
    constexpr bool HAS_A = (J == 3);
    constexpr int FAR = 1 << 24;
    constexpr unsigned LIM_HI = (1023u + 450u) << 20;      // high word of FAST_LIMIT: e < 2^450  <=>  hi(e) < LIM_HI  (e >= 0)
    const AccDev& acc = L.acc;
    const int U = L.U;
    const int T1 = L.ntiles - 1;
    const double pi0 = L.pi[J][0], pi1 = L.pi[J][1], pi2 = L.pi[J][2], pi3 = HAS_A ? L.pi[J][3] : 0.0;
    const double shd[2] = {sqrt(L.st[0].hd), sqrt(L.st[1].hd)};   // sqrt(d/2): folded into the residuals, so that the exponent is a plain square
    auto wsumX = [&](const double (&g)[3], bool both) -> double {   // prior-weighted cell value
        return both ? fma(g[0], pi1, fma(g[1], pi2, g[2] * pi3)) : fma(g[0], pi0, fma(g[1], pi1, g[2] * pi2));
    };
    auto sumY = [&](const double (&g)[3]) -> double { return g[0] + g[1] + g[2]; };
    // chunk-lifetime accumulators (lane private, plain doubles): the scalars.  The a cells GA[state][a'] live as long as a
    // does, UNWEIGHTED (the prior weights are applied when a is left), in the lane's column of win.ga: they are only touched
    // once per tile (below)
    double accT = 0.0, accNC0 = 0.0, accNC1 = 0.0;
    if (HAS_A) {
#pragma unroll
        for (int i = 0; i < 9; i++) win.ga[i][lane] = 0.0;
    }
    auto take_ga = [&](double (&g)[3][3]) {      // read the a cells and leave them empty
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int q = 0; q < 3; q++) { g[i][q] = win.ga[i * 3 + q][lane]; win.ga[i * 3 + q][lane] = 0.0; }
    };
    unsigned nconf = 0;
    int bad = 0;
    int ha[2] = {0, 0}, la[2] = {-1, -1};
    double invAa[2] = {1.0, 1.0}, ua[2] = {0.0, 0.0}, v1[2] = {0.0, 0.0};
    bool okA = true;
    int nstates_a = 1, nb = 0;
    bool new_a = true, new_win = true, win_has1 = true;           // win_has1: some b of the window is in study 1
    auto cells_of = [&](const double (&g)[3][3], double (&c)[5]) {   // X1 X2 X3 YS YN from the unweighted sums
        c[X1] = wsumX(g[0], false); c[X2] = wsumX(g[1], false); c[X3] = wsumX(g[2], true);
        c[YS] = sumY(g[2]);
        c[YN] = sumY(g[0]) + sumY(g[1]);
    };
    auto flush_a = [&]() {        // a cells of the a that is being left: five sums over the warp, one atomic each
        if (HAS_A) {
            double GA[3][3], accA[5], mine = 0.0;
            take_ga(GA);
            cells_of(GA, accA);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                double r = accA[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                if (lane == k) mine = r;
            }
            if (lane < 5) bin_add(acc, lane, a, mine, 0);
        }
    };
    auto flush_window = [&]() {   // b cells of the window that is being left
        __syncwarp();
        if (lane < nb) {
#pragma unroll
            for (int k = 0; k < 5; k++) bin_add(acc, k, b0 + lane, win.part[lane][k], 0);
        }
    };
    while (remaining > 0) {
        // ---- uniform values of a ------------------------------------------------------------------------------
        if (new_a) {
            new_a = false;
            okA = true;
            if (HAS_A) {
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    la[s] = L.loc[s][a];
                    ha[s] = la[s] >= 0;
                    invAa[s] = 1.0; ua[s] = 0.0; v1[s] = 0.0;
                    if (ha[s]) {
                        invAa[s] = L.st[s].invA[la[s]];
                        ua[s] = L.st[s].u[la[s]];
                        const int n = L.st[s].e1n[la[s]];
                        okA = okA && n < 440;
                        v1[s] = scale2(L.st[s].e1m[la[s]], min(n, 900));
                    }
                }
                if (!okA) { v1[0] = 0.0; v1[1] = 0.0; }     // every subset with this a takes the slow path
                nstates_a = ha[0] && ha[1] ? 3 : (ha[0] || ha[1] ? 1 : 0);
            }
        }

        // ---- window table: lane t prepares b = b0 + t -----------------------------------------------------
        if (new_win) {
            new_win = false;
            nb = min(EXH_BW, U - 1 - b0);
            __syncwarp();
            const int b = b0 + lane;
            const bool bv = lane < nb;
            int okb = 1, hbs = 0;
            WinRec rec[2];
            int rowv[2];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const StudyDev& S = L.st[s];
                const int lb = bv ? L.loc[s][b] : -1;
                const bool hb = lb >= 0;
                hbs += hb;
                double Wab = 0.0, Ab = 1.0, zb = 0.0, v2 = 0.0, v3 = 0.0;
                if (hb) {
                    Ab = S.A[lb]; zb = S.z[lb];
                    const int n = S.e1n[lb];
                    okb &= n < 440;
                    v2 = scale2(S.e1m[lb], min(n, 900));
                    if (HAS_A && ha[s]) {
                        const double2 wp = S.WP[(size_t)la[s] * S.ldp + lb];     // { d Sigma[a][b], E{a,b} }
                        Wab = wp.x; v3 = wp.y;
                        okb &= v3 < FAST_LIMIT;
                    }
                }
                // bordered step a -> {a,b}:  Schur = A_b - W_ab^2 / A_a,  residual = z_b - W_ab z_a / A_a
                const double s22 = fma(-Wab * Wab, invAa[s], Ab);
                const double r2 = fma(-Wab, ua[s], zb);
                const double inv22 = 1.0 / s22;
                if (HAS_A && ha[s] && hb) bad |= !(s22 > 0.25);
                rec[s].Wab = Wab; rec[s].inv22 = inv22; rec[s].c2 = shd[s] * (r2 * inv22);
                rec[s].v2 = v2; rec[s].v3 = v3;
                rowv[s] = hb ? lb : S.n;
            }
#pragma unroll
            for (int s = 0; s < 2; s++) {
                WinStudy& w = win.st[s];
                if (!okb) { rec[s].v2 = 0.0; rec[s].v3 = 0.0; }
                const int nxt_row = __shfl_down_sync(0xffffffffu, rowv[s], 1);
                rec[s].row1 = lane < EXH_BW - 1 ? nxt_row : L.st[s].n;
                rec[s].ok = okb;
                w.rec[lane] = rec[s];
                w.row[lane] = rowv[s];
                if (lane < 4) w.row[EXH_BW + lane] = L.st[s].n;
            }
            win_has1 = __any_sync(0xffffffffu, rowv[1] != L.st[1].n);
            {   // prefix sums of the b's numbers of states
                const int ns = hbs == 2 ? 3 : hbs;
                int c = ns;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += v; }
                win.cum[lane + 1] = c;
                if (lane == 0) win.cum[0] = 0;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) (&win.part[0][0])[k * 32 + lane] = 0.0;
            __syncwarp();
        }

        {
            const int x_lo = L.tile_lo[xt];
            const int x = x_lo + lane;
            const bool xin = x >= L.tile_vmin[xt];                       // (the lowest tile of a group of SNPs starts below the group)
            // ---- per-tile lane values of x: masks {x} (4) and {a,x} (5) ----------------------------------
            int hx[2];
            const double2* wpx[2];                                       // column of x in the study's WP table (the zero column when absent)
            double px[2], cx[2], rx[2], v4[2], v5[2];
            bool okX = true;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const StudyDev& S = L.st[s];
                const int lx = xin ? L.loc[s][x] : -1;
                hx[s] = lx >= 0;
                double Wax = 0.0, Ax = 1.0, zx = 0.0;
                v4[s] = 0.0; v5[s] = 0.0;
                wpx[s] = S.WP + (hx[s] ? lx : S.n);
                if (hx[s]) {
                    Ax = S.A[lx]; zx = S.z[lx];
                    const int n = S.e1n[lx];
                    okX = okX && n < 440;
                    v4[s] = scale2(S.e1m[lx], min(n, 900));
                    if (HAS_A && ha[s]) {
                        const double2 wp = wpx[s][(size_t)la[s] * S.ldp];          // { d Sigma[a][x], E{a,x} }
                        Wax = wp.x; v5[s] = wp.y;
                        okX = okX && v5[s] < FAST_LIMIT;
                    }
                }
                px[s] = Wax * invAa[s];
                cx[s] = fma(-Wax, px[s], Ax);
                rx[s] = shd[s] * fma(-Wax, ua[s], zx);
                if (HAS_A && ha[s] && hx[s]) bad |= !(cx[s] > 0.25);
            }
            okX = okX && okA;
            if (!okX) { v4[0] = 0.0; v4[1] = 0.0; v5[0] = 0.0; v5[1] = 0.0; }
            const int nstates_x = hx[0] && hx[1] ? 3 : (hx[0] || hx[1] ? 1 : 0);
            // ---- the steps [tA, tB) of this segment in which the lane's subset {a, b0 + t, x} exists and lies in the rank range
            // (lexicographic bounds: for fixed a and x the admissible b's are an interval)
            int tA = 0, tB = xin ? x - b0 : 0;                           // b < x
            if (P.partial) {
                int bmin, bmax;                                          // b in [bmin, bmax)
                if (HAS_A) {
                    bmin = a > P.lo[0] ? -FAR : (a == P.lo[0] ? (x >= P.lo[2] ? P.lo[1] : P.lo[1] + 1) : FAR);
                    bmax = a < P.hi[0] ? FAR : (a == P.hi[0] ? (x < P.hi[2] ? P.hi[1] + 1 : P.hi[1]) : -FAR);
                } else {
                    bmin = x >= P.lo[1] ? P.lo[0] : P.lo[0] + 1;
                    bmax = x < P.hi[1] ? P.hi[0] + 1 : P.hi[0];
                }
                tA = max(tA, bmin - b0);
                tB = min(tB, bmax - b0);
            }
            // SX[ta][tx][b in both studies?]: the tile's sums over the steps of the 27 products, b's state reduced to its class.
            // Both the x cells (summed over a's states) and the tile's share of the a cells (summed over x's states) follow
            // from it at the end of the tile: ONE accumulation per product and step instead of one for x and one for a.
            double SX[3][3][2];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int q = 0; q < 3; q++) { SX[i][q][0] = 0.0; SX[i][q][1] = 0.0; }
            // b cells: `ns` steps starting at step tb are staged; lane 5 slot + cell sums row (slot, cell) over the 32 lanes
            auto flush_stage = [&](int tb, int ns) {
                __syncwarp();
                if (lane < 5 * ns) {
                    const double2* row = reinterpret_cast<const double2*>(&win.stage[lane][0]);
                    double2 s0 = row[0], s1 = row[1];
#pragma unroll
                    for (int j = 2; j < 16; j += 2) {
                        const double2 u0 = row[j], u1 = row[j + 1];
                        s0.x += u0.x; s0.y += u0.y; s1.x += u1.x; s1.y += u1.y;
                    }
                    win.part[tb + lane / 5][lane % 5] += (s0.x + s0.y) + (s1.x + s1.y);
                }
                __syncwarp();
            };
            int slot = 0;                                                // staged steps
            // steps of this segment: the b's of the window that have an x of this tile beyond them, from t_lo on, as
            // far as the chunk reaches
            const int t_hi = min(min(nb, x_lo + 31 - b0), t_lo + remaining);
            remaining -= t_hi - t_lo;
            {   // expanded configurations of the lane's subsets in this segment (the reference's mycount)
                const int c_lo = max(tA, t_lo), c_hi = min(tB, t_hi);
                if (c_hi > c_lo) nconf += (unsigned)(nstates_a * nstates_x * (win.cum[c_hi] - win.cum[c_lo]));
            }
            const size_t ldp0 = L.st[0].ldp, ldp1 = L.st[1].ldp;
            unsigned slowmask = 0;
            // The steps of the segment, specialised at compile time on what the segment cannot contain (all warp-uniform):
            //   XT   0: the tile has x's of both studies;  1: no lane's x is in study 1;  2: none is in study 0
            //        -- a study without x has E = 0 for every mask with x: its loads, its chain and two thirds of the products go;
            //   CH0 / CH1: study s can have a non-zero E{a,b,x} here (a in the study, some b of the window, some x of the tile);
            //        without it the bordered step -> rsqrt -> exp chain of the study goes.
            // With the internal SNP order (union SNPs sorted by type: both studies, study 0 only, study 1 only) a < b < x are in
            // type order, windows and tiles are mostly of one type, and only the triples of three shared SNPs -- 30 % on the
            // synthetic loci -- need everything.  The generic version (0, true, true) is correct for any segment.
            auto run_steps = [&](auto xt_, auto ch0_, auto ch1_) {
            constexpr int XT = decltype(xt_)::value;
            constexpr bool USE[2] = {XT != 2, XT != 1};                  // study s has x's in this tile
            constexpr bool CH[2] = {HAS_A && decltype(ch0_)::value && USE[0], HAS_A && decltype(ch1_)::value && USE[1]};
            double2 nxt[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};   // { W[b][x], E{b,x} } of the NEXT step (software prefetch)
            if (USE[0]) nxt[0] = wpx[0][(size_t)win.st[0].row[t_lo] * ldp0];
            if (USE[1]) nxt[1] = wpx[1][(size_t)win.st[1].row[t_lo] * ldp1];
            for (int t = t_lo; t < t_hi; t++) {
                const bool active = t >= tA && t < tB;
                double v[2][8];
                unsigned emax = 0;                                       // largest high word of the step's E's (non-negative doubles order like integers)
                int smin = 0x7fffffff;                                   // smallest high word of the Schur complements
                // Branch-free on purpose: both studies' chains (bordered Cholesky step -> rsqrt -> exp) sit in ONE basic
                // block so that the compiler interleaves them.  Absent SNPs are "virtual" (W = 0, A = 1, z = 0, E = 0: the zero
                // row / column of the WP table): the arithmetic stays finite and the zero base E{a,b} or the final select
                // switch the expansion off.
                // (written stage by stage over both studies, see xexp_pair)
                double Wbx[2] = {0.0, 0.0}, e6[2] = {0.0, 0.0}, e7[2] = {0.0, 0.0};
                double wWab[2], wInv22[2], wC2[2], wV3[2], wV2[2];       // the window table's record of this step, both studies
                int wRow1[2], wOk[2];
#pragma unroll
                for (int s = 0; s < 2; s++) win_load(win.st[s].rec[t], wWab[s], wInv22[s], wC2[s], wV3[s], wV2[s], wRow1[s], wOk[s]);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    if (!USE[s]) continue;
                    Wbx[s] = nxt[s].x;
                    e6[s] = nxt[s].y;                                    // E{b,x} from the pair table (0 when b or x is absent)
                    nxt[s] = wpx[s][(size_t)wRow1[s] * (s ? ldp1 : ldp0)];
                }
                if (CH[0] || CH[1]) {
                    double tt[2], s7[2], r7[2], y0[2], ee[2], rs[2], uu[2], arg[2] = {0.0, 0.0}, em[2] = {0.0, 0.0};
                    int en[2] = {0, 0};
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) tt[s] = fma(-wWab[s], px[s], Wbx[s]);
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) s7[s] = fma(-tt[s] * tt[s], wInv22[s], cx[s]);
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[s]) : "d"(s7[s]));
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) r7[s] = fma(-tt[s], wC2[s], rx[s]);                 // sqrt(d/2) x the residual
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) ee[s] = fma(s7[s], -(y0[s] * y0[s]), 1.0);        // rsqrt_fast, both studies
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) rs[s] = fma(fma(ee[s], 0.375, 0.5), y0[s] * ee[s], y0[s]);
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) uu[s] = r7[s] * rs[s];
#pragma unroll
                    for (int s = 0; s < 2; s++) if (CH[s]) arg[s] = uu[s] * uu[s];
                    if (CH[0] && CH[1]) xexp_pair(arg, em, en);
                    else if (CH[0]) xexp(arg[0], em[0], en[0]);
                    else xexp(arg[1], em[1], en[1]);
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        if (!CH[s]) continue;
                        en[s] = min(en[s], 1000);
                        const double ex = __hiloint2double(__double2hiint(em[s]) + (en[s] << 20), __double2loint(em[s]));
                        const double e = wV3[s] * (ex * rs[s]);                                     // v3 = 0 when a or b is absent
                        e7[s] = hx[s] ? e : 0.0;
                        smin = min(smin, __double2hiint(s7[s]));                  // (signed: negative values fail too; NaN shows up in e)
                        emax = max(emax, (unsigned)__double2hiint(e));
                    }
                }
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    if (USE[s]) emax = max(emax, (unsigned)__double2hiint(e6[s]));
                    v[s][0] = 1.0; v[s][1] = v1[s]; v[s][2] = wV2[s]; v[s][3] = wV3[s];
                    v[s][4] = v4[s]; v[s][5] = v5[s]; v[s][6] = e6[s]; v[s][7] = e7[s];
                }
                // fast path: every E below 2^450 (as unsigned high words: negative or NaN fails), every Schur complement >= 1/4
                const bool ok = okX && wOk[0] && emax < LIM_HI && smin >= 0x3fd00000;
                if (active && !ok) slowmask |= 1u << t;             // rare: re-evaluated after the loop (slow_subset)
                {                                                   // a lane that is off contributes nothing on the fast path
                    const bool on = active && ok;
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        if (!USE[s]) continue;
                        v[s][4] = on ? v[s][4] : 0.0; v[s][5] = on ? v[s][5] : 0.0;
                        v[s][6] = on ? v[s][6] : 0.0; v[s][7] = on ? v[s][7] : 0.0;
                    }
                }

                // ---- cells: G[state][a'] = sum over the expansions with one SNP in that state, a' = number of OTHER SNPs causal
                // in both studies.  The x and a cells accumulate over the steps of the tile through SX (above): the 27 products are
                // all the arithmetic they cost per step.  The b cells of the step get their prior weights here.
                double GB[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
                for (int tx = 0; tx < 3; tx++)
#pragma unroll
                    for (int tb = 0; tb < 3; tb++)
#pragma unroll
                        for (int ta = 0; ta < (HAS_A ? 3 : 1); ta++) {
                            const int m0 = (HAS_A && in0(ta) ? 1 : 0) | (in0(tb) ? 2 : 0) | (in0(tx) ? 4 : 0);
                            const int m1 = (HAS_A && in1(ta) ? 1 : 0) | (in1(tb) ? 2 : 0) | (in1(tx) ? 4 : 0);
                            // (an expansion whose x state needs a study the tile does not have, or an E{a,b,x} that cannot exist here)
                            if ((in0(tx) && !USE[0]) || (in1(tx) && !USE[1])) continue;
                            if (HAS_A && ((m0 == 7 && !CH[0]) || (m1 == 7 && !CH[1]))) continue;
                            const int ac = (HAS_A && ta == 2 ? 1 : 0) + (tb == 2 ? 1 : 0) + (tx == 2 ? 1 : 0);
                            SX[ta][tx][tb == 2 ? 1 : 0] = fma(v[0][m0], v[1][m1], SX[ta][tx][tb == 2 ? 1 : 0]);
                            GB[tb][ac - (tb == 2 ? 1 : 0)] = fma(v[0][m0], v[1][m1], GB[tb][ac - (tb == 2 ? 1 : 0)]);
                        }
                if (HAS_A ? CH[0] : USE[0]) accNC1 = fma(pi0, v[0][HAS_A ? 7 : 6], accNC1);
                if (HAS_A ? CH[1] : USE[1]) accNC0 = fma(pi0, v[1][HAS_A ? 7 : 6], accNC0);
                {   // b cells: staged lane by lane; the rows are summed every EXH_STG steps
                    double q[5];
                    cells_of(GB, q);
#pragma unroll
                    for (int k = 0; k < 5; k++) win.stage[slot * 5 + k][lane] = q[k];
                    if (++slot == EXH_STG) { flush_stage(t + 1 - EXH_STG, EXH_STG); slot = 0; }
                }
            }  // b window
            };  // run_steps
            {
                typedef std::integral_constant<int, 0> X0; typedef std::integral_constant<int, 1> X1_; typedef std::integral_constant<int, 2> X2_;
                const bool tile0 = __any_sync(0xffffffffu, hx[0]), tile1 = __any_sync(0xffffffffu, hx[1]);
                if (!tile1) run_steps(X1_{}, std::true_type{}, std::false_type{});
                else if (!tile0) {
                    if (!HAS_A || (ha[1] && win_has1)) run_steps(X2_{}, std::false_type{}, std::true_type{});
                    else run_steps(X2_{}, std::false_type{}, std::false_type{});
                } else run_steps(X0{}, std::true_type{}, std::true_type{});
            }
            if (slot) flush_stage(t_hi - slot, slot);

            double GX[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};   // x cells of the tile, a cells of (a, tile): unweighted
            double GT[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
            for (int ta = 0; ta < (HAS_A ? 3 : 1); ta++)
#pragma unroll
                for (int tx = 0; tx < 3; tx++)
#pragma unroll
                    for (int cb = 0; cb < 2; cb++) {
                        GX[tx][cb + (ta == 2 ? 1 : 0)] += SX[ta][tx][cb];
                        if (HAS_A) GT[ta][cb + (tx == 2 ? 1 : 0)] += SX[ta][tx][cb];
                    }
            if (HAS_A) {
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int q = 0; q < 3; q++) win.ga[i * 3 + q][lane] += GT[i][q];
            }
            double accX[5];
            cells_of(GX, accX);
            accT += (accX[X1] + accX[X2]) + accX[X3];   // every expansion is in exactly one x cell: the total, once per tile
            if (xin) {   // flush the x cells of this tile
#pragma unroll
                for (int k = 0; k < 5; k++) bin_add(acc, k, x, accX[k], 0);
            }
            while (slowmask) {                          // rare, divergent, self-contained
                const int t = __ffs(slowmask) - 1;
                slowmask &= slowmask - 1;
                slow_subset<J>(*Lg, a, b0 + t, x);
            }
            __syncwarp();
        }  // segment
        // ---- advance: next tile, else next window, else next a --------------------------------------------------------
        if (remaining > 0) {
            t_lo = 0;
            xt++;
            if (xt > T1) {
                flush_window();
                b0 += EXH_BW;
                new_win = true;
                if (b0 > U - 2) {
                    if (!HAS_A) { nb = 0; break; }                  // (the plan never runs past the end of the pairs)
                    flush_a();
                    a++;
                    b0 = a + 1;
                    new_a = true;
                }
                xt = L.tile_of[b0 + 1];
            }
        }
    }
    // ---- end of the chunk: b window, a cells, scalars ---------------------------------------------------------------
    flush_window();
    {
        double r[8], accA[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (HAS_A) {
            double GA[3][3];
            take_ga(GA);
            cells_of(GA, accA);
        }
#pragma unroll
        for (int k = 0; k < 5; k++) r[k] = accA[k];
        r[5] = accT; r[6] = accNC0; r[7] = accNC1;
        unsigned cnt = nconf;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = HAS_A ? 0 : 5; k < 8; k++) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        {   // every lane holds all eight sums after the butterfly: lane k flushes sum k (one bin_add deep, not eight)
            double mine = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) if (lane == k) mine = r[k];
            if (lane < 5) { if (HAS_A) bin_add(acc, lane, a, mine, 0); }
            else if (lane < 8) bin_add(acc, SCAL, lane == 5 ? S_TOTAL : (lane == 6 ? S_NC0 : S_NC1), mine, 0);
            if (lane == 0) count_add(acc, (u64)cnt);
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) flag_set(acc, ERR_NOT_PD);
    }
    __syncwarp();
}

// Size class 1: one lane per union SNP of a 32-wide tile; three expansions at most.  Mantissa/exponent
// arithmetic straight into the bins (U subsets in total: cost is irrelevant).
__device__ inline void singles_tile(const LocusDev& L, int tile, int x_lo, int x_hi, int lane) {
    const AccDev& acc = L.acc;
    const int x = tile * 32 + lane;
    unsigned cnt = 0;
    if (x >= x_lo && x < x_hi) {
        const int l0 = L.loc[0][x], l1 = L.loc[1][x];
        const double p0 = L.pi[1][0], p1 = L.pi[1][1];
        E2 e0{1.0, 0}, e1{1.0, 0};
        if (l0 >= 0) e0 = E2{L.st[0].e1m[l0], L.st[0].e1n[l0]};
        if (l1 >= 0) e1 = E2{L.st[1].e1m[l1], L.st[1].e1n[l1]};
        if (l0 >= 0) {   // causal in study 0 only
            bin_add(acc, X1, x, p0 * e0.m, e0.n); bin_add(acc, YN, x, e0.m, e0.n);
            bin_add(acc, SCAL, S_TOTAL, p0 * e0.m, e0.n); bin_add(acc, SCAL, S_NC1, p0 * e0.m, e0.n);
            cnt++;
        }
        if (l1 >= 0) {   // causal in study 1 only
            bin_add(acc, X2, x, p0 * e1.m, e1.n); bin_add(acc, YN, x, e1.m, e1.n);
            bin_add(acc, SCAL, S_TOTAL, p0 * e1.m, e1.n); bin_add(acc, SCAL, S_NC0, p0 * e1.m, e1.n);
            cnt++;
        }
        if (l0 >= 0 && l1 >= 0) {   // causal in both
            const double m = e0.m * e1.m;
            const int n = e0.n + e1.n;
            bin_add(acc, X3, x, p1 * m, n); bin_add(acc, YS, x, m, n);
            bin_add(acc, SCAL, S_TOTAL, p1 * m, n);
            cnt++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) count_add(acc, (u64)cnt);
}

// All size classes 0..3 of one pipsort_run_exhaustive call in ONE launch: a single work queue of chunk descriptors
// (exh_plan.h), ordered from expensive (triples) to cheap (pairs, tiles of singles, the null configuration).
struct ExhAll {
    ExhParams p3, p2;
    const int4* chunks;    // [n_total] ExhChunkDesc
    int x1_lo, x1_hi;      // size class 1: SNPs [x1_lo, x1_hi)
    unsigned n_total;
    unsigned* counter;     // work-queue head; [1]: blocks that have finished (the last one re-arms the queue)
};

__global__ void __launch_bounds__(EXH_WARPS * 32, EXH_MINBLOCKS)
exhaustive_all_kernel(LocusDev L, ExhAll A, const LocusDev* __restrict__ Lg) {
    extern __shared__ __align__(16) unsigned char exh_smem[];      // EXH_WARPS x WarpWin (more than the 48 KB static limit)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpWin& win = reinterpret_cast<WarpWin*>(exh_smem)[wib];
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(A.counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= A.n_total) break;
        const int4 d = __ldg(A.chunks + item);
        const unsigned kind = (unsigned)d.w >> 28;
        const int nsteps = d.w & 0x0fffffff, xt = d.z & 0xffff, t_lo = (unsigned)d.z >> 16;
        if (kind == 3) {
            exh_chunk<3>(L, A.p3, d.x, d.y, xt, t_lo, nsteps, win, Lg, lane);
        } else if (kind == 2) {
            exh_chunk<2>(L, A.p2, -1, d.y, xt, t_lo, nsteps, win, Lg, lane);
        } else if (kind == 1) {
            singles_tile(L, d.x, A.x1_lo, A.x1_hi, lane);
        } else if (lane == 0) {   // postcal.cpp:793-822: -K/2 - 1 + U log(1-gamma)
            const double einv = 0.36787944117144233;
            bin_add(L.acc, SCAL, S_TOTAL, einv, 0);
            bin_add(L.acc, SCAL, S_NC0, einv, 0);
            bin_add(L.acc, SCAL, S_NC1, einv, 0);
            count_add(L.acc, 1ull);
        }
    }
    // this block has no more work: a dependent launch (finalize, launched with programmatic stream serialization) may
    // start becoming resident; it still waits for the whole grid before it reads the accumulators
    asm volatile("griddepcontrol.launch_dependents;");
    // the last block to get here re-arms the queue for the next launch (no memset node between passes)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(A.counter + 1, 1u);
        if (done == gridDim.x - 1) { A.counter[0] = 0u; A.counter[1] = 0u; }
    }
}

static_assert(sizeof(WarpWin) % 16 == 0 && (EXH_STG_LD * 8) % 16 == 0, "16-byte loads of the staging rows");
constexpr size_t EXH_SMEM_BYTES = sizeof(WarpWin) * EXH_WARPS;

}  // namespace pipsort
