// Work decomposition of the exhaustive launch (host side, plain C++: also compiled by the CPU tests).
//
// The register kernel (exhaustive_dev.cuh) walks, for every size-3 union subset {a < b < x} (size 2: {b < x}), a fixed
// order of WARP-STEPS: a outermost, then 32-wide windows of b, then 32-wide tiles of x (lanes), then the b's of the window
// -- one step = one (a, b) against 32 x's.  A CHUNK is a contiguous run of that order, described by where it starts and
// how many steps it has; the kernel's work queue hands out chunks.  Cutting the order at arbitrary steps (not at window
// or tile boundaries) is what lets a small locus -- fewer steps than twenty per resident warp -- be dealt out evenly:
// one chunk per resident warp, all of (nearly) the same cost.  Large loci get a dozen chunks per resident warp.
//
// x tiles (ExhTiles) never straddle a boundary between SNP types when the internal order is grouped by type (both studies /
// study 0 only / study 1 only): a tile of one type runs a specialised, cheaper step loop.  Inside a group the tiles are
// aligned to the TOP of the group (the partly empty tile is the lowest one, which the fewest pairs visit).
//
// Replaces the static OpenMP schedule of postcal.cpp:769-770 (chunks of total/1000 ranks).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace pipsort {

struct ExhChunkDesc {          // 16 bytes, read by the kernel as one int4
    int32_t a;                 // kind 3: first SNP of the triple; kind 1: tile of singles; else unused
    int32_t b0;                // first b of the window the chunk starts in
    uint32_t xt_tlo;           // x tile | (first step inside the (window, tile) segment) << 16
    uint32_t nsteps_kind;      // number of warp-steps | kind << 28   (kind 3 triples, 2 pairs, 1 singles tile, 0 null)
};

struct ExhCost {               // relative cost of the set-up work, in warp-steps (measured, scripts/sweep_chunks.py)
    double seg = 0.6;          // per (window, tile) segment: the lane values of x, flush of the x cells
    double win = 0.9;          // per window: the table of the 32 b's, flush of the b cells
    double a = 0.5;            // per change of a: flush of the a cells
    double chunk = 0.8;        // per chunk: queue pop, final flushes
};

inline int exh_nb(int U, int b0) { return std::min(32, U - 1 - b0); }                       // b in [b0, min(b0 + 31, U - 2)]

// The 32-wide tiles of x.  Lane l of tile t is x = lo[t] + l, valid when x >= vmin[t] (the lowest tile of a group starts
// below its group); hi(t) = lo[t] + 31 is always a SNP of the tile.  `types` (per internal SNP, may be null): a new group
// starts wherever the type changes, provided the order is grouped (at most four runs); otherwise one group.
struct ExhTiles {
    int U = 0, T = 0;
    std::vector<int> lo, vmin, of;    // [T], [T], [U]: tile of x
    ExhTiles() {}
    ExhTiles(int U_, const int* types) : U(U_) {
        std::vector<int> cut{0};
        if (types) {
            for (int i = 1; i < U; i++) if (types[i] != types[i - 1]) cut.push_back(i);
            if (cut.size() > 4) cut.assign(1, 0);
        }
        cut.push_back(U);
        of.assign(U, 0);
        for (size_t g = 0; g + 1 < cut.size(); g++) {
            const int g0 = cut[g], g1 = cut[g + 1], nt = (g1 - g0 + 31) / 32;
            for (int j = 0; j < nt; j++) {
                const int l = g1 - 32 * (nt - j);
                for (int x = std::max(l, g0); x < l + 32; x++) of[x] = T;
                lo.push_back(l); vmin.push_back(std::max(l, g0));
                T++;
            }
        }
    }
    int hi(int t) const { return lo[t] + 31; }
    int first_tile(int b0) const { return of[b0 + 1]; }                                      // the tile that holds x = b0 + 1
    // steps of the segment (window at b0, tile xt): the b's of the window that have an x beyond them in the tile
    int seg_steps(int b0, int xt) const { return std::min(exh_nb(U, b0), hi(xt) - b0); }
};

// What a warp-step costs depends on which of the kernel's specialised step loops its segment runs (exhaustive_dev.cuh,
// run_steps): generic (the tile has x's of both studies), single-study tile with the study's bordered-step chain, or
// single-study tile without a chain.  The model restates the kernel's dispatch from the SNP types in the internal order
// (0: in both studies, 1: study 0 only, 2: study 1 only, 3: neither); with no types every step costs 1.
struct ExhCostModel {
    int U = 0;
    ExhTiles tiles;
    double c_chain = 0.70;     // single-study tile, chain        (relative to a generic size-3 step; measured, scripts/sweep_costs.py)
    double c_plain = 0.40;     // single-study tile, no chain
    double c_pair = 0.45, c_pair1 = 0.30;   // size-2 steps: generic / single-study tile
    std::vector<unsigned char> cls;         // per x tile: 0 generic, 1 no x of study 1, 2 no x of study 0
    std::vector<int> sufG, suf1, suf2;      // tiles of each class from tile t to the last one
    std::vector<int> pre1;                  // pre1[i] = number of SNPs < i that are in study 1
    std::vector<unsigned char> in1;         // SNP is in study 1
    uint64_t hash = 0;                      // of the type layout (plans are cached per layout)
    ExhCostModel() {}
    ExhCostModel(int U_, const int* types) : U(U_), tiles(U_, types) {
        const int nt = tiles.T;
        cls.assign(std::max(nt, 0), 0);
        sufG.assign(nt + 1, 0); suf1.assign(nt + 1, 0); suf2.assign(nt + 1, 0);
        pre1.assign(U + 1, 0); in1.assign(U, 1);
        if (types) {
            hash = 1469598103934665603ull;
            for (int i = 0; i < U; i++) {
                in1[i] = types[i] == 0 || types[i] == 2;
                hash = (hash ^ (uint64_t)(types[i] + 1)) * 1099511628211ull;
            }
            for (int t = 0; t < nt; t++) {
                bool h0 = false, h1 = false;
                for (int x = tiles.vmin[t]; x <= tiles.hi(t); x++) {
                    h0 |= types[x] == 0 || types[x] == 1;
                    h1 |= types[x] == 0 || types[x] == 2;
                }
                cls[t] = !h1 ? 1 : (!h0 ? 2 : 0);
            }
        }
        for (int i = 0; i < U; i++) pre1[i + 1] = pre1[i] + in1[i];
        for (int t = nt - 1; t >= 0; t--) {
            sufG[t] = sufG[t + 1] + (cls[t] == 0); suf1[t] = suf1[t + 1] + (cls[t] == 1); suf2[t] = suf2[t + 1] + (cls[t] == 2);
        }
    }
    // study 1 can have a non-zero E{a,b,x} in the window: a is in it and so is some b of the window (kernel: ha[1] && win_has1)
    bool chain1(int J, int a, int b0, int nb) const { return J == 3 && a >= 0 && in1[a] && pre1[b0 + nb] - pre1[b0] > 0; }
    double w_class(int J, int c, bool ch1) const {
        if (J == 2) return c == 0 ? c_pair : c_pair1;
        return c == 0 ? 1.0 : (c == 1 ? c_chain : (ch1 ? c_chain : c_plain));
    }
    double step_cost(int J, int a, int b0, int nb, int xt) const { return w_class(J, cls[xt], chain1(J, a, b0, nb)); }
    // cost of ONE step in each of the tiles xt..T-1 of the window (to be multiplied by the steps per tile)
    double tiles_cost(int J, int a, int b0, int nb, int xt) const {
        const bool ch1 = chain1(J, a, b0, nb);
        return sufG[xt] * w_class(J, 0, ch1) + suf1[xt] * w_class(J, 1, ch1) + suf2[xt] * w_class(J, 2, ch1);
    }
    // the tiles xt.. of the window at b0: steps (n) and cost (c) of all of them.  Every tile beyond the first few has all nb
    // steps (a tile has fewer only while its top x lies inside the window's own range of b)
    void window_rest(int J, int a, int b0, int nb, int xt, double& n, double& c) const {
        n = 0; c = 0;
        int t = xt;
        for (; t < tiles.T && tiles.hi(t) - b0 < nb; t++) { const int k = tiles.seg_steps(b0, t); n += k; c += k * step_cost(J, a, b0, nb, t); }
        if (t < tiles.T) { n += (double)(tiles.T - t) * nb; c += nb * tiles_cost(J, a, b0, nb, t); }
    }
};

// Total steps of class J (3: a in [a_lo, a_hi]; 2: a ignored) -- O(U^2 / 32) arithmetic; `cost`: their modelled cost instead.
inline double exh_class_steps(const ExhCostModel& M, int J, int a_lo, int a_hi, bool cost = false) {
    const int U = M.U;
    auto of_a = [&](int a) {
        double tot = 0;
        for (int b0 = a + 1; b0 <= U - 2; b0 += 32) {
            double n, c;
            M.window_rest(J, a, b0, exh_nb(U, b0), M.tiles.first_tile(b0), n, c);
            tot += cost ? c : n;
        }
        return tot;
    };
    if (J == 2) return of_a(-1);
    double n = 0;
    for (int a = a_lo; a <= a_hi; a++) n += of_a(a);
    return n;
}

// Cuts class J into chunks of cost ~target (steps + set-up costs) and appends them.  Returns the modelled total cost.
inline double exh_plan_class(const ExhCostModel& M, int J, int a_lo, int a_hi, double target, const ExhCost& cs, std::vector<ExhChunkDesc>& out) {
    const int U = M.U, T1 = M.tiles.T - 1;
    if (U < J || (J == 3 && a_lo > a_hi)) return 0.0;
    int a = J == 3 ? a_lo : -1, b0 = a + 1, t = 0;
    if (b0 > U - 2) return 0.0;
    int xt = M.tiles.first_tile(b0);
    double total = 0.0;
    bool done = false;
    while (!done) {
        ExhChunkDesc d{a, b0, (uint32_t)xt | ((uint32_t)t << 16), 0};
        double acc = cs.chunk + cs.win + cs.seg + (J == 3 ? cs.a : 0.0);
        uint32_t nsteps = 0;
        for (;;) {
            const int nb = exh_nb(U, b0);
            int left = M.tiles.seg_steps(b0, xt) - t;                     // steps left in this segment
            const double w = M.step_cost(J, a, b0, nb, xt);               // cost of a step of this segment
            // whole rest of the window at once when it fits (large targets: avoids walking every tile on the host)
            if (t == 0) {
                double wn, wc;
                M.window_rest(J, a, b0, nb, xt, wn, wc);
                const double wcost = wc + (double)(T1 - xt) * cs.seg;
                if (acc + wcost <= target && nsteps + (uint64_t)wn < (1u << 27)) {
                    nsteps += (uint32_t)wn;
                    acc += wcost;
                    xt = T1;
                    left = 0;
                    t = M.tiles.seg_steps(b0, xt);
                }
            }
            if (left > 0) {
                int take = left;
                const double room = (target - acc) / w;
                if (room < left) take = std::max(room >= 1.0 ? (int)room : 0, nsteps == 0 ? 1 : 0);
                if ((uint64_t)nsteps + take >= (1u << 27)) take = 0;
                nsteps += take; acc += take * w; t += take; left -= take;
                if (left > 0) break;                                       // chunk is full; the next one resumes here
            }
            // advance to the next segment
            t = 0;
            xt++;
            double adv = cs.seg;
            if (xt > T1) {
                b0 += 32;
                adv += cs.win;
                if (b0 > U - 2) {
                    if (J == 3 && a < a_hi) { a++; b0 = a + 1; adv += cs.a; }
                    else { done = true; break; }
                }
                xt = M.tiles.first_tile(b0);
            }
            if (acc + adv + M.c_plain > target) break;                     // no room for another step: close the chunk here
            acc += adv;
        }
        if (nsteps) {
            d.nsteps_kind = nsteps | ((uint32_t)J << 28);
            out.push_back(d);
            total += acc;
        }
    }
    return total;
}

}  // namespace pipsort
