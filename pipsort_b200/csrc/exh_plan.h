// Work decomposition of the exhaustive launch (host side, plain C++: also compiled by the CPU tests).
//
// The register kernel (exhaustive_dev.cuh) walks, for every size-3 union subset {a < b < x} (size 2: {b < x}), a fixed
// order of WARP-STEPS: a outermost, then 32-wide windows of b, then 32-wide tiles of x (lanes), then the b's of the window
// -- one step = one (a, b) against 32 x's.  A CHUNK is a contiguous run of that order, described by where it starts and
// how many steps it has; the kernel's work queue hands out chunks.  Cutting the order at arbitrary steps (not at window
// or tile boundaries) is what lets a small locus -- fewer steps than twenty per resident warp -- be dealt out evenly:
// one chunk per resident warp, all of (nearly) the same cost.  Large loci get a dozen chunks per resident warp.
//
// x tiles are aligned to the TOP of the SNP range (tile j covers x in [32 j - off, 32 j - off + 32), off = 32 T - U):
// the partly empty tile is then the lowest one, which only the few pairs with b < 32 - off ever visit, instead of the
// highest one, which every pair visits.
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

inline int exh_tile_off(int U) { return U > 0 ? (32 - (U & 31)) & 31 : 0; }
inline int exh_nb(int U, int b0) { return std::min(32, U - 1 - b0); }                       // b in [b0, min(b0 + 31, U - 2)]
inline int exh_first_tile(int off, int b0) { return (b0 + 1 + off) >> 5; }                  // the tile that holds x = b0 + 1
inline int exh_last_tile(int U, int off) { return (U - 1 + off) >> 5; }
// steps of the segment (window at b0, tile xt): the b's of the window that have an x beyond them in the tile
inline int exh_seg_steps(int U, int off, int b0, int xt) { return std::min(exh_nb(U, b0), xt * 32 + 31 - off - b0); }

// Total steps of class J (3: a in [a_lo, a_hi]; 2: a ignored) -- O(U^2 / 32) arithmetic.
inline double exh_class_steps(int U, int J, int a_lo, int a_hi) {
    const int off = exh_tile_off(U), T1 = exh_last_tile(U, off);
    auto of_a = [&](int a) {
        double n = 0;
        for (int b0 = a + 1; b0 <= U - 2; b0 += 32) {
            const int nb = exh_nb(U, b0), t0 = exh_first_tile(off, b0);
            n += exh_seg_steps(U, off, b0, t0) + (double)(T1 - t0) * nb;
        }
        return n;
    };
    if (J == 2) return of_a(-1);
    double n = 0;
    for (int a = a_lo; a <= a_hi; a++) n += of_a(a);
    return n;
}

// Cuts class J into chunks of cost ~target (steps + set-up costs) and appends them.  Returns the modelled total cost.
inline double exh_plan_class(int U, int J, int a_lo, int a_hi, double target, const ExhCost& cs, std::vector<ExhChunkDesc>& out) {
    const int off = exh_tile_off(U), T1 = exh_last_tile(U, off);
    if (U < J || (J == 3 && a_lo > a_hi)) return 0.0;
    int a = J == 3 ? a_lo : -1, b0 = a + 1, xt = exh_first_tile(off, b0), t = 0;
    if (b0 > U - 2) return 0.0;
    double total = 0.0;
    bool done = false;
    while (!done) {
        ExhChunkDesc d{a, b0, (uint32_t)xt | ((uint32_t)t << 16), 0};
        double acc = cs.chunk + cs.win + cs.seg + (J == 3 ? cs.a : 0.0);
        uint32_t nsteps = 0;
        for (;;) {
            const int nb = exh_nb(U, b0);
            int left = exh_seg_steps(U, off, b0, xt) - t;                 // steps left in this segment
            // whole rest of the window at once when it fits (large targets: avoids walking every tile on the host)
            if (t == 0) {
                const double wcost = (double)left + (double)(T1 - xt) * (nb + cs.seg);
                if (acc + wcost <= target && nsteps + (uint64_t)left + (uint64_t)(T1 - xt) * nb < (1u << 27)) {
                    nsteps += left + (T1 - xt) * nb;
                    acc += wcost;
                    xt = T1;
                    left = 0;
                    t = exh_seg_steps(U, off, b0, xt);
                }
            }
            if (left > 0) {
                int take = left;
                const double room = target - acc;
                if (room < left) take = std::max(room >= 1.0 ? (int)room : 0, nsteps == 0 ? 1 : 0);
                if ((uint64_t)nsteps + take >= (1u << 27)) take = 0;
                nsteps += take; acc += take; t += take; left -= take;
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
                xt = exh_first_tile(off, b0);
            }
            if (acc + adv + 1.0 > target) break;                           // no room for another step: close the chunk here
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
