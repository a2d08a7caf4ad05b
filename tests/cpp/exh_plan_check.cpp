// CPU check of the exhaustive work decomposition (pipsort_b200/csrc/exh_plan.h): replays every chunk with the walk the
// kernel performs (exhaustive_dev.cuh: exh_chunk) and verifies that the chunks tile the warp-step space of the class
// exactly once and that every (a, b, x) subset is visited by exactly one active lane.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <tuple>

#include "../../pipsort_b200/csrc/exh_plan.h"

using namespace pipsort;

static int check(const ExhCostModel& M, int J, int a_lo, int a_hi, double target) {
    ExhCost cs;
    std::vector<ExhChunkDesc> ch;
    const int U = M.U;
    const ExhTiles& T = M.tiles;
    const double modelled = exh_plan_class(M, J, a_lo, a_hi, target, cs, ch);
    {   // the modelled cost of the chunks' steps must add up to the class cost (set-up costs come on top)
        const double want = exh_class_steps(M, J, a_lo, a_hi, true);
        if (modelled < want * (1.0 - 1e-9)) { printf("modelled cost %g below the class cost %g (U=%d J=%d target=%g)\n", modelled, want, U, J, target); return 1; }
    }
    const int T1 = T.T - 1;
    std::map<std::tuple<int, int, int>, int> steps;          // (a, b, xt) -> visits
    std::map<std::tuple<int, int, int>, int> subsets;        // (a, b, x) -> visits
    for (const ExhChunkDesc& d : ch) {
        if ((int)(d.nsteps_kind >> 28) != J) { printf("kind mismatch\n"); return 1; }
        int a = J == 3 ? d.a : -1, b0 = d.b0, xt = d.xt_tlo & 0xffff, t_lo = d.xt_tlo >> 16;
        long remaining = d.nsteps_kind & 0x0fffffff;
        if (remaining <= 0) { printf("empty chunk\n"); return 1; }
        while (remaining > 0) {                                // === the kernel's walk (exh_chunk) ===
            if (b0 > U - 2 || xt > T1 || (J == 3 && (a < a_lo || a > a_hi))) { printf("walk left the class U=%d J=%d\n", U, J); return 1; }
            const int nb = exh_nb(U, b0);
            const int x_lo = T.lo[xt];
            const int tmax = std::min(nb, x_lo + 31 - b0);
            const int t_hi = (int)std::min<long>(tmax, t_lo + remaining);
            if (t_hi <= t_lo) { printf("empty segment U=%d J=%d a=%d b0=%d xt=%d t_lo=%d\n", U, J, a, b0, xt, t_lo); return 1; }
            for (int t = t_lo; t < t_hi; t++) {
                const int b = b0 + t;
                steps[{a, b, xt}]++;
                for (int lane = 0; lane < 32; lane++) {
                    const int x = x_lo + lane;
                    if (x >= T.vmin[xt] && x > b) {
                        if (x >= U) { printf("x out of range\n"); return 1; }
                        subsets[{a, b, x}]++;
                    }
                }
            }
            remaining -= t_hi - t_lo;
            if (remaining > 0) {
                t_lo = 0; xt++;
                if (xt > T1) {
                    b0 += 32;
                    if (b0 > U - 2) { if (J == 3) { a++; b0 = a + 1; } else { printf("pairs walk ran off the end\n"); return 1; } }
                    xt = T.of[b0 + 1];
                }
            }
        }
    }
    // expectation
    long want_sub = 0, want_steps = 0;
    for (int a = (J == 3 ? a_lo : -1); a <= (J == 3 ? a_hi : -1); a++)
        for (int b = a + 1; b <= U - 2; b++) {
            for (int x = b + 1; x < U; x++) {
                want_sub++;
                auto it = subsets.find({a, b, x});
                if (it == subsets.end() || it->second != 1) { printf("subset (%d,%d,%d) visited %d times (U=%d J=%d target=%g)\n", a, b, x, it == subsets.end() ? 0 : it->second, U, J, target); return 1; }
            }
            for (int xt = T.of[b + 1]; xt <= T1; xt++) {
                want_steps++;
                auto it = steps.find({a, b, xt});
                if (it == steps.end() || it->second != 1) { printf("step (%d,%d,%d) visited %d times\n", a, b, xt, it == steps.end() ? 0 : it->second); return 1; }
            }
        }
    if ((long)subsets.size() != want_sub || (long)steps.size() != want_steps) { printf("extra work: %zu/%ld subsets %zu/%ld steps\n", subsets.size(), want_sub, steps.size(), want_steps); return 1; }
    if ((double)want_steps != exh_class_steps(M, J, a_lo, a_hi)) { printf("exh_class_steps wrong: %g vs %ld\n", exh_class_steps(M, J, a_lo, a_hi), want_steps); return 1; }
    return 0;
}

int main() {
    int n = 0;
    for (int U : {2, 3, 4, 5, 31, 32, 33, 34, 63, 64, 65, 70, 97, 105, 131}) {
        const ExhCostModel M(U, nullptr);                      // one group of SNPs, every step costs the same
        for (double target : {1.0, 2.0, 3.5, 7.0, 13.0, 24.7, 40.0, 200.0, 5000.0, 1e9}) {
            if (U >= 2 && check(M, 2, 0, 0, target)) return 1;
            if (U >= 3) {
                if (check(M, 3, 0, U - 3, target)) return 1;
                if (U > 40 && check(M, 3, 7, U / 2, target)) return 1;
                if (check(M, 3, U - 3, U - 3, target)) return 1;
                n++;
            }
        }
    }
    // SNP types: tiles end at the type boundaries and steps of single-study tiles are cheaper -- grouped layouts (also with
    // groups smaller than a tile) and an ungrouped one (one group)
    for (int U : {5, 33, 64, 97, 131, 180})
        for (int layout = 0; layout < 4; layout++) {
            std::vector<int> types(U);
            for (int i = 0; i < U; i++)
                types[i] = layout == 0 ? (i < (2 * U) / 3 ? 0 : (i < (5 * U) / 6 ? 1 : 2))
                         : layout == 1 ? (i < U / 4 ? 1 : 2)
                         : layout == 2 ? (i < U - 7 ? 0 : (i < U - 3 ? 1 : 2))
                                       : (i * 7 + i / 3) % 3;
            const ExhCostModel M(U, types.data());
            for (double target : {1.0, 2.6, 7.0, 21.3, 200.0, 1e9}) {
                if (check(M, 2, 0, 0, target)) return 1;
                if (check(M, 3, 0, U - 3, target)) return 1;
                if (U > 40 && check(M, 3, 7, U / 2, target)) return 1;
                n++;
            }
        }
    printf("ok %d\n", n);
    return 0;
}
