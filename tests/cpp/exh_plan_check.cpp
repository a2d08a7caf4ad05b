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

static int check(int U, int J, int a_lo, int a_hi, double target) {
    ExhCost cs;
    std::vector<ExhChunkDesc> ch;
    exh_plan_class(U, J, a_lo, a_hi, target, cs, ch);
    const int off = exh_tile_off(U), T1 = exh_last_tile(U, off);
    std::map<std::tuple<int, int, int>, int> steps;          // (a, b, xt) -> visits
    std::map<std::tuple<int, int, int>, int> subsets;        // (a, b, x) -> visits
    for (const ExhChunkDesc& d : ch) {
        if ((int)(d.nsteps_kind >> 28) != J) { printf("kind mismatch\n"); return 1; }
        int a = J == 3 ? d.a : -1, b0 = d.b0, xt = d.xt_tlo & 0xffff, t_lo = d.xt_tlo >> 16;
        long remaining = d.nsteps_kind & 0x0fffffff;
        if (remaining <= 0) { printf("empty chunk\n"); return 1; }
        while (remaining > 0) {                                // === the kernel's walk ===
            if (b0 > U - 2 || xt > T1 || (J == 3 && (a < a_lo || a > a_hi))) { printf("walk left the class U=%d J=%d\n", U, J); return 1; }
            const int nb = exh_nb(U, b0);
            const int tmax = std::min(nb, xt * 32 + 31 - off - b0);
            const int t_hi = (int)std::min<long>(tmax, t_lo + remaining);
            if (t_hi <= t_lo) { printf("empty segment U=%d J=%d a=%d b0=%d xt=%d t_lo=%d\n", U, J, a, b0, xt, t_lo); return 1; }
            for (int t = t_lo; t < t_hi; t++) {
                const int b = b0 + t;
                steps[{a, b, xt}]++;
                for (int lane = 0; lane < 32; lane++) {
                    const int x = xt * 32 + lane - off;
                    if (x >= 0 && x < U && x > b) subsets[{a, b, x}]++;
                }
            }
            remaining -= t_hi - t_lo;
            if (remaining > 0) {
                t_lo = 0; xt++;
                if (xt > T1) {
                    b0 += 32;
                    if (b0 > U - 2) { if (J == 3) { a++; b0 = a + 1; } else { printf("pairs walk ran off the end\n"); return 1; } }
                    xt = exh_first_tile(off, b0);
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
            for (int xt = exh_first_tile(off, b); xt <= T1; xt++) {
                want_steps++;
                auto it = steps.find({a, b, xt});
                if (it == steps.end() || it->second != 1) { printf("step (%d,%d,%d) visited %d times\n", a, b, xt, it == steps.end() ? 0 : it->second); return 1; }
            }
        }
    if ((long)subsets.size() != want_sub || (long)steps.size() != want_steps) { printf("extra work: %zu/%ld subsets %zu/%ld steps\n", subsets.size(), want_sub, steps.size(), want_steps); return 1; }
    if ((double)want_steps != exh_class_steps(U, J, a_lo, a_hi)) { printf("exh_class_steps wrong: %g vs %ld\n", exh_class_steps(U, J, a_lo, a_hi), want_steps); return 1; }
    return 0;
}

int main() {
    int n = 0;
    for (int U : {2, 3, 4, 5, 31, 32, 33, 34, 63, 64, 65, 70, 97, 105, 131})
        for (double target : {1.0, 2.0, 3.5, 7.0, 13.0, 24.7, 40.0, 200.0, 5000.0, 1e9}) {
            if (U >= 2 && check(U, 2, 0, 0, target)) return 1;
            if (U >= 3) {
                if (check(U, 3, 0, U - 3, target)) return 1;
                if (U > 40 && check(U, 3, 7, U / 2, target)) return 1;
                if (check(U, 3, U - 3, U - 3, target)) return 1;
                n++;
            }
        }
    printf("ok %d\n", n);
    return 0;
}
