// placeholder until the register kernel lands
#pragma once
#include "common.cuh"
namespace pipsort {
inline int exhaustive_launch(const LocusDev&, int, unsigned long long, unsigned long long, int, cudaStream_t, bool* done,
                             unsigned long long*, int*) { *done = false; return 0; }
}
