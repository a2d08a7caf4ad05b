// The multi-GPU combine step over NVLink / NVSwitch peer memory -- instead of an NCCL all-reduce.
//
// One process per GPU.  Every engine owns a MAILBOX in device memory (cudaMalloc + CUDA IPC handle, mapped by every
// peer): one inbox slot per rank, each with the layout of the accumulator store, plus three control words.  After a
// rank's exhaustive launch
//   non-root:  p2p_push_kernel  copies its store into ITS slot of the root's mailbox with plain 16-byte stores that travel
//              over NVLink as posted writes (86 KB for the 150-SNP locus: nothing next to 900 GB/s, and no remote
//              read-modify-write: system-scope fp64 atomics over the link took ~20 us for the same job), fences, and
//              bumps the root's `arrivals` word;
//   root:      p2p_merge_kernel polls its LOCAL `arrivals` word until all peers of this epoch have arrived, adds the
//              peers' slots to its store, and writes `consumed = epoch` into every peer's control word (flow control: a
//              peer may only overwrite its slot for epoch k once the root has consumed epoch k-1; peers poll their
//              LOCAL word).
// Both are ordinary stream-ordered launches: no host synchronisation, no collective library; the epoch counters live in
// device memory so the launches have no per-step argument and replay from a CUDA graph.  Spins are bounded (about a
// minute of clock64) and raise ERR_P2P_TIMEOUT instead of hanging the device.
#pragma once
#include "common.cuh"

namespace pipsort {

typedef unsigned long long u64;

constexpr int P2P_MAX_WORLD = 16;
constexpr long long P2P_SPIN_CYCLES = 120000000000ll;   // ~60 s at 1.9 GHz: ranks may reach the combine step far apart

struct P2PPeers {
    u64* ctrl[P2P_MAX_WORLD];   // every rank's control words: [0] arrivals (root's is used), [1] consumed, [2] epoch (local)
    int world, root;
};

__device__ inline bool p2p_spin_ge(const volatile u64* p, u64 target) {
    const long long t0 = clock64();
    while (*p < target) {
        if (clock64() - t0 > P2P_SPIN_CYCLES) return false;
        __nanosleep(64);
    }
    return true;
}

__global__ void __launch_bounds__(256)
p2p_push_kernel(const double* __restrict__ store, size_t n, double* __restrict__ my_slot_at_root, u64* __restrict__ root_ctrl,
                u64* __restrict__ my_ctrl, unsigned* __restrict__ done, double* __restrict__ err_flag) {
    __shared__ int ok;
    // the epoch lives in device memory (control word 2, bumped by the last block) so that the launch has no per-step
    // argument and can be replayed from a CUDA graph
    if (threadIdx.x == 0) ok = p2p_spin_ge(my_ctrl + 1, *(volatile u64*)(my_ctrl + 2)) ? 1 : 0;   // consumed >= epoch - 1
    __syncthreads();
    if (ok) {
        const size_t n2 = n >> 1;                    // 16-byte posted writes (both buffers are 256-byte aligned)
        const double2* __restrict__ src = reinterpret_cast<const double2*>(store);
        double2* __restrict__ dst = reinterpret_cast<double2*>(my_slot_at_root);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) my_slot_at_root[n - 1] = store[n - 1];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) atomicAdd(err_flag, 1.0);
        const unsigned d = atomicAdd(done, 1u);
        if (d == gridDim.x - 1) {                    // last block of this rank: everything above is visible system-wide
            *done = 0;
            my_ctrl[2] += 1;
            __threadfence_system();
            atomicAdd_system(root_ctrl, 1ull);
        }
    }
}

// slots: the root's own mailbox, world slots of n doubles each (slot r is written by rank r)
__global__ void __launch_bounds__(256)
p2p_merge_kernel(double* __restrict__ store, const double* __restrict__ slots, size_t n, size_t slot_stride, u64* __restrict__ my_ctrl,
                 P2PPeers peers, int my_rank, unsigned* __restrict__ done, double* __restrict__ err_flag) {
    __shared__ int ok;
    __shared__ u64 epoch_s;
    if (threadIdx.x == 0) {
        epoch_s = *(volatile u64*)(my_ctrl + 2) + 1;
        ok = p2p_spin_ge(my_ctrl, epoch_s * (u64)(peers.world - 1)) ? 1 : 0;    // every peer of this epoch has arrived
    }
    __syncthreads();
    const u64 epoch = epoch_s;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double v = store[i];
        for (int r = 0; r < peers.world; r++)
            if (r != my_rank) v += __ldcg(slots + (size_t)r * slot_stride + i);   // written over the link into L2: bypass L1
        store[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) atomicAdd(err_flag, 1.0);
        const unsigned d = atomicAdd(done, 1u);
        if (d == gridDim.x - 1) {
            *done = 0;
            my_ctrl[2] = epoch;
            __threadfence_system();
            for (int r = 0; r < peers.world; r++)
                if (r != my_rank) atomicExch_system(peers.ctrl[r] + 1, epoch);
        }
    }
}

}  // namespace pipsort
