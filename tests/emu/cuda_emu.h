// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a minimal host stand-in for the CUDA execution model, enough to run the plain
// (no tensor core, no TMA) kernels of csrc/ on the CPU so that their control flow and indexing can be checked without a
// GPU.  One OS thread per CUDA thread, one block at a time; __syncthreads is a std::barrier, warp shuffles exchange
// through a per-warp slot array between two warp barriers (every lane of the warp must call them, as on the device),
// __shared__ variables become function-local statics (valid because blocks run one after the other).
// Nothing in the product includes this file; the kernels under test are cut from the .cu sources by tests/test_emu.py.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct EmuDim { unsigned x = 1, y = 1, z = 1; };
static thread_local EmuDim threadIdx, blockIdx;
static EmuDim blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

static std::unique_ptr<std::barrier<>> emu_block_bar;
static std::vector<std::unique_ptr<std::barrier<>>> emu_warp_bar;
static double emu_slots[32][32];

inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }
inline double __shfl_xor_sync(unsigned, double v, int o) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    emu_slots[w][l] = v;
    emu_warp_bar[w]->arrive_and_wait();
    const double r = emu_slots[w][l ^ o];
    emu_warp_bar[w]->arrive_and_wait();
    return r;
}
inline double __ldg(const double* p) { return *p; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
inline int min(int a, int b) { return a < b ? a : b; }

namespace esper {
inline double warp_sum(double v) {                 // devutil.cuh: xor-shuffle tree, 16 down to 1
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
}  // namespace esper

// Run kernel(args...) on a grid of `grid` blocks of `threads` threads, one block after the other.
template <class K, class... A>
void emu_launch(int grid, int threads, K kernel, A... args) {
    gridDim.x = (unsigned)grid; blockDim.x = (unsigned)threads;
    const int warps = (threads + 31) / 32;
    for (int b = 0; b < grid; ++b) {
        emu_block_bar = std::make_unique<std::barrier<>>(threads);
        emu_warp_bar.clear();
        for (int w = 0; w < warps; ++w) emu_warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32, threads - 32 * w)));
        std::vector<std::thread> th;
        th.reserve(threads);
        for (int t = 0; t < threads; ++t)
            th.emplace_back([=]() { threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b; kernel(args...); });
        for (auto& x : th) x.join();
    }
}
