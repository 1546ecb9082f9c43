// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a minimal host stand-in for the CUDA execution model, enough to run the plain
// (no tensor core, no TMA) kernels of csrc/ on the CPU so that their control flow and indexing can be checked without a
// GPU.  One OS thread per CUDA thread, one block at a time; __syncthreads is a std::barrier, warp shuffles exchange
// through a per-warp slot array between two warp barriers (every lane of the warp must call them, as on the device),
// __shared__ variables become function-local statics (valid because blocks run one after the other).
// Nothing in the product includes this file; the kernels under test are cut from the .cu sources by tests/test_emu.py.
#pragma once
#include <atomic>
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
#define __align__(x)
#ifndef EMU_ARENA_SHARED
#define __shared__ static       // one block at a time: a function-local static is the block's shared variable
#endif

static std::unique_ptr<std::barrier<>> emu_block_bar;
static std::vector<std::unique_ptr<std::barrier<>>> emu_warp_bar;
static double emu_slots[32][32];

#ifndef EMU_ARENA_SHARED
inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }
inline double __shfl_xor_sync(unsigned, double v, int o) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    emu_slots[w][l] = v;
    emu_warp_bar[w]->arrive_and_wait();
    const double r = emu_slots[w][l ^ o];
    emu_warp_bar[w]->arrive_and_wait();
    return r;
}
#else
namespace emu2 { inline void syncthreads(); inline double shfl_xor(double v, int o); inline int syncthreads_or(int v); inline void syncwarp(); }
inline void __syncthreads() { emu2::syncthreads(); }
inline int __syncthreads_or(int v);
inline void __syncwarp();
inline double __shfl_xor_sync(unsigned, double v, int o) { return emu2::shfl_xor(v, o); }
#endif
inline double __ldg(const double* p) { return *p; }
inline float __ldg(const float* p) { return *p; }
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float4 __ldg(const float4* p) { return *p; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline uint32_t emu_f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float emu_u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// IEEE binary16, round to nearest even (what __float2half_rn does), subnormals and overflow to infinity included
struct __half { uint16_t bits; };
inline __half __float2half_rn(float f) {
    const uint32_t u = emu_f2u(f), sign = (u >> 16) & 0x8000u;
    const int32_t e = (int32_t)((u >> 23) & 0xFF) - 127;
    uint32_t man = u & 0x7FFFFFu;
    __half h;
    if (((u >> 23) & 0xFF) == 0xFF) { h.bits = (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0)); return h; }
    if (e > 15) { h.bits = (uint16_t)(sign | 0x7C00u); return h; }
    if (e >= -14) {                                        // normal half
        uint32_t m10 = man >> 13, rest = man & 0x1FFFu, he = (uint32_t)(e + 15);
        uint32_t v = (he << 10) | m10;
        if (rest > 0x1000u || (rest == 0x1000u && (m10 & 1))) ++v;       // carries into the exponent correctly
        h.bits = (uint16_t)(sign | v);
        return h;
    }
    if (e < -25) { h.bits = (uint16_t)sign; return h; }      // below half the smallest subnormal
    man |= 0x800000u;                                        // subnormal: shift the 24-bit significand
    const int shift = -14 - e + 13;
    uint32_t m = man >> shift, rest = man & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rest > half || (rest == half && (m & 1))) ++m;
    h.bits = (uint16_t)(sign | m);
    return h;
}
inline float __half2float(__half h) {
    const uint32_t sign = (uint32_t)(h.bits & 0x8000u) << 16, e = (h.bits >> 10) & 0x1Fu, m = h.bits & 0x3FFu;
    if (e == 0x1F) return emu_u2f(sign | 0x7F800000u | (m << 13));
    if (e == 0) { const float v = std::ldexp((float)m, -24); return sign ? -v : v; }
    return emu_u2f(sign | ((e + 112) << 23) | (m << 13));
}
inline unsigned short __half_as_ushort(__half h) { return h.bits; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
struct uint4 { uint32_t x, y, z, w; };
inline double __ldcg(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
inline long long clock64() { return 0; }                                  // no watchdog in the stand-in
inline unsigned int atomicExch(unsigned int* p, unsigned int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline void emu_red_release_add(unsigned int* p) { __atomic_fetch_add(p, 1u, __ATOMIC_RELEASE); }
inline unsigned int emu_ld_acquire(const unsigned int* p) { std::this_thread::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

#ifndef EMU_NO_WARP_SUM
namespace esper {
inline double warp_sum(double v) {                 // devutil.cuh: xor-shuffle tree, 16 down to 1
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
}  // namespace esper
#endif

// ---- second model (EMU_ARENA_SHARED): blocks of a thread-block cluster run CONCURRENTLY, so shared memory is a per-block
// arena; the test rewrites `__shared__ T name[N];` into `T* name = emu_smem<T>(N);` (same allocation order in every thread ->
// same offsets) and `extern __shared__ T name[];` into `T* name = emu_dyn<T>();`.  map_shared_rank translates an address of
// this block's arena into the peer's, which is what distributed shared memory does.
#ifndef EMU_STATIC_BYTES_DEF
#define EMU_STATIC_BYTES_DEF (16 * 1024)
#endif
constexpr size_t EMU_STATIC_BYTES = EMU_STATIC_BYTES_DEF;
static thread_local char* emu_arena = nullptr;
static thread_local size_t emu_off = 0;
static std::vector<std::vector<char>> emu_arenas;          // one per block of the running cluster
static thread_local unsigned emu_cluster_rank = 0;
static unsigned emu_cluster_size = 1;
static std::unique_ptr<std::barrier<>> emu_cluster_bar;
static std::vector<std::unique_ptr<std::barrier<>>> emu_block_bars;
static thread_local std::barrier<>* emu_my_block_bar = nullptr;
static thread_local std::atomic<int>* emu_my_or = nullptr;     // [2] per block: __syncthreads_or
static thread_local int emu_or_parity = 0;
static thread_local std::barrier<>* emu_my_warp_bar = nullptr;
static thread_local double* emu_my_slots = nullptr;

template <class T> T* emu_smem(size_t n) {
    emu_off = (emu_off + 15) & ~size_t(15);
    T* p = reinterpret_cast<T*>(emu_arena + emu_off);
    emu_off += n * sizeof(T);
    if (emu_off > EMU_STATIC_BYTES) { fprintf(stderr, "emu: static shared memory exhausted\n"); abort(); }
    return p;
}
template <class T> T* emu_dyn() { return reinterpret_cast<T*>(emu_arena + EMU_STATIC_BYTES); }

#ifdef EMU_ARENA_SHARED
#undef EMU_SYNC_DEFINED
namespace emu2 {
inline void syncthreads() { emu_my_block_bar->arrive_and_wait(); }
inline int syncthreads_or(int v) {
    const int p = emu_or_parity; emu_or_parity ^= 1;
    if (v) emu_my_or[p].store(1);
    emu_my_block_bar->arrive_and_wait();
    const int r = emu_my_or[p].load();
    emu_my_block_bar->arrive_and_wait();
    if (threadIdx.x == 0) emu_my_or[p].store(0);
    return r;
}
inline void syncwarp() { emu_my_warp_bar->arrive_and_wait(); }
inline double shfl_xor(double v, int o) {
    const int l = threadIdx.x & 31;
    emu_my_slots[l] = v;
    emu_my_warp_bar->arrive_and_wait();
    const double r = emu_my_slots[l ^ o];
    emu_my_warp_bar->arrive_and_wait();
    return r;
}
}  // namespace emu2
inline int __syncthreads_or(int v) { return emu2::syncthreads_or(v); }
inline void __syncwarp() { emu2::syncwarp(); }
namespace cooperative_groups {
struct cluster_group {
    unsigned num_blocks() const { return emu_cluster_size; }
    unsigned block_rank() const { return emu_cluster_rank; }
    void sync() const { emu_cluster_bar->arrive_and_wait(); }
    template <class T> T* map_shared_rank(T* p, unsigned r) const {
        return reinterpret_cast<T*>(emu_arenas[r].data() + (reinterpret_cast<char*>(p) - emu_arena));
    }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups

// grid blocks of `threads` threads in clusters of `cs` consecutive blocks (cs = 1: plain launch); dyn = dynamic shared bytes.
template <class K, class... A>
void emu_launch_cluster(int grid, int threads, int cs, size_t dyn, K kernel, A... args) {
    gridDim.x = (unsigned)grid; blockDim.x = (unsigned)threads;
    emu_cluster_size = (unsigned)cs;
    const int warps = (threads + 31) / 32;
    static std::vector<double> slots;
    for (int c0 = 0; c0 < grid; c0 += cs) {
        emu_arenas.assign(cs, std::vector<char>(EMU_STATIC_BYTES + dyn + 64, 0));
        emu_cluster_bar = std::make_unique<std::barrier<>>(cs * threads);
        emu_block_bars.clear();
        std::vector<std::unique_ptr<std::barrier<>>> wbars;
        std::vector<std::atomic<int>> orflags(2 * (size_t)cs);
        for (auto& f : orflags) f.store(0);
        slots.assign((size_t)cs * warps * 32, 0.0);
        for (int r = 0; r < cs; ++r) {
            emu_block_bars.push_back(std::make_unique<std::barrier<>>(threads));
            for (int w = 0; w < warps; ++w) wbars.push_back(std::make_unique<std::barrier<>>(std::min(32, threads - 32 * w)));
        }
        std::vector<std::thread> th;
        for (int r = 0; r < cs; ++r)
            for (int t = 0; t < threads; ++t)
                th.emplace_back([=, &wbars, &orflags]() {
                    threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)(c0 + r);
                    emu_my_or = orflags.data() + 2 * (size_t)r; emu_or_parity = 0;
                    emu_cluster_rank = (unsigned)r;
                    emu_arena = emu_arenas[r].data(); emu_off = 0;
                    emu_my_block_bar = emu_block_bars[r].get();
                    emu_my_warp_bar = wbars[(size_t)r * warps + (t >> 5)].get();
                    emu_my_slots = slots.data() + ((size_t)r * warps + (t >> 5)) * 32;
                    kernel(args...);
                });
        for (auto& x : th) x.join();
    }
}
#endif

#ifndef EMU_ARENA_SHARED
// 2-D grid of kernels without __syncthreads / shuffles (epilogue kernels): the threads of a block run one after the other.
template <class K, class... A>
void emu_launch_serial(int gx, int gy, int threads, K kernel, A... args) {
    gridDim.x = (unsigned)gx; gridDim.y = (unsigned)gy; blockDim.x = (unsigned)threads;
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx)
            for (int t = 0; t < threads; ++t) { threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)bx; blockIdx.y = (unsigned)by; kernel(args...); }
    gridDim.y = 1;
}
#endif

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
