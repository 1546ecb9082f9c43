// ll_microbench.cu -- measurement tool (not part of the library): what a CTA-to-CTA hand-over through L2 costs on this GPU.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/ll_microbench tools/ll_microbench.cu
// Tests (cooperative launch, one CTA per SM, all co-resident):
//   1 ping-pong between CTA 0 and CTA k with flag-carrying 16-byte packets: cycles per one-way hand-over, for several
//     store / load flavours (volatile, relaxed.gpu, release/acquire.gpu) and for k on the same / the other die
//   2 __nanosleep(t) real duration
//   3 all-gather pattern of the Lanczos vector: every CTA publishes 14 packets, every CTA reads all packets (P4)
//   4 reduce + broadcast pattern (P1-P3): every CTA publishes S partials, slot owners add them, everyone reads the S sums
//   5 plain grid barrier (red.release / ld.acquire)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ unsigned g_timeout;
#define SPIN_GUARD(t0) if (*((volatile unsigned*)&g_timeout) != 0u) return; if (clock64() - (t0) > 400000000LL) { g_timeout = 1; return; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int F> __device__ __forceinline__ void st_pk(uint4* p, unsigned lo, unsigned hi, unsigned tag) {
    if (F == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
    if (F == 1) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
    if (F == 2) asm volatile("st.release.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
    if (F == 3) asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
template <int F> __device__ __forceinline__ uint4 ld_pk(const uint4* p) {
    uint4 v;
    if (F == 0) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <int F>
__global__ void pingpong_kernel(uint4* buf, int peer, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    const int b = blockIdx.x;
    if (b != 0 && b != peer) return;
    uint4* mine = buf + (b == 0 ? 0 : 8);       // different 128-byte lines
    uint4* theirs = buf + (b == 0 ? 8 : 0);
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        if (b == 0) {
            st_pk<F>(theirs, it, it, it);
            uint4 v;
            { const long long tw_ = clock64(); do { v = ld_pk<F>(mine); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
        } else {
            uint4 v;
            { const long long tw_ = clock64(); do { v = ld_pk<F>(mine); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
            st_pk<F>(theirs, it, it, it);
        }
    }
    if (b == 0) out[0] = clock64() - t0;
}

__global__ void nanosleep_kernel(long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned ts[6] = {0, 20, 40, 100, 200, 1000};
    for (int k = 0; k < 6; ++k) {
        const long long t0 = clock64();
        for (int i = 0; i < 100; ++i) __nanosleep(ts[k]);
        out[k] = (clock64() - t0) / 100;
    }
}

// P4 pattern: every CTA (nb of them) writes `per` packets of its own, then reads all nb*per packets.  poll_mode 0: every thread
// polls its packets directly; 1: sentinel per producer first (its last packet), then one bulk read.
template <int F>
__global__ void __launch_bounds__(512, 1) allgather_kernel(uint4* buf, int per, int iters, int poll_mode, long long* out, double* sink) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    const int n = nb * per;
    double acc = 0.0;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        if (tid < per) st_pk<F>(buf + b * per + tid, it + tid, b, it);
        if (poll_mode == 1) {
            for (int c = tid; c < nb; c += blockDim.x) {
                uint4 v;
                { const long long tw_ = clock64(); do { v = ld_pk<F>(buf + c * per + per - 1); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
            }
            __syncthreads();
        }
        for (int c = tid; c < n; c += blockDim.x) {
            uint4 v;
            { const long long tw_ = clock64(); do { v = ld_pk<F>(buf + c); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
            acc += (double)v.x;
        }
        __syncthreads();
    }
    if (tid == 0) { out[b] = clock64() - t0; }
    if (acc == 1.2345) sink[0] = acc;
}

// P1-P3 pattern: S slots; every CTA writes S partial packets pbuf[s][b]; owner of slot s (CTA s % nb, warp s / nb) sums the nb
// partials and writes hbuf[s]; warps 8..11 of every CTA read all S sums.
template <int F>
__global__ void __launch_bounds__(512, 1) reduce_bcast_kernel(uint4* pbuf, uint4* hbuf, int S, int iters, long long* out, double* sink) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ double s_h[512];
    double acc = 0.0;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        for (int s = tid; s < S; s += blockDim.x) st_pk<F>(pbuf + (size_t)s * nb + b, s + it, b, it);
        for (int s = b + warp * nb; s < S && warp < 8; s += 8 * nb) {
            unsigned tot = 0;
            for (int c = lane; c < nb; c += 32) {
                uint4 v;
                { const long long tw_ = clock64(); do { v = ld_pk<F>(pbuf + (size_t)s * nb + c); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
                tot += v.x;
            }
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) st_pk<F>(hbuf + s, tot, 0, it);
        }
        if (warp >= 8 && warp < 12) {
            for (int s = tid - 256; s < S; s += 128) {
                uint4 v;
                { const long long tw_ = clock64(); do { v = ld_pk<F>(hbuf + s); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
                s_h[s] = (double)v.x;
            }
        }
        __syncthreads();
        acc += s_h[tid % S];
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (acc == 1.2345) sink[0] = acc;
}


// ---- plain data + one flag per producer -----------------------------------------------------------------------------
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_flag(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double2 ld_d2(const double* p) { double2 v; asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_d(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// A2: every CTA writes `per` doubles (plain), fence, flag; warp 0 of every CTA polls all flags, then all threads load the vector.
// variant 1: flags are 128 bytes apart (one line per producer) instead of packed.
__global__ void __launch_bounds__(512, 1) allgather_flag_kernel(double* vec, unsigned* flags, int per, int iters, int variant, long long* out, double* sink) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = nb * per, fstride = variant == 1 ? 32 : 1;
    double acc = 0.0;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* dst = vec + (size_t)(it & 1) * n;                      // two buffers by parity
        if (warp == 0) {
            if (lane < per) st_d(dst + b * per + lane, (double)(it + lane));
            __syncwarp();
            if (lane == 0) { __threadfence(); st_flag(flags + (size_t)b * fstride, (unsigned)it); }
            // poll
            const long long tw_ = clock64();
            for (int c = lane; c < nb; c += 32) {
                unsigned v;
                do { v = ld_flag(flags + (size_t)c * fstride); SPIN_GUARD(tw_) } while ((int)(v - (unsigned)it) < 0);
            }
            __threadfence();
        }
        __syncthreads();
        for (int c = 2 * tid; c < n; c += 2 * blockDim.x) { const double2 v = ld_d2(dst + c); acc += v.x + v.y; }
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (acc == 1.2345) sink[0] = acc;
}

// A3: as A2 with one counter (red.release / ld.acquire) instead of per-producer flags.
__global__ void __launch_bounds__(512, 1) allgather_counter_kernel(double* vec, unsigned* counter, int per, int iters, long long* out, double* sink) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = nb * per;
    double acc = 0.0;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* dst = vec + (size_t)(it & 1) * n;
        if (warp == 0) {
            if (lane < per) st_d(dst + b * per + lane, (double)(it + lane));
            __syncwarp();
            if (lane == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
                unsigned seen;
                const long long tw_ = clock64();
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); SPIN_GUARD(tw_) } while (seen < (unsigned)it * nb);
            }
        }
        __syncthreads();
        for (int c = 2 * tid; c < n; c += 2 * blockDim.x) { const double2 v = ld_d2(dst + c); acc += v.x + v.y; }
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (acc == 1.2345) sink[0] = acc;
}

// B4: all-reduce of S values in two flag hand-overs: partials [nb][S] plain + flag per CTA; slot owners (warp per slot) add the
// nb partials of their slots and write sums[S] + flag per owner CTA; warp 8 polls the owner flags, everyone reads the S sums.
__global__ void __launch_bounds__(512, 1) allreduce_flag_kernel(double* part, double* sums, unsigned* flags1, unsigned* flags2, int S, int iters,
                                                                long long* out, double* sink) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_h[512];
    double acc = 0.0;
    long long t0 = 0;
    const int n_owner = S < nb ? S : nb;                              // CTAs that own at least one slot
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* pp = part + (size_t)(it & 1) * nb * S;
        double* ss = sums + (size_t)(it & 1) * S;
        for (int s = tid; s < S; s += blockDim.x) st_d(pp + (size_t)b * S + s, (double)(s + it));
        __syncthreads();
        if (tid == 0) { __threadfence(); st_flag(flags1 + b, (unsigned)it); }
        // owners: wait for all partial flags (each owner warp polls them itself), add, publish
        bool owner_any = false;
        for (int s = b + warp * nb; s < S && warp < 8; s += 8 * nb) {
            owner_any = true;
            const long long tw_ = clock64();
            for (int c = lane; c < nb; c += 32) {
                unsigned v;
                do { v = ld_flag(flags1 + c); SPIN_GUARD(tw_) } while ((int)(v - (unsigned)it) < 0);
            }
            __threadfence();
            __syncwarp();
            double x[5] = {0, 0, 0, 0, 0};
            for (int q = 0; q < 5; ++q) { const int c = lane + 32 * q; if (c < nb) { asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(x[q]) : "l"(pp + (size_t)c * S + s) : "memory"); } }
            double tot = (x[0] + x[1]) + (x[2] + x[3]) + x[4];
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) st_d(ss + s, tot);
        }
        // (a CTA owns at most a few slots, all handled by its warps 0..7; the last of them publishes)
        if (warp < 8) {
            __syncwarp();
        }
        // named barrier among owner warps is overkill here: warp 0 owns the first slot, later warps fewer; publish after a block barrier
        __syncthreads();
        if (tid == 0 && b < n_owner) { __threadfence(); st_flag(flags2 + b, (unsigned)it); }
        if (warp == 8) {
            const long long tw_ = clock64();
            for (int c = lane; c < n_owner; c += 32) {
                unsigned v;
                do { v = ld_flag(flags2 + c); SPIN_GUARD(tw_) } while ((int)(v - (unsigned)it) < 0);
            }
            __threadfence();
        }
        __syncthreads();
        for (int s = tid; s < S; s += blockDim.x) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(ss + s) : "memory"); s_h[s] = v; }
        __syncthreads();
        acc += s_h[tid % S];
        __syncthreads();
        (void)owner_any;
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (acc == 1.2345) sink[0] = acc;
}

// B6: all-reduce with floating-point atomics (order of the additions not fixed) + counter.
__global__ void __launch_bounds__(512, 1) allreduce_atomic_kernel(double* sums, unsigned* counter, int S, int iters, long long* out, double* sink) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    __shared__ double s_h[512];
    double acc = 0.0;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* ss = sums + (size_t)(it % 3) * S;                       // three buffers: the one of round it+1 is cleared in round it
        double* nx = sums + (size_t)((it + 1) % 3) * S;
        for (int s = tid; s < S; s += blockDim.x) {
            asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(ss + s), "d"((double)(s + it)) : "memory");
            if (b == 0) st_d(nx + s, 0.0);
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned seen;
            const long long tw_ = clock64();
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); SPIN_GUARD(tw_) } while (seen < (unsigned)it * nb);
        }
        __syncthreads();
        for (int s = tid; s < S; s += blockDim.x) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(ss + s) : "memory"); s_h[s] = v; }
        __syncthreads();
        acc += s_h[tid % S];
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (acc == 1.2345) sink[0] = acc;
}


// ---- optimised counter hand-overs: fence + relaxed atomic on the producer side, relaxed polling + ONE fence on the consumer side,
// bulk data with ordinary (coalesced) loads after the fence (a gpu-scope fence invalidates the SM's L1).  Results are verified.
__device__ __forceinline__ void arrive_counter(unsigned* counter) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ bool wait_counter(const unsigned* counter, unsigned target) {
    const long long tw_ = clock64();
    for (;;) {
        unsigned seen;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if ((int)(seen - target) >= 0) break;
        if (*((volatile unsigned*)&g_timeout) != 0u) return false;
        if (clock64() - tw_ > 400000000LL) { g_timeout = 1; return false; }
    }
    __threadfence();
    return true;
}

// A5: all-gather: per doubles per CTA; data by ordinary stores / loads.  load_kind 0: ld.global (default caching), 1: ld.relaxed.gpu
__global__ void __launch_bounds__(512, 1) allgather_opt_kernel(double* vec, unsigned* counter, int per, int iters, int load_kind, long long* out, unsigned* errs) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    const int n = nb * per;
    __shared__ double s_vec[2304];
    long long t0 = 0;
    unsigned bad = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* dst = vec + (size_t)(it & 1) * 4096;
        if (tid < per) dst[b * per + tid] = (double)(it * 7 + b * per + tid);
        __syncthreads();
        if (tid == 0) { arrive_counter(counter); if (!wait_counter(counter, (unsigned)it * nb)) bad |= 2; }
        __syncthreads();
        if (load_kind == 0) { for (int c = 2 * tid; c < n; c += 2 * blockDim.x) { const double2 v = *reinterpret_cast<const double2*>(dst + c); s_vec[c] = v.x; s_vec[c + 1] = v.y; } }
        else { for (int c = 2 * tid; c < n; c += 2 * blockDim.x) { const double2 v = ld_d2(dst + c); s_vec[c] = v.x; s_vec[c + 1] = v.y; } }
        __syncthreads();
        for (int c = tid; c < n; c += blockDim.x) if (s_vec[c] != (double)(it * 7 + c)) bad |= 1;
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (bad) atomicOr(errs, bad);
}

// B8: all-reduce with floating-point atomics, optimised counter (three rotating buffers; the buffer of round it+1 is cleared in round it-... by CTA 0 before its arrive)
__global__ void __launch_bounds__(512, 1) allreduce_atomic_opt_kernel(double* sums, unsigned* counter, int S, int iters, long long* out, unsigned* errs) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    __shared__ double s_h[512];
    long long t0 = 0;
    unsigned bad = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* ss = sums + (size_t)(it % 3) * 512;
        double* nx = sums + (size_t)((it + 1) % 3) * 512;
        for (int s = tid; s < S; s += blockDim.x) {
            asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(ss + s), "d"((double)(s + it)) : "memory");
            if (b == 0) nx[s] = 0.0;
        }
        __syncthreads();
        if (tid == 0) { arrive_counter(counter); if (!wait_counter(counter, (unsigned)it * nb)) bad |= 2; }
        __syncthreads();
        for (int s = tid; s < S; s += blockDim.x) s_h[s] = ss[s];
        __syncthreads();
        for (int s = tid; s < S; s += blockDim.x) if (s_h[s] != (double)(s + it) * nb) bad |= 1;
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (bad) atomicOr(errs, bad);
}

// B9: deterministic all-reduce in two counter hand-overs: partials [nb][S] -> owners (slot s: CTA s % nb, warp s / nb) add the nb
// partials of a slot in a fixed order -> sums[S] -> everyone.
__global__ void __launch_bounds__(512, 1) allreduce_det_opt_kernel(double* part, double* sums, unsigned* counter1, unsigned* counter2, int S, int iters,
                                                                   long long* out, unsigned* errs) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_h[512];
    long long t0 = 0;
    unsigned bad = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        double* pp = part + (size_t)(it & 1) * 160 * 512;
        double* ss = sums + (size_t)(it & 1) * 512;
        for (int s = tid; s < S; s += blockDim.x) pp[(size_t)b * S + s] = (double)(s + it);
        __syncthreads();
        if (tid == 0) { arrive_counter(counter1); if (!wait_counter(counter1, (unsigned)it * nb)) bad |= 2; }
        __syncthreads();
        for (int s = b + warp * nb; s < S; s += 16 * nb) {
            double x[5] = {0, 0, 0, 0, 0};
            for (int q = 0; q < 5; ++q) { const int c = lane + 32 * q; if (c < nb) x[q] = pp[(size_t)c * S + s]; }
            double tot = (x[0] + x[1]) + (x[2] + x[3]) + x[4];
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) ss[s] = tot;
        }
        __syncthreads();
        if (tid == 0) { arrive_counter(counter2); if (!wait_counter(counter2, (unsigned)it * nb)) bad |= 2; }
        __syncthreads();
        for (int s = tid; s < S; s += blockDim.x) s_h[s] = ss[s];
        __syncthreads();
        for (int s = tid; s < S; s += blockDim.x) if (s_h[s] != (double)(s + it) * nb) bad |= 1;
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (bad) atomicOr(errs, bad);
}

__global__ void fence_cost_kernel(double* scratch, long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long t0 = clock64();
    for (int i = 0; i < 200; ++i) __threadfence();
    out[0] = (clock64() - t0) / 200;
    t0 = clock64();
    for (int i = 0; i < 200; ++i) { scratch[i & 7] = (double)i; __threadfence(); }
    out[1] = (clock64() - t0) / 200;
}


// ---- Design D skeleton: the three hand-overs of one Lanczos step with dummy arithmetic ------------------------------------
//   hop 1  workers: S partial packets pbuf[s][b]; owner of slot s (worker s % nW, warp (s / nW) % 8) adds the nW partials
//   hop 2  owner writes the sum to R replicas hrep[r][s]; worker b reads replica b % R (warps 8..11)
//   hop 3  workers: `per` packets of the next vector qbuf[row]; the nA aggregator CTAs (blockIdx >= nW) each collect a quarter of the
//          vector, write it as plain doubles, fence, and set their word of the ready packet in R replicas; worker thread 0 polls
//          its replica, fences, and all threads read the plain vector with ordinary loads.
template <int R>
__global__ void __launch_bounds__(512, 1) skeleton_kernel(uint4* pbuf, uint4* hrep, uint4* qbuf, double* qplain, unsigned* ready, int nW, int nA,
                                                          int S, int per, int iters, long long* out, unsigned* errs) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = nW * per;
    __shared__ double s_h[512];
    __shared__ double s_vec[2304];
    long long t0 = 0, t_h1 = 0, t_h3 = 0, tq = 0;
    unsigned bad = 0;
    if (b >= nW) {                                               // aggregator a
        const int a = b - nW;
        if (a >= nA) return;
        const int c0 = (int)((long long)n * a / nA), c1 = (int)((long long)n * (a + 1) / nA);
        for (int it = 1; it <= iters; ++it) {
            double* dst = qplain + (size_t)(it & 1) * 4096;
            for (int c = c0 + tid; c < c1; c += blockDim.x) {
                uint4 v;
                { const long long tw_ = clock64(); do { v = ld_pk<0>(qbuf + c); SPIN_GUARD(tw_) } while ((int)(v.y - (unsigned)it) < 0 || (int)(v.w - (unsigned)it) < 0); }
                dst[c] = __hiloint2double((int)v.z, (int)v.x);
            }
            __syncthreads();
            if (tid == 0) __threadfence();
            __syncthreads();
            if (tid < R) st_flag(ready + (size_t)tid * 32 + a, (unsigned)it);     // replica r: 128-byte line r, word a
        }
        return;
    }
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        tq = clock64();
        // hop 1
        for (int s = tid; s < S; s += blockDim.x) st_pk<0>(pbuf + (size_t)s * nW + b, (unsigned)(s + it), 0u, (unsigned)it);
        for (int s = b + warp * nW; s < S && warp < 8; s += 8 * nW) {
            unsigned x[5] = {0, 0, 0, 0, 0};
            unsigned pend = 0;
            for (int q = 0; q < 5; ++q) if (lane + 32 * q < nW) pend |= 1u << q;
            const long long tw_ = clock64();
            while (pend) {
                const unsigned cur = pend;
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    if ((cur >> q) & 1u) { const uint4 v = ld_pk<0>(pbuf + (size_t)s * nW + lane + 32 * q); if (v.y == (unsigned)it && v.w == (unsigned)it) { x[q] = v.x; pend &= ~(1u << q); } }
                SPIN_GUARD(tw_)
            }
            unsigned tot = x[0] + x[1] + x[2] + x[3] + x[4];
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane < R) st_pk<0>(hrep + (size_t)lane * 512 + s, tot, 0u, (unsigned)it);
        }
        // hop 2
        if (warp >= 8 && warp < 12) {
            const int t = tid - 256;
            unsigned pend = 0;
            for (int q = 0; q < 4; ++q) if (t + 128 * q < S) pend |= 1u << q;
            const long long tw_ = clock64();
            while (pend) {
                const unsigned cur = pend;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((cur >> q) & 1u) { const uint4 v = ld_pk<0>(hrep + (size_t)(b % R) * 512 + t + 128 * q); if (v.y == (unsigned)it && v.w == (unsigned)it) { s_h[t + 128 * q] = (double)v.x; pend &= ~(1u << q); } }
                SPIN_GUARD(tw_)
            }
        }
        __syncthreads();
        if (tid == 0) t_h1 += clock64() - tq;
        for (int s = tid; s < S; s += blockDim.x) if (s_h[s] != (double)((unsigned)(s + it) * (unsigned)nW)) bad |= 1;
        tq = clock64();
        // hop 3
        if (tid < per) { const double v = (double)(it * 7 + b * per + tid); st_pk<0>(qbuf + b * per + tid, (unsigned)__double2loint(v), (unsigned)__double2hiint(v), (unsigned)it); }
        if (tid == 0) {
            const long long tw_ = clock64();
            for (;;) {
                const uint4 v = ld_pk<0>(reinterpret_cast<const uint4*>(ready + (size_t)(b % R) * 32));
                bool all = (int)(v.x - (unsigned)it) >= 0;
                if (nA > 1) all = all && (int)(v.y - (unsigned)it) >= 0;
                if (nA > 2) all = all && (int)(v.z - (unsigned)it) >= 0;
                if (nA > 3) all = all && (int)(v.w - (unsigned)it) >= 0;
                if (all) break;
                SPIN_GUARD(tw_)
            }
            __threadfence();
        }
        __syncthreads();
        const double* src = qplain + (size_t)(it & 1) * 4096;
        for (int c = 2 * tid; c < n; c += 2 * blockDim.x) { const double2 v = *reinterpret_cast<const double2*>(src + c); s_vec[c] = v.x; s_vec[c + 1] = v.y; }
        __syncthreads();
        if (tid == 0) t_h3 += clock64() - tq;
        for (int c = tid; c < n; c += blockDim.x) if (s_vec[c] != (double)(it * 7 + c)) bad |= 2;
        __syncthreads();
    }
    if (tid == 0) { out[b] = clock64() - t0; out[256 + b] = t_h1; out[512 + b] = t_h3; }
    if (bad) atomicOr(errs, bad);
}

__global__ void __launch_bounds__(512, 1) barrier_kernel(unsigned* bar, int iters, long long* out) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned seen;
            { const long long tw_ = clock64(); do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); SPIN_GUARD(tw_) } while (seen < (unsigned)it * nb); }
        }
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
}


// two-level grid barrier: groups of GS CTAs arrive on their own counter (one 128-byte line each); the last arriver of a group
// arrives on the top counter; everybody polls the top counter.  poll_kind 0: ld.acquire polling, 1: ld.relaxed polling + one fence
__global__ void __launch_bounds__(512, 1) barrier2_kernel(unsigned* bar, int GS, int poll_kind, int iters, long long* out) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    const int g = b / GS, ng = (nb + GS - 1) / GS, gsize = min(GS, nb - g * GS);
    unsigned* gcnt = bar + 64 + 32 * g;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        __syncthreads();
        if (tid == 0) {
            unsigned old;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(gcnt) : "memory");
            if (old + 1 == (unsigned)it * gsize) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned seen;
            const long long tw_ = clock64();
            if (poll_kind == 0) {
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); SPIN_GUARD(tw_) } while (seen < (unsigned)it * ng);
            } else {
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); SPIN_GUARD(tw_) } while (seen < (unsigned)it * ng);
                __threadfence();
            }
        }
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
}
// flat barrier with relaxed polling + fence
__global__ void __launch_bounds__(512, 1) barrier_relaxed_kernel(unsigned* bar, int iters, long long* out) {
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 2) t0 = clock64();
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned seen;
            const long long tw_ = clock64();
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); SPIN_GUARD(tw_) } while (seen < (unsigned)it * nb);
            __threadfence();
        }
        __syncthreads();
    }
    if (tid == 0) out[b] = clock64() - t0;
}

template <class K, class... A> static void coop(K kern, int grid, int threads, A... args) {
    void* argv[] = {(void*)&args...};
    unsigned z = 0; CK(cudaMemcpyToSymbol(g_timeout, &z, 4));
    CK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(threads), argv, 0, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(&z, g_timeout, 4));
    if (z) printf("  [spin wait TIMED OUT in the next line's test]\n");
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, nsm);
    uint4* buf; long long* out; double* sink; unsigned* bar;
    CK(cudaMalloc(&buf, 64 << 20)); CK(cudaMalloc(&out, 8 * 1024)); CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&bar, 8192));
    std::vector<long long> h(1024);
    const int iters = 1000;
    const char* names[4] = {"st.volatile/ld.volatile", "st.relaxed.gpu/ld.relaxed.gpu", "st.release.gpu/ld.acquire.gpu", "st.cg/ld.cg"};
    for (int peer : {1, 74, 147}) {
        if (peer >= nsm) continue;
        for (int f = 0; f < 3; ++f) {
            CK(cudaMemset(buf, 0, 4096));
            if (f == 0) coop(pingpong_kernel<0>, nsm, 32, buf, peer, iters, out);
            if (f == 1) coop(pingpong_kernel<1>, nsm, 32, buf, peer, iters, out);
            if (f == 2) coop(pingpong_kernel<2>, nsm, 32, buf, peer, iters, out);
            if (f == 3) coop(pingpong_kernel<3>, nsm, 32, buf, peer, iters, out);
            CK(cudaMemcpy(h.data(), out, 8, cudaMemcpyDeviceToHost));
            printf("pingpong CTA0<->CTA%-3d %-32s one-way %6.0f cycles\n", peer, names[f], (double)h[0] / (2.0 * iters));
        }
    }
    nanosleep_kernel<<<1, 32>>>(out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), out, 48, cudaMemcpyDeviceToHost));
    printf("nanosleep(0,20,40,100,200,1000) -> %lld %lld %lld %lld %lld %lld cycles\n", h[0], h[1], h[2], h[3], h[4], h[5]);
    for (int per : {14}) for (int mode = 0; mode < 2; ++mode) for (int f : {0, 1}) {
        CK(cudaMemset(buf, 0, 1 << 20));
        const int grid = nsm - 5;
        if (f == 0) coop(allgather_kernel<0>, grid, 512, buf, per, 500, mode, out, sink);
        if (f == 1) coop(allgather_kernel<1>, grid, 512, buf, per, 500, mode, out, sink);
        if (f == 3) coop(allgather_kernel<3>, grid, 512, buf, per, 500, mode, out, sink);
        CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("allgather %d CTAs x %d packets, poll mode %d, %-30s %7.0f cycles per round\n", grid, per, mode, names[f], (double)mx / 499.0);
    }
    for (int S : {64, 200, 400}) for (int f : {0, 1}) {
        CK(cudaMemset(buf, 0, 8 << 20));
        const int grid = nsm - 5;
        uint4* hb = buf + (4 << 20) / 16;
        if (f == 0) coop(reduce_bcast_kernel<0>, grid, 512, buf, hb, S, 500, out, sink);
        if (f == 1) coop(reduce_bcast_kernel<1>, grid, 512, buf, hb, S, 500, out, sink);
        if (f == 3) coop(reduce_bcast_kernel<3>, grid, 512, buf, hb, S, 500, out, sink);
        CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("reduce+bcast %d CTAs, %3d slots, %-30s %7.0f cycles per round\n", grid, S, names[f], (double)mx / 499.0);
    }
    {
        CK(cudaMemset(bar, 0, 64));
        coop(barrier_kernel, nsm, 512, bar, 1000, out);
        CK(cudaMemcpy(h.data(), out, 8 * nsm, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid barrier (red.release/ld.acquire), %d CTAs: %7.0f cycles per barrier\n", nsm, (double)mx / 999.0);
    }

    {
        const int grid = nsm - 5;
        double* dv = reinterpret_cast<double*>(buf);
        unsigned* fl = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(buf) + (16 << 20));
        for (int variant = 0; variant < 2; ++variant) {
            CK(cudaMemset(buf, 0, 32 << 20));
            coop(allgather_flag_kernel, grid, 512, dv, fl, 14, 500, variant, out, sink);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("allgather plain data + flag per producer (%s), %d CTAs x 14 doubles: %7.0f cycles per round\n", variant ? "flags 128 B apart" : "flags packed", grid, (double)mx / 499.0);
        }
        CK(cudaMemset(buf, 0, 32 << 20));
        coop(allgather_counter_kernel, grid, 512, dv, fl, 14, 500, out, sink);
        CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
        { long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("allgather plain data + one counter, %d CTAs x 14 doubles: %7.0f cycles per round\n", grid, (double)mx / 499.0); }
        for (int S : {64, 200, 400}) {
            CK(cudaMemset(buf, 0, 32 << 20));
            double* part = dv; double* sums = dv + (1 << 20);
            coop(allreduce_flag_kernel, grid, 512, part, sums, fl, fl + 4096, S, 500, out, sink);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("allreduce plain + flags (2 hand-overs), %d CTAs, %3d values: %7.0f cycles per round\n", grid, S, (double)mx / 499.0);
            CK(cudaMemset(buf, 0, 32 << 20));
            coop(allreduce_atomic_kernel, grid, 512, sums, fl, S, 500, out, sink);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
            mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("allreduce red.add.f64 + counter, %d CTAs, %3d values: %7.0f cycles per round\n", grid, S, (double)mx / 499.0);
        }
    }

    {
        const int grid = nsm - 5;
        double* dv = reinterpret_cast<double*>(buf);
        unsigned* fl = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(buf) + (16 << 20));
        unsigned* errs = fl + 1024;
        unsigned he = 0;
        fence_cost_kernel<<<1, 32>>>(dv, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, 16, cudaMemcpyDeviceToHost));
        printf("__threadfence alone %lld cycles, after a store %lld cycles\n", h[0], h[1]);
        for (int lk = 0; lk < 2; ++lk) {
            CK(cudaMemset(buf, 0, 32 << 20));
            coop(allgather_opt_kernel, grid, 512, dv, fl, 14, 500, lk, out, errs);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&he, errs, 4, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("allgather OPT (fence+relaxed counter, %s loads), %d CTAs x 14 doubles: %7.0f cycles per round, errors %u\n", lk ? "ld.relaxed.gpu" : "ordinary", grid, (double)mx / 499.0, he);
        }
        for (int S : {64, 200, 400}) {
            CK(cudaMemset(buf, 0, 32 << 20));
            coop(allreduce_atomic_opt_kernel, grid, 512, dv, fl, S, 500, out, errs);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&he, errs, 4, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("allreduce OPT red.add.f64 + counter, %d CTAs, %3d values: %7.0f cycles per round, errors %u\n", grid, S, (double)mx / 499.0, he);
            CK(cudaMemset(buf, 0, 32 << 20));
            coop(allreduce_det_opt_kernel, grid, 512, dv, dv + (2 << 20) / 8 * 4, fl, fl + 64, S, 500, out, errs);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&he, errs, 4, cudaMemcpyDeviceToHost));
            mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("allreduce OPT deterministic two counters, %d CTAs, %3d values: %7.0f cycles per round, errors %u\n", grid, S, (double)mx / 499.0, he);
        }
    }

    {
        const int nW = nsm - 5;
        char* base = reinterpret_cast<char*>(buf);
        uint4* pbuf = reinterpret_cast<uint4*>(base);                       // 512 slots x 160 workers x 16 B = 1.3 MB
        uint4* hrep = reinterpret_cast<uint4*>(base + (2 << 20));           // 8 replicas x 512
        uint4* qb = reinterpret_cast<uint4*>(base + (3 << 20));
        double* qplain = reinterpret_cast<double*>(base + (4 << 20));
        unsigned* ready = reinterpret_cast<unsigned*>(base + (5 << 20));
        unsigned* errs = reinterpret_cast<unsigned*>(base + (6 << 20));
        unsigned he = 0;
        for (int nA : {1, 2, 4}) for (int S : {64, 200, 400}) {
            CK(cudaMemset(buf, 0, 8 << 20));
            coop(skeleton_kernel<8>, nW + nA, 512, pbuf, hrep, qb, qplain, ready, nW, nA, S, 14, 500, out, errs);
            CK(cudaMemcpy(h.data(), out, 8 * 768, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&he, errs, 4, cudaMemcpyDeviceToHost));
            long long mx = 0, m1 = 0, m3 = 0; for (int i = 0; i < nW; ++i) { mx = h[i] > mx ? h[i] : mx; m1 = h[256 + i] > m1 ? h[256 + i] : m1; m3 = h[512 + i] > m3 ? h[512 + i] : m3; }
            printf("skeleton D: %d workers, %d aggregators, %3d slots, 8 replicas: %7.0f cycles per step (reduce+bcast %6.0f, allgather %6.0f), errors %u\n",
                   nW, nA, S, (double)mx / 499.0, (double)m1 / 500.0, (double)m3 / 500.0, he);
        }
    }

    {
        for (int GS : {4, 8, 12, 16, 24, 37}) for (int pk = 0; pk < 2; ++pk) {
            CK(cudaMemset(bar, 0, 8192));
            coop(barrier2_kernel, nsm, 512, bar, GS, pk, 1000, out);
            CK(cudaMemcpy(h.data(), out, 8 * nsm, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("two-level barrier, groups of %2d, %s polling, %d CTAs: %7.0f cycles per barrier\n", GS, pk ? "relaxed+fence" : "acquire", nsm, (double)mx / 999.0);
        }
        CK(cudaMemset(bar, 0, 8192));
        coop(barrier_relaxed_kernel, nsm, 512, bar, 1000, out);
        CK(cudaMemcpy(h.data(), out, 8 * nsm, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("flat barrier, relaxed polling + fence, %d CTAs: %7.0f cycles per barrier\n", nsm, (double)mx / 999.0);
    }
    return 0;
}
