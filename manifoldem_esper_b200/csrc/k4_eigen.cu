// k4_eigen.cu -- K4: leading eigenpairs of the symmetric Markov matrix Ms (m x m, FP64), replacing the
// full SVD of 1_Embedding/DM/DiffusionMaps.py:166-169 (Ms is symmetric PSD, so singular values =
// eigenvalues and U = eigenvectors; the reference saves s**2 and U).
//
// Method: Lanczos with full reorthogonalisation on the operator deflated by the analytically known
// top pair (lambda_0 = 1, v_0 = sqrt(d)/|sqrt(d)|), which is kept as a locked vector in the basis.
// The iteration runs as ONE persistent cooperative kernel per batch of steps (one CTA per SM, grid
// barriers between the phases of a step), so a step costs three grid barriers instead of six kernel
// launches:
//     S0  w = Ms q_j - alpha_{j-1} q_j - beta_{j-1} q_{j-1}   each CTA owns a row block; q_j staged in shared memory; SYMV reads Ms once
//     S1  h = V^T w         one warp per basis vector (classical Gram-Schmidt coefficients)
//     S2  w -= V h          CTA-owned rows; partial |w|^2; a second S1/S2 pass only when the DGKS
//                           test (|w_after|^2 < |w_before|^2 / 2, "twice is enough") asks for it; S0
//                           removes the predicted three-term components so that it rarely does
// Every pass is a streaming FP64 pass (Ms: m^2 x 8 B per step, L2-resident at m = 2000); all
// reductions have a fixed order (deterministic).  The host solves the small tridiagonal problem
// (bisection + inverse iteration) to test convergence while the next batch already runs.
#include "common.cuh"
#include "devutil.cuh"
#include <algorithm>
namespace esper {
namespace lanczos {

// h[i] = V_i . w for i < nb (one block per basis vector).
__global__ void __launch_bounds__(256) dots_kernel(const double* __restrict__ V, int m, const double* __restrict__ w,
                                                   double* __restrict__ h) {
    __shared__ double sh[32];
    const double* v = V + (size_t)blockIdx.x * m;
    double s = 0.0;
    for (int r = threadIdx.x; r < m; r += blockDim.x) s = fma(__ldg(v + r), __ldg(w + r), s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) h[blockIdx.x] = s;
}

// w -= sum_i h[i] V_i ; the component along the current Lanczos vector accumulates into alpha.
__global__ void __launch_bounds__(256) update_kernel(const double* __restrict__ V, int m, int nb,
                                                     const double* __restrict__ h, double* __restrict__ w,
                                                     double* __restrict__ alpha_slot) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) {
        double acc = w[r];
        for (int i = 0; i < nb; ++i) acc = fma(-__ldg(h + i), __ldg(V + (size_t)i * m + r), acc);
        w[r] = acc;
    }
    if (alpha_slot && blockIdx.x == 0 && threadIdx.x == 0) *alpha_slot += h[nb - 1];
}

// Row-block (replicated) orthogonalisation at large m: the same two operations with enough loads in flight to stream V
// from HBM.  dots: grid (basis vector, DOT_SPLIT row ranges), four independent accumulators per thread, partial sums
// h_part[i][s]; update: every CTA first adds the DOT_SPLIT partials of every coefficient in a fixed order (shared
// memory; every rank computes the same bits), then w -= sum_i h[i] V_i with four independent accumulators.
constexpr int DOT_SPLIT = 8;
__global__ void __launch_bounds__(256) dots_split_kernel(const double* __restrict__ V, int m, const double* __restrict__ w,
                                                         double* __restrict__ h_part) {
    __shared__ double sh[32];
    const double* v = V + (size_t)blockIdx.x * m;
    const int chunk = (m + DOT_SPLIT - 1) / DOT_SPLIT;
    const int r0 = blockIdx.y * chunk, r1 = min(m, r0 + chunk);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = r0 + threadIdx.x;
    for (; r + 768 < r1; r += 1024) {
        const double x0 = __ldg(v + r), x1 = __ldg(v + r + 256), x2 = __ldg(v + r + 512), x3 = __ldg(v + r + 768);
        s0 = fma(x0, __ldg(w + r), s0);       s1 = fma(x1, __ldg(w + r + 256), s1);
        s2 = fma(x2, __ldg(w + r + 512), s2); s3 = fma(x3, __ldg(w + r + 768), s3);
    }
    for (; r < r1; r += 256) s0 = fma(__ldg(v + r), __ldg(w + r), s0);
    const double s = block_sum((s0 + s1) + (s2 + s3), sh);
    if (threadIdx.x == 0) h_part[(size_t)blockIdx.x * DOT_SPLIT + blockIdx.y] = s;
}

__global__ void __launch_bounds__(256) update_split_kernel(const double* __restrict__ V, int m, int nb,
                                                           const double* __restrict__ h_part, double* __restrict__ w,
                                                           double* __restrict__ alpha_slot) {
    extern __shared__ double hs[];                     // [nb] summed coefficients
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        double t = h_part[(size_t)i * DOT_SPLIT];
#pragma unroll
        for (int q = 1; q < DOT_SPLIT; ++q) t += h_part[(size_t)i * DOT_SPLIT + q];
        hs[i] = t;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) {
        double a0 = w[r], a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = 0;
        for (; i + 3 < nb; i += 4) {
            const double x0 = __ldg(V + (size_t)i * m + r), x1 = __ldg(V + (size_t)(i + 1) * m + r),
                         x2 = __ldg(V + (size_t)(i + 2) * m + r), x3 = __ldg(V + (size_t)(i + 3) * m + r);
            a0 = fma(-hs[i], x0, a0);     a1 = fma(-hs[i + 1], x1, a1);
            a2 = fma(-hs[i + 2], x2, a2); a3 = fma(-hs[i + 3], x3, a3);
        }
        for (; i < nb; ++i) a0 = fma(-hs[i], __ldg(V + (size_t)i * m + r), a0);
        w[r] = (a0 + a1) + (a2 + a3);
    }
    if (alpha_slot && blockIdx.x == 0 && threadIdx.x == 0) *alpha_slot += hs[nb - 1];
}

__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }

struct StepParams {
    const double* Ms; int m;
    double* V;            // basis: V[0] = v0 (locked), V[j] = q_j
    double* wbuf;         // [2][m] ping-pong work vectors
    double* h;            // [>= max basis + 2] Gram-Schmidt coefficients (+ slot for w.w)
    double* part;         // [gridDim.x] partial |w|^2
    double* alpha;        // alpha[j-1] for q_j
    double* beta;         // beta[j-1] = |w| after step j
    int j_start;          // first Lanczos vector index of this batch (>= 1); V[j_start] is normalised
    int n_steps;
    unsigned int* bar;    // bar[0] arrival counter (zeroed before every launch), bar[1] abort flag
    int rows_cached;      // resident kernel: owned rows of Ms kept in shared memory (the rest, <= 2, in registers)
    long long* prof;      // optional [8] cycle counters of CTA 0 per phase (debug), or NULL
};

// Grid-wide barrier for a co-resident (cooperatively launched) grid: monotone arrival counter, the
// k-th barrier completes when the counter reaches k * gridDim.x.  A barrier that does not complete
// within ~2 s sets the abort flag instead of spinning forever; every CTA then leaves the kernel.
struct GridBarrier {
    unsigned int* bar;
    unsigned int epoch = 0;
    bool aborted = false;
    __device__ GridBarrier(unsigned int* b) : bar(b) {}
    __device__ __forceinline__ bool sync() {
        __shared__ int s_abort;
        __syncthreads();
        if (threadIdx.x == 0) {
            ++epoch;
            int ab = 0;
            // release-arrive / relaxed poll + ONE acquire fence after the last poll (an acquire on every polling load costs a
            // fence per poll: measured 2790 -> 2440 cycles per barrier, profiles/r2_handover_microbench.md); bar.sync above
            // orders the CTA's writes before the release
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            const unsigned int target = epoch * gridDim.x;
            const long long t0 = clock64();
            for (unsigned int polls = 1;; ++polls) {
                unsigned int seen;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
                if (seen >= target) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); break; }
                if ((polls & 255u) == 0u) {                      // watchdog: another CTA gave up, or ~2 s without progress
                    if (*((volatile unsigned int*)(bar + 1)) != 0u) { ab = 1; break; }
                    if (clock64() - t0 > 4000000000LL) { atomicExch(bar + 1, 1u); ab = 1; break; }
                }
            }
            s_abort = ab;
        }
        __syncthreads();
        if (threadIdx.x != 0) ++epoch;
        aborted = aborted || (s_abort != 0);
        return !aborted;
    }
};

constexpr int LAN_THREADS = 512;
constexpr int LAN_WARPS = LAN_THREADS / 32;
constexpr int LAN_CH = 2048;          // columns of the vector staged in shared memory at a time
constexpr int LAN_H = 1040;           // Gram-Schmidt coefficients staged in shared memory (max_iter <= 1024)

// Sum of the per-CTA partial norms, identical in every thread of every CTA (fixed reduction tree).
__device__ __forceinline__ double sum_partials(const double* part, int nb, double* s_tmp /*[LAN_WARPS + 1]*/) {
    double v = 0.0;
    for (int q = threadIdx.x; q < nb; q += LAN_THREADS) v += ld_cg(part + q);
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < LAN_WARPS; ++w) tot += s_tmp[w];
        s_tmp[LAN_WARPS] = tot;
    }
    __syncthreads();
    return s_tmp[LAN_WARPS];
}

// Persistent cooperative kernel: n_steps Lanczos steps with full reorthogonalisation.
__global__ void __launch_bounds__(LAN_THREADS, 2) lanczos_steps_kernel(StepParams p) {
    GridBarrier grid(p.bar);
    __shared__ double s_vec[LAN_CH];
    __shared__ double s_red[32 * 16];
    __shared__ double s_nrm[16];
    __shared__ double s_tmp[LAN_WARPS + 1];
    const int m = p.m;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = gridDim.x, b = blockIdx.x;
    const int chunk = (m + nb - 1) / nb;
    const int row0 = min(b * chunk, m), row1 = min(row0 + chunk, m);

    int cur = 0;                                     // wbuf[cur] holds the un-normalised q_j (after the first step)
    double s2_last = 0.0;                            // |w|^2 of the previous step (set by the DGKS test)
    for (int step = 0; step < p.n_steps; ++step) {
        const int j = p.j_start + step;
        // ---- source of q_j: normalised V[j] on the first step, else wbuf[cur] / beta ----
        const double* src; double scale = 1.0;
        // Predicted large components, removed locally before the Gram-Schmidt pass (three-term recurrence
        // with alpha_j predicted by alpha_{j-1}); the pass then measures only small corrections, which is
        // what makes a single classical pass stable (Kahan-Parlett / DGKS: |w_after| >= 0.7 |w_before|).
        double sigma = 0.0, beta_prev = 0.0;
        if (j >= 2) { sigma = ld_cg(p.alpha + j - 2); beta_prev = ld_cg(p.beta + j - 2); }
        if (step == 0) {
            src = p.V + (size_t)j * m;
        } else {
            const double bt = sqrt(s2_last);
            scale = bt > 0.0 ? 1.0 / bt : 0.0;
            src = p.wbuf + (size_t)cur * m;
            beta_prev = bt;
            if (b == 0 && tid == 0) p.beta[j - 2] = bt;              // beta of the previous step (vector j-1)
            for (int r = row0 + tid; r < row1; r += LAN_THREADS)       // owners publish the normalised q_j
                p.V[(size_t)j * m + r] = ld_cg(src + r) * scale;
        }
        double* wnew = p.wbuf + (size_t)(cur ^ 1) * m;

        // ---- S0: w = Ms q_j - sigma q_j - beta q_{j-1} for the owned rows (one warp per row) ----
        for (int rbase = row0; rbase < row1; rbase += LAN_WARPS) {
            const int r = rbase + warp;
            double acc = 0.0;
            for (int c0 = 0; c0 < m; c0 += LAN_CH) {
                const int cn = min(LAN_CH, m - c0);
                __syncthreads();
                for (int cidx = tid; cidx < cn; cidx += LAN_THREADS) s_vec[cidx] = ld_cg(src + c0 + cidx);
                __syncthreads();
                if (r < row1) {
                    const double* a = p.Ms + (size_t)r * m + c0;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    for (int cidx = lane; cidx < cn; cidx += 256) {    // 8 independent (predicated) loads in flight per lane
                        double x[8], y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int cc = cidx + 32 * u;
                            const bool ok = cc < cn;
                            x[u] = ok ? __ldg(a + cc) : 0.0;
                            y[u] = ok ? s_vec[cc] : 0.0;
                        }
                        a0 = fma(x[0], y[0], a0); a1 = fma(x[1], y[1], a1); a2 = fma(x[2], y[2], a2); a3 = fma(x[3], y[3], a3);
                        a0 = fma(x[4], y[4], a0); a1 = fma(x[5], y[5], a1); a2 = fma(x[6], y[6], a2); a3 = fma(x[7], y[7], a3);
                    }
                    acc += (a0 + a1) + (a2 + a3);
                }
            }
            const double tot = warp_sum(acc);
            if (r < row1 && lane == 0) {
                double x = fma(-sigma, ld_cg(src + r) * scale, tot * scale);
                if (j >= 2) x = fma(-beta_prev, ld_cg(p.V + (size_t)(j - 1) * m + r), x);
                wnew[r] = x;
            }
        }
        if (!grid.sync()) return;

        for (int pass = 0; pass < 2; ++pass) {
            // ---- S1: h[i] = V_i . w for i <= j, and (first pass only) h[j+1] = w . w ----
            // (the second pass must not touch h[j+1]: slower CTAs may still be reading it for the DGKS test)
            // Eight warps share one basis vector, two vectors per CTA per round.
            const int i_end = (pass == 0) ? j + 1 : j;
            const int grp = warp >> 3, gw = warp & 7;
            for (int i0 = 0; i0 <= i_end; i0 += 2 * nb) {
                const int i = i0 + grp * nb + b;
                double sacc = 0.0;
                if (i <= i_end) {
                    const double* v = (i <= j) ? p.V + (size_t)i * m : wnew;
                    double a0 = 0.0, a1 = 0.0;
                    for (int r = gw * 32 + lane; r < m; r += 2048) {  // 8 independent (predicated) load pairs per lane
                        double x[8], y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int rr = r + 256 * u;
                            const bool ok = rr < m;
                            x[u] = ok ? ld_cg(v + rr) : 0.0;
                            y[u] = ok ? ld_cg(wnew + rr) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u += 2) { a0 = fma(x[u], y[u], a0); a1 = fma(x[u + 1], y[u + 1], a1); }
                    }
                    sacc = warp_sum(a0 + a1);
                }
                __syncthreads();
                if (lane == 0) s_red[warp] = sacc;
                __syncthreads();
                if (gw == 0 && lane == 0 && i <= i_end) {
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) tot += s_red[grp * 8 + q];
                    p.h[i] = tot;
                }
            }
            if (!grid.sync()) return;

            // ---- S2: w -= V h on the owned rows; partial |w|^2 ----
            double nrm = 0.0;
            for (int rbase = row0; rbase < row1; rbase += 16) {
                const int r = rbase + (tid & 15), sub = tid >> 4;
                double sacc = 0.0;
                if (r < row1) {
                    double a0 = 0.0, a1 = 0.0;
                    for (int i = sub; i <= j; i += 256) {             // 8 independent (predicated) load pairs per thread
                        double x[8], y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int ii = i + 32 * u;
                            const bool ok = ii <= j;
                            x[u] = ok ? ld_cg(p.h + ii) : 0.0;
                            y[u] = ok ? ld_cg(p.V + (size_t)ii * m + r) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u += 2) { a0 = fma(x[u], y[u], a0); a1 = fma(x[u + 1], y[u + 1], a1); }
                    }
                    sacc = a0 + a1;
                }
                __syncthreads();
                s_red[sub * 16 + (tid & 15)] = sacc;
                __syncthreads();
                if (tid < 16 && r < row1) {
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < 32; ++q) tot += s_red[q * 16 + tid];
                    const double x = ld_cg(wnew + r) - tot;
                    wnew[r] = x;
                    nrm = fma(x, x, nrm);
                }
            }
            if (tid < 16) s_nrm[tid] = nrm;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) tot += s_nrm[q];
                p.part[b] = tot;
                if (b == 0) { if (pass == 0) p.alpha[j - 1] = sigma + ld_cg(p.h + j); else p.alpha[j - 1] += ld_cg(p.h + j); }
            }
            if (!grid.sync()) return;
            s2_last = sum_partials(p.part, nb, s_tmp);
            // DGKS ("twice is enough"): second pass iff |w_after|^2 < |w_before|^2 / 2
            if (pass == 0 && !(s2_last < 0.5 * ld_cg(p.h + j + 1))) break;
        }
        cur ^= 1;
    }
    // ---- tail: publish beta of the last step and the normalised next vector ----
    {
        const int j = p.j_start + p.n_steps;                           // index of the next vector
        const double bt = sqrt(s2_last);
        const double scale = bt > 0.0 ? 1.0 / bt : 0.0;
        const double* src = p.wbuf + (size_t)cur * m;
        if (b == 0 && tid == 0) p.beta[j - 2] = bt;
        for (int r = row0 + tid; r < row1; r += LAN_THREADS) p.V[(size_t)j * m + r] = ld_cg(src + r) * scale;
    }
}

// ------------------------------------------------------------------------------------------------
// Resident variant for m <= LAN_CH (the BASELINE shape, m = 2000): Ms never leaves the chip.
//
// Ms is the same at every step and a CTA always needs the same rows, so each CTA loads its rows ONCE per launch:
// up to ~13 rows into shared memory (208 kB) and at most two more into registers.  148 SMs x 13.5 rows x 16 kB
// is the whole 32 MB matrix.  Per step, the only global traffic is the Lanczos vector (16 kB per CTA), the
// basis rows for the Gram-Schmidt passes and a few scalars; every phase is arranged so that it waits for at
// most ONE L2 round trip (measured cycle budget per step in DESIGN.md).  Same arithmetic and the same three
// grid barriers as lanczos_steps_kernel.
// ------------------------------------------------------------------------------------------------
#define RES_PROF(slot)                                                                      \
    do {                                                                                    \
        if (p.prof && b == 0 && tid == 0) { const long long t_ = clock64(); p.prof[slot] += t_ - t_prof; t_prof = t_; } \
    } while (0)

// 16 per-lane values -> 16 warp totals with 16 shuffles (instead of 16 x 5): every round halves the number of values
// a lane carries by trading the other half with its partner.  Lane l ends with the total of value index
// ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1); fixed order, deterministic.
__device__ __forceinline__ double warp_multi_sum16(double (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 16;
        const double send = up ? v[i] : v[i + 8];
        const double keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 8;
        const double send = up ? v[i] : v[i + 4];
        const double keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 4;
        const double send = up ? v[i] : v[i + 2];
        const double keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool up = lane & 2;
        const double send = up ? v[0] : v[1];
        const double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// RES = true : rows of Ms resident in shared memory / registers, one CTA per SM (lowest latency for one PD).
// RES = false: rows streamed from L2 every step, <= 64 registers and 20 kB of shared memory, so that two such grids
//              (two PDs in flight) share every SM and hide each other's barrier latency (highest throughput).
template <bool RES>
__global__ void __launch_bounds__(LAN_THREADS, RES ? 1 : 2) lanczos_resident_kernel(StepParams p) {
    GridBarrier grid(p.bar);
    extern __shared__ double res_smem[];
    double* s_vec = res_smem;                        // [LAN_CH] q_j, normalised, all columns
    double* s_rows = res_smem + LAN_CH;              // [rows_cached][m]
    __shared__ double s_red[32 * 16];
    __shared__ double s_prev[16];                    // q_{j-1} on the owned rows
    __shared__ double s_h[LAN_H];                    // Gram-Schmidt coefficients of the step (staged once per CTA)
    __shared__ double s_nrm[16];
    __shared__ double s_tmp[LAN_WARPS + 1];
    const int m = p.m;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = gridDim.x, b = blockIdx.x;
    const int chunk = (m + nb - 1) / nb;
    const int row0 = min(b * chunk, m), row1 = min(row0 + chunk, m);
    const int n_own = row1 - row0;                   // <= 16
    const int n_sm = min(p.rows_cached, n_own);      // rows in shared memory; rows n_sm, n_sm+1 live in registers
    long long t_prof = clock64();

    // ---- once per launch: rows of Ms -> shared memory / registers; q_j, q_{j-1}, alpha_{j-1}, beta_{j-1} ----
    int j = p.j_start;
    double rrow[2][4];
    if (RES) {
        for (int rr = 0; rr < n_sm; ++rr) {
            const double* a = p.Ms + (size_t)(row0 + rr) * m;
            for (int c = tid; c < m; c += LAN_THREADS) s_rows[(size_t)rr * m + c] = __ldg(a + c);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = tid + LAN_THREADS * u;
                rrow[e][u] = (n_sm + e < n_own && c < m) ? __ldg(p.Ms + (size_t)(row0 + n_sm + e) * m + c) : 0.0;
            }
    }
    for (int c = tid; c < m; c += LAN_THREADS) s_vec[c] = ld_cg(p.V + (size_t)j * m + c);
    if (tid < 16) s_prev[tid] = (j >= 2 && tid < n_own) ? ld_cg(p.V + (size_t)(j - 1) * m + row0 + tid) : 0.0;
    double sigma = 0.0, beta_prev = 0.0;             // alpha_{j-1} (shift predictor), beta_{j-1}; carried in registers
    if (j >= 2) { sigma = ld_cg(p.alpha + j - 2); beta_prev = ld_cg(p.beta + j - 2); }
    __syncthreads();
    RES_PROF(7);

    int cur = 0;
    for (int step = 0; step < p.n_steps; ++step, ++j) {
        double* wnew = p.wbuf + (size_t)cur * m;

        // ---- S0: w = Ms q_j - sigma q_j - beta_{j-1} q_{j-1} on the owned rows ----
        if (RES) {                                   // on-chip operands only
            double acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = tid + LAN_THREADS * u;
                if (c < m) {
                    const double qv = s_vec[c];
#pragma unroll
                    for (int q = 0; q < 14; ++q)
                        if (q < n_sm) acc[q] = fma(s_rows[(size_t)q * m + c], qv, acc[q]);
                    acc[14] = fma(rrow[0][u], qv, acc[14]);
                    acc[15] = fma(rrow[1][u], qv, acc[15]);
                }
            }
            const double t = warp_multi_sum16(acc, lane);
            if ((lane & 1) == 0) s_red[warp * 16 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = t;
            __syncthreads();
            if (tid < n_own) {
                const int slot = tid < n_sm ? tid : 14 + (tid - n_sm);      // register rows sit in slots 14, 15
                double tot = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < LAN_WARPS; ++w2) tot += s_red[w2 * 16 + slot];
                const int r = row0 + tid;
                const double qr = s_vec[r];
                double x = fma(-sigma, qr, tot);
                x = fma(-beta_prev, s_prev[tid], x);
                wnew[r] = x;
                s_prev[tid] = qr;                                           // q_j on the owned rows, for the next step
            }
        } else {                                     // rows streamed from L2: one warp per row, 8 loads in flight per lane
            const int r = row0 + warp;
            if (warp < n_own) {
                const double* a = p.Ms + (size_t)r * m;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int cidx = lane; cidx < m; cidx += 256) {
                    double x[8], y[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int cc = cidx + 32 * u;
                        const bool ok = cc < m;
                        x[u] = ok ? __ldg(a + cc) : 0.0;
                        y[u] = ok ? s_vec[cc] : 0.0;
                    }
                    a0 = fma(x[0], y[0], a0); a1 = fma(x[1], y[1], a1); a2 = fma(x[2], y[2], a2); a3 = fma(x[3], y[3], a3);
                    a0 = fma(x[4], y[4], a0); a1 = fma(x[5], y[5], a1); a2 = fma(x[6], y[6], a2); a3 = fma(x[7], y[7], a3);
                }
                const double tot = warp_sum((a0 + a1) + (a2 + a3));
                if (lane == 0) {
                    const double qr = s_vec[r];
                    double x = fma(-sigma, qr, tot);
                    x = fma(-beta_prev, s_prev[warp], x);
                    wnew[r] = x;
                    s_prev[warp] = qr;
                }
            }
        }
        RES_PROF(0);
        if (!grid.sync()) return;
        RES_PROF(1);

        double alpha_j = sigma, s2_last = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // ---- S1: h[i] = V_i . w for i <= j, and (first pass only) h[j+1] = w . w ----
            const int i_end = (pass == 0) ? j + 1 : j;
            const int grp = warp >> 3, gw = warp & 7;
            for (int i0 = 0; i0 <= i_end; i0 += 2 * nb) {
                const int i = i0 + grp * nb + b;
                double sacc = 0.0;
                if (i <= i_end) {
                    const double* v = (i <= j) ? p.V + (size_t)i * m : wnew;
                    double x[8], y[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int rr = gw * 32 + lane + 256 * u;
                        const bool ok = rr < m;
                        x[u] = ok ? ld_cg(v + rr) : 0.0;
                        y[u] = ok ? ld_cg(wnew + rr) : 0.0;
                    }
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int u = 0; u < 8; u += 2) { a0 = fma(x[u], y[u], a0); a1 = fma(x[u + 1], y[u + 1], a1); }
                    sacc = warp_sum(a0 + a1);
                }
                __syncthreads();
                if (lane == 0) s_red[warp] = sacc;
                __syncthreads();
                if (gw == 0 && lane == 0 && i <= i_end) {
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) tot += s_red[grp * 8 + q];
                    p.h[i] = tot;
                }
            }
            RES_PROF(2);
            if (!grid.sync()) return;
            RES_PROF(3);

            // ---- S2: w -= V h on the owned rows; partial |w|^2.  All loads of the phase are issued together. ----
            // h is read by every thread of every CTA: stage it once per CTA (148 x 512 threads loading the same dozen
            // cache lines directly made this phase 3x slower -- an L2 slice hot spot).
            const int rloc = tid & 15, sub = tid >> 4;
            const int r = row0 + rloc;
            const bool h_sm = (j + 2 <= LAN_H);
            double wv = 0.0;
            double yv[8];                                   // basis entries V[ii][r], ii = sub + 32u (first 256 vectors)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int ii = sub + 32 * u;
                yv[u] = (rloc < n_own && ii <= j) ? ld_cg(p.V + (size_t)ii * m + r) : 0.0;
            }
            if (rloc < n_own && sub == 0) wv = ld_cg(wnew + r);
            if (h_sm) {                                     // same round trip as the loads above
                for (int i = tid; i <= j + 1; i += LAN_THREADS) s_h[i] = ld_cg(p.h + i);
                __syncthreads();
            }
            const double hj = h_sm ? s_h[j] : ld_cg(p.h + j);
            const double ww = (pass == 0) ? (h_sm ? s_h[j + 1] : ld_cg(p.h + j + 1)) : 0.0;
            double sacc = 0.0;
            if (rloc < n_own) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const int i0 = sub + 32 * u, i1 = i0 + 32;
                    const double h0 = i0 <= j ? (h_sm ? s_h[i0] : ld_cg(p.h + i0)) : 0.0;
                    const double h1 = i1 <= j ? (h_sm ? s_h[i1] : ld_cg(p.h + i1)) : 0.0;
                    a0 = fma(h0, yv[u], a0); a1 = fma(h1, yv[u + 1], a1);
                }
                for (int i = sub + 256; i <= j; i += 32)     // more than 256 basis vectors: the tail, rarely reached
                    a0 = fma(h_sm ? s_h[i] : ld_cg(p.h + i), ld_cg(p.V + (size_t)i * m + r), a0);
                sacc = a0 + a1;
            }
            __syncthreads();
            s_red[sub * 16 + rloc] = sacc;
            __syncthreads();
            double nrm = 0.0;
            if (tid < 16 && tid < n_own) {
                double tot = 0.0;
#pragma unroll
                for (int q = 0; q < 32; ++q) tot += s_red[q * 16 + tid];
                const double x = wv - tot;
                wnew[r] = x;
                nrm = x * x;
            }
            if (tid < 16) s_nrm[tid] = nrm;
            __syncthreads();
            alpha_j += hj;
            if (tid == 0) {
                double tot = 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) tot += s_nrm[q];
                p.part[b] = tot;
                if (b == 0) p.alpha[j - 1] = alpha_j;
            }
            RES_PROF(4);
            if (!grid.sync()) return;
            RES_PROF(5);
            // norm of the updated vector; prefetch the new vector elements in the same round trip
            double nx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int c = tid + LAN_THREADS * u; nx[u] = c < m ? ld_cg(wnew + c) : 0.0; }
            s2_last = sum_partials(p.part, nb, s_tmp);
            // DGKS ("twice is enough"): second pass iff |w_after|^2 < |w_before|^2 / 2
            if (pass == 0 && !(s2_last < 0.5 * ww)) {
                const double bt = sqrt(s2_last);
                const double inv = bt > 0.0 ? 1.0 / bt : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {                               // q_{j+1} for all columns, locally
                    const int c = tid + LAN_THREADS * u;
                    if (c < m) {
                        const double x = nx[u] * inv;
                        s_vec[c] = x;
                        if (c >= row0 && c < row1) p.V[(size_t)(j + 1) * m + c] = x;
                    }
                }
                break;
            }
            if (pass == 1) {
                const double bt = sqrt(s2_last);
                const double inv = bt > 0.0 ? 1.0 / bt : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = tid + LAN_THREADS * u;
                    if (c < m) {
                        const double x = nx[u] * inv;
                        s_vec[c] = x;
                        if (c >= row0 && c < row1) p.V[(size_t)(j + 1) * m + c] = x;
                    }
                }
            }
        }
        beta_prev = sqrt(s2_last);
        sigma = alpha_j;
        if (b == 0 && tid == 0) p.beta[j - 1] = beta_prev;
        __syncthreads();                                                    // s_vec complete before the next S0
        cur ^= 1;
        RES_PROF(6);
    }
}

// ------------------------------------------------------------------------------------------------
// Small-m variant (the whole Lanczos vector fits in shared memory): ONE grid barrier per step.
//
//   S0   w' = Ms q_j - sigma q_j - beta_{j-1} q_{j-1} on the owned rows (q_j in shared memory), written to
//        global, with three per-CTA partial sums: q_j.w', v0.w', |w'|^2
//   --- grid barrier ---
//   every CTA sums the partials (fixed order): alpha_j = sigma + q_j.w', gamma = v0.w',
//        beta_j^2 = |w'|^2 - (q_j.w')^2 - gamma^2, and rebuilds the next vector for ALL columns in its own
//        shared memory: q_{j+1} = (w' - (q_j.w') q_j - gamma v0) / beta_j   -> no second barrier.
//
// Orthogonality against older vectors is maintained by PARTIAL reorthogonalisation (Simon 1984; as in
// PROPACK): every CTA advances the omega-recurrence that bounds q_{j+1}.q_k from the alphas and betas; when
// the bound passes sqrt(eps) the current and the next vector are orthogonalised explicitly against the whole
// basis (the S1/S2 phases of the kernel above, with the DGKS second pass).  The basis stays semi-orthogonal
// (<= ~1e-8), for which the tridiagonal matrix equals the exact projection to working precision and the
// recurrence A V = V T + beta v e^T still holds to rounding, so every Ritz pair keeps its small residual
// (mutual orthogonality of the returned vectors is then residual/gap, ~1e-9, instead of 1e-15).  The locked vector v0 (eigenvalue 1, far outside the rest of the
// spectrum) is projected out at every step.
// ------------------------------------------------------------------------------------------------
struct ProParams {
    const double* Ms; int m;
    double* V;            // basis: V[0] = v0, V[j] = q_j (explicit, normalised; written by the row owners)
    double* wbuf;         // [2][m] ping-pong w' / w''
    double* h;            // Gram-Schmidt coefficients of the explicit reorthogonalisation
    double* part;         // [2][3][gridDim.x] ping-pong partial sums
    double* alpha; double* beta;
    double* omega;        // [2][cap] persistent omega_j, omega_{j-1} between batches; omega[2*cap] = force flag
    int omega_cap;
    int j_start, n_steps;
    unsigned int* bar;
    int* stats;           // stats[0] += explicit reorthogonalisations
};

constexpr int PRO_MAX_M = 8192;

__global__ void __launch_bounds__(LAN_THREADS, 2) lanczos_pro_kernel(ProParams p) {
    GridBarrier grid(p.bar);
    extern __shared__ double pro_smem[];
    const int m = p.m, cap = p.omega_cap;
    double* s_q = pro_smem;                       // [m]   current Lanczos vector q_j
    double* s_om0 = s_q + m;                      // [cap] omega_{j,k}
    double* s_om1 = s_om0 + cap;                  // [cap] omega_{j-1,k}
    double* s_om2 = s_om1 + cap;                  // [cap] scratch for omega_{j+1,k}
    __shared__ double s_red[32 * 16];
    __shared__ double s_tmp[LAN_WARPS + 1];
    __shared__ double s_bc[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = gridDim.x, b = blockIdx.x;
    const int chunk = (m + nb - 1) / nb;
    const int row0 = min(b * chunk, m), row1 = min(row0 + chunk, m);
    const double EPS = 1.1102230246251565e-16;
    const double DELTA = 1.4901161193847656e-08;  // sqrt(eps): semi-orthogonality level
    const double* v0 = p.V;

    // ---- load the persistent state: q_{j_start}, omega arrays, force flag ----
    int j = p.j_start;
    for (int c = tid; c < m; c += LAN_THREADS) s_q[c] = ld_cg(p.V + (size_t)j * m + c);
    for (int k = tid; k < cap; k += LAN_THREADS) { s_om0[k] = ld_cg(p.omega + k); s_om1[k] = ld_cg(p.omega + cap + k); }
    bool force = ld_cg(p.omega + 2 * cap) != 0.0;
    __syncthreads();
    int cur = 0;
    int n_reorth = 0;

    // alpha_{j-1} (shift predictor) and beta_{j-1}: from the previous launch at batch start, afterwards carried in
    // registers (every CTA computes them identically; block 0's global copies may not be visible yet).
    double sigma = 0.0, beta_prev = 0.0;
    if (j >= 2) { sigma = ld_cg(p.alpha + j - 2); beta_prev = ld_cg(p.beta + j - 2); }
    for (int step = 0; step < p.n_steps; ++step, ++j) {
        double* wcur = p.wbuf + (size_t)cur * m;
        double* pcur = p.part + (size_t)cur * 3 * nb;

        // ---- S0: owned rows of w' and the three partial sums ----
        double pa = 0.0, pg = 0.0, pn = 0.0;                 // lane 0 of each warp accumulates
        for (int rbase = row0; rbase < row1; rbase += LAN_WARPS) {
            const int r = rbase + warp;
            if (r < row1) {
                const double* a = p.Ms + (size_t)r * m;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int c = lane;
                for (; c + 224 < m; c += 256) {
                    const double x0 = __ldg(a + c), x1 = __ldg(a + c + 32), x2 = __ldg(a + c + 64), x3 = __ldg(a + c + 96),
                                 x4 = __ldg(a + c + 128), x5 = __ldg(a + c + 160), x6 = __ldg(a + c + 192), x7 = __ldg(a + c + 224);
                    a0 = fma(x0, s_q[c], a0);       a1 = fma(x1, s_q[c + 32], a1);
                    a2 = fma(x2, s_q[c + 64], a2);  a3 = fma(x3, s_q[c + 96], a3);
                    a0 = fma(x4, s_q[c + 128], a0); a1 = fma(x5, s_q[c + 160], a1);
                    a2 = fma(x6, s_q[c + 192], a2); a3 = fma(x7, s_q[c + 224], a3);
                }
                for (; c < m; c += 32) a0 = fma(__ldg(a + c), s_q[c], a0);
                const double tot = warp_sum((a0 + a1) + (a2 + a3));
                if (lane == 0) {
                    const double qr = s_q[r];
                    double x = fma(-sigma, qr, tot);
                    if (j >= 2) x = fma(-beta_prev, ld_cg(p.V + (size_t)(j - 1) * m + r), x);
                    wcur[r] = x;
                    pa = fma(qr, x, pa); pg = fma(__ldg(v0 + r), x, pg); pn = fma(x, x, pn);
                }
            }
        }
        __syncthreads();
        if (lane == 0) { s_red[warp] = pa; s_red[16 + warp] = pg; s_red[32 + warp] = pn; }
        __syncthreads();
        if (tid < 3) {
            double t = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < LAN_WARPS; ++w2) t += s_red[tid * 16 + w2];
            pcur[tid * nb + b] = t;
        }
        if (!grid.sync()) return;

        // ---- every CTA: the three global sums (identical arithmetic everywhere) ----
        const double qa = sum_partials(pcur, nb, s_tmp);
        const double ga = sum_partials(pcur + nb, nb, s_tmp);
        const double n2 = sum_partials(pcur + 2 * nb, nb, s_tmp);
        double beta2 = n2 - qa * qa - ga * ga;
        const double alpha_j = sigma + qa;

        // ---- omega recurrence: bound on q_{j+1}.q_k, k = 1 .. j (1-based Lanczos indices) ----
        // Needs beta_j; a cancelled beta_j^2 (tiny or negative) forces the explicit path, which recomputes it.
        bool need = force || !(beta2 > 1e-4 * n2) || (j + 2 >= cap);
        double beta_j = beta2 > 0.0 ? sqrt(beta2) : 0.0;
        if (!need) {
            double mx = 0.0;
            for (int k = tid + 1; k < j; k += LAN_THREADS) {           // k = 1 .. j-1
                const double ak = ld_cg(p.alpha + k - 1), bk = ld_cg(p.beta + k - 1);
                const double bkm1 = (k >= 2) ? ld_cg(p.beta + k - 2) : 0.0;
                const double om_k1 = (k + 1 == j) ? 1.0 : s_om0[k + 1];   // omega_{j,k+1}; omega_{j,j} = 1
                const double om_km1 = (k >= 2) ? s_om0[k - 1] : 0.0;
                double t = bk * om_k1 + (ak - alpha_j) * s_om0[k] + bkm1 * om_km1 - beta_prev * s_om1[k];
                t += copysign(EPS * (bk + beta_j) * 0.6, t);
                t /= beta_j;
                s_om2[k] = t;
                mx = fmax(mx, fabs(t));
            }
            // warp/block max
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            __syncthreads();
            if (lane == 0) s_tmp[warp] = mx;
            __syncthreads();
            mx = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < LAN_WARPS; ++w2) mx = fmax(mx, s_tmp[w2]);
            __syncthreads();
            if (mx > DELTA) need = true;
        }

        if (!need) {
            // ---- fast path: rebuild q_{j+1} for all columns locally ----
            const double inv = 1.0 / beta_j;
            if (b == 0 && tid == 0) { p.alpha[j - 1] = alpha_j; p.beta[j - 1] = beta_j; }
            for (int c = tid; c < m; c += LAN_THREADS) {
                const double x = (ld_cg(wcur + c) - qa * s_q[c] - ga * __ldg(v0 + c)) * inv;
                s_q[c] = x;
                if (c >= row0 && c < row1) p.V[(size_t)(j + 1) * m + c] = x;
            }
            for (int k = tid + 1; k < j; k += LAN_THREADS) { s_om1[k] = s_om0[k]; }
            __syncthreads();
            for (int k = tid + 1; k < j; k += LAN_THREADS) { s_om0[k] = s_om2[k]; }
            if (tid == 0) {
                s_om1[j] = 1.0;                                        // omega_{j,j}
                s_om0[j] = EPS * sqrt((double)m) * 0.6;                // omega_{j+1,j}
                s_om0[j + 1] = 1.0;
            }
            __syncthreads();
            force = false;
            sigma = alpha_j; beta_prev = beta_j;
        } else {
            // ---- explicit path: w'' = w' - qa q_j - ga v0 on the owned rows, then Gram-Schmidt against V[0..j] ----
            ++n_reorth;
            for (int r = row0 + tid; r < row1; r += LAN_THREADS)
                wcur[r] = ld_cg(wcur + r) - qa * s_q[r] - ga * __ldg(v0 + r);
            if (!grid.sync()) return;
            double alpha_fix = 0.0, s2_last = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                const int i_end = (pass == 0) ? j + 1 : j;          // slot j+1: w.w before the pass
                const int grp = warp >> 3, gw = warp & 7;
                for (int i0 = 0; i0 <= i_end; i0 += 2 * nb) {
                    const int i = i0 + grp * nb + b;
                    double sacc = 0.0;
                    if (i <= i_end) {
                        const double* v = (i <= j) ? p.V + (size_t)i * m : wcur;
                        double a0 = 0.0, a1 = 0.0;
                        int r = gw * 32 + lane;
                        for (; r + 768 < m; r += 1024) {
                            const double v0_ = ld_cg(v + r), v1 = ld_cg(v + r + 256), v2 = ld_cg(v + r + 512), v3 = ld_cg(v + r + 768);
                            const double w0 = ld_cg(wcur + r), w1 = ld_cg(wcur + r + 256), w2 = ld_cg(wcur + r + 512), w3 = ld_cg(wcur + r + 768);
                            a0 = fma(v0_, w0, a0); a1 = fma(v1, w1, a1); a0 = fma(v2, w2, a0); a1 = fma(v3, w3, a1);
                        }
                        for (; r < m; r += 256) a0 = fma(ld_cg(v + r), ld_cg(wcur + r), a0);
                        sacc = warp_sum(a0 + a1);
                    }
                    __syncthreads();
                    if (lane == 0) s_red[warp] = sacc;
                    __syncthreads();
                    if (gw == 0 && lane == 0 && i <= i_end) {
                        double tot = 0.0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) tot += s_red[grp * 8 + q];
                        p.h[i] = tot;
                    }
                }
                if (!grid.sync()) return;
                double nrm = 0.0;
                for (int rbase = row0; rbase < row1; rbase += 16) {
                    const int r = rbase + (tid & 15), sub = tid >> 4;
                    double sacc = 0.0;
                    if (r < row1)
                        for (int i = sub; i <= j; i += 32) sacc = fma(ld_cg(p.h + i), ld_cg(p.V + (size_t)i * m + r), sacc);
                    __syncthreads();
                    s_red[sub * 16 + (tid & 15)] = sacc;
                    __syncthreads();
                    if (tid < 16 && r < row1) {
                        double tot = 0.0;
#pragma unroll
                        for (int q = 0; q < 32; ++q) tot += s_red[q * 16 + tid];
                        const double x = ld_cg(wcur + r) - tot;
                        wcur[r] = x;
                        nrm = fma(x, x, nrm);
                    }
                }
                __syncthreads();
                if (tid < 16) s_red[tid] = nrm;
                __syncthreads();
                if (tid == 0) {
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < 16; ++q) tot += s_red[q];
                    pcur[b] = tot;                                   // reuse slot 0 of this step's partials
                }
                alpha_fix += ld_cg(p.h + j);
                if (!grid.sync()) return;
                s2_last = sum_partials(pcur, nb, s_tmp);
                if (pass == 0 && !(s2_last < 0.5 * ld_cg(p.h + j + 1))) break;
            }
            beta_j = sqrt(s2_last);
            const double inv = beta_j > 0.0 ? 1.0 / beta_j : 0.0;
            if (b == 0 && tid == 0) { p.alpha[j - 1] = alpha_j + alpha_fix; p.beta[j - 1] = beta_j; }
            for (int c = tid; c < m; c += LAN_THREADS) {
                const double x = ld_cg(wcur + c) * inv;
                s_q[c] = x;
                if (c >= row0 && c < row1) p.V[(size_t)(j + 1) * m + c] = x;
            }
            // orthogonality restored for q_{j+1}; standard practice: also do the next vector explicitly
            for (int k = tid + 1; k <= j; k += LAN_THREADS) { s_om1[k] = (k == j) ? 1.0 : (force ? EPS : s_om0[k]); s_om0[k] = EPS; }
            if (tid == 0) s_om0[j + 1] = 1.0;
            __syncthreads();
            force = !force;
            sigma = alpha_j + alpha_fix; beta_prev = beta_j;
        }
        cur ^= 1;
    }
    // ---- persist the state for the next batch ----
    if (b == 0) {
        for (int k = tid; k < cap; k += LAN_THREADS) { p.omega[k] = s_om0[k]; p.omega[cap + k] = s_om1[k]; }
        if (tid == 0) { p.omega[2 * cap] = force ? 1.0 : 0.0; atomicAdd(p.stats, n_reorth); }
    }
}

// beta = |w| ; out = w / beta (single block).
__global__ void __launch_bounds__(1024) normalize_kernel(const double* __restrict__ w, int m, double* __restrict__ out,
                                                         double* __restrict__ beta_slot) {
    __shared__ double sh[32];
    __shared__ double nrm;
    double s = 0.0;
    for (int r = threadIdx.x; r < m; r += blockDim.x) { const double x = w[r]; s = fma(x, x, s); }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) { nrm = sqrt(s); if (beta_slot) *beta_slot = nrm; }
    __syncthreads();
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int r = threadIdx.x; r < m; r += blockDim.x) out[r] = w[r] * inv;
}

__global__ void __launch_bounds__(256) init_kernel(const double* __restrict__ deg, int m, double* __restrict__ w_sqrt,
                                                   double* __restrict__ w_rand) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    w_sqrt[r] = sqrt(deg[r]);
    uint32_t x = (uint32_t)r * 2654435761u + 0x9E3779B9u;     // deterministic start vector
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    w_rand[r] = (double)x * (1.0 / 4294967296.0) - 0.5;
}

// U[r][1+i] = sum_l V_{1+l}[r] S[l][i]; U[r][0] = V_0[r].  One thread per (row, Ritz vector).
__global__ void __launch_bounds__(128) ritz_kernel(const double* __restrict__ V, int m, int steps,
                                                   const double* __restrict__ S, int kw, double* __restrict__ U, int k) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;                       // 0 -> steady state, 1.. -> Ritz vector i-1
    if (r >= m) return;
    if (i == 0) { U[(size_t)r * k] = V[r]; return; }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int l = 0;
    for (; l + 3 < steps; l += 4) {
        a0 = fma(__ldg(V + (size_t)(1 + l) * m + r), __ldg(S + (size_t)l * kw + (i - 1)), a0);
        a1 = fma(__ldg(V + (size_t)(2 + l) * m + r), __ldg(S + (size_t)(l + 1) * kw + (i - 1)), a1);
        a2 = fma(__ldg(V + (size_t)(3 + l) * m + r), __ldg(S + (size_t)(l + 2) * kw + (i - 1)), a2);
        a3 = fma(__ldg(V + (size_t)(4 + l) * m + r), __ldg(S + (size_t)(l + 3) * kw + (i - 1)), a3);
    }
    for (; l < steps; ++l) a0 = fma(__ldg(V + (size_t)(1 + l) * m + r), __ldg(S + (size_t)l * kw + (i - 1)), a0);
    U[(size_t)r * k + i] = (a0 + a1) + (a2 + a3);
}

// Normalise each column to unit length and make its largest-magnitude entry positive (first on ties).
// V[0] = sqrt(deg) / |sqrt(deg)| (the locked vector) and, with_start, V[1] = the deterministic start vector of init_kernel
// orthogonalised twice against V[0] and normalised.  One block: seven launches of the kernels above in one.
__global__ void __launch_bounds__(1024) start_vectors_kernel(const double* __restrict__ deg, int m, double* __restrict__ V, int with_start) {
    __shared__ double sh[32];
    __shared__ double s_val;
    const int tid = threadIdx.x;
    double* V0 = V; double* V1 = V + m;
    double s = 0.0;
    for (int r = tid; r < m; r += 1024) { const double x = sqrt(deg[r]); V0[r] = x; s = fma(x, x, s); }
    s = block_sum(s, sh);
    if (tid == 0) { const double n = sqrt(s); s_val = n > 0.0 ? 1.0 / n : 0.0; }
    __syncthreads();
    const double inv0 = s_val;
    for (int r = tid; r < m; r += 1024) V0[r] *= inv0;
    if (!with_start) return;
    for (int r = tid; r < m; r += 1024) {
        uint32_t x = (uint32_t)r * 2654435761u + 0x9E3779B9u;
        x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
        V1[r] = (double)x * (1.0 / 4294967296.0) - 0.5;
    }
    for (int pass = 0; pass < 2; ++pass) {
        double d = 0.0;
        for (int r = tid; r < m; r += 1024) d = fma(V0[r], V1[r], d);
        __syncthreads();
        d = block_sum(d, sh);
        if (tid == 0) s_val = d;
        __syncthreads();
        const double h = s_val;
        for (int r = tid; r < m; r += 1024) V1[r] = fma(-h, V0[r], V1[r]);
    }
    double q = 0.0;
    for (int r = tid; r < m; r += 1024) { const double x = V1[r]; q = fma(x, x, q); }
    __syncthreads();
    q = block_sum(q, sh);
    if (tid == 0) { const double n = sqrt(q); s_val = n > 0.0 ? 1.0 / n : 0.0; }
    __syncthreads();
    const double inv1 = s_val;
    for (int r = tid; r < m; r += 1024) V1[r] *= inv1;
}

__global__ void __launch_bounds__(256) fix_columns_kernel(double* __restrict__ U, int m, int k) {
    __shared__ double sh[32];
    __shared__ double s_scale;
    const int col = blockIdx.x;
    double s = 0.0, best = -1.0; int best_r = 0x7fffffff;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const double x = U[(size_t)r * k + col];
        s = fma(x, x, s);
        if (fabs(x) > best) { best = fabs(x); best_r = r; }
    }
    // arg-max of |x| over the block (ties -> smallest row)
    __shared__ double sb[256]; __shared__ int si[256];
    sb[threadIdx.x] = best; si[threadIdx.x] = best_r;
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        double b = -1.0; int br = 0x7fffffff;
        for (int t = 0; t < blockDim.x; ++t)
            if (sb[t] > b || (sb[t] == b && si[t] < br)) { b = sb[t]; br = si[t]; }
        const double nrm = sqrt(s);
        double sc = nrm > 0.0 ? 1.0 / nrm : 0.0;
        if (br < m && U[(size_t)br * k + col] < 0.0) sc = -sc;
        s_scale = sc;
    }
    __syncthreads();
    const double sc = s_scale;
    for (int r = threadIdx.x; r < m; r += blockDim.x) U[(size_t)r * k + col] *= sc;
}

// res2[col] += |Ms u - theta u|^2 partials for selected columns: grid (ceil(m/8), ncols); deterministic per-block partials.
__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ Ms, int m, const double* __restrict__ U, int k,
                                                       const int* __restrict__ cols, const double* __restrict__ theta,
                                                       double* __restrict__ partial /*[ncols][gridDim.x]*/) {
    __shared__ double sh[8];
    const int col = cols[blockIdx.y];
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    double r2 = 0.0;
    if (row < m) {
        const double* a = Ms + (size_t)row * m;
        double s = 0.0;
        for (int c = lane; c < m; c += 32) s = fma(__ldg(a + c), U[(size_t)c * k + col], s);
        s = warp_sum(s);
        const double d = s - theta[blockIdx.y] * U[(size_t)row * k + col];
        r2 = d * d;
    }
    if (lane == 0) sh[threadIdx.x >> 5] = r2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// Host-side symmetric tridiagonal eigen-solver (T = tridiag(beta, alpha, beta), size n).
// ---------------------------------------------------------------------------------------------
static int sturm_count_below(const std::vector<double>& a, const std::vector<double>& b, int n, double x, double pivmin) {
    int cnt = 0;
    double q = a[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) ++cnt;
    for (int i = 1; i < n; ++i) {
        q = a[i] - x - b[i - 1] * b[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        if (q < 0.0) ++cnt;
    }
    return cnt;   // number of eigenvalues < x
}

// The `want` largest eigenvalues (descending) by bisection.
static void tridiag_top_eigenvalues(const std::vector<double>& a, const std::vector<double>& b, int n, int want,
                                    std::vector<double>& theta) {
    double lo = a[0], hi = a[0], tn = 0.0;
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < n - 1 ? fabs(b[i]) : 0.0);
        lo = std::min(lo, a[i] - r); hi = std::max(hi, a[i] + r);
        tn = std::max(tn, fabs(a[i]) + r);
    }
    const double pivmin = std::max(tn * tn * 1e-300, 1e-300);
    const double span = hi - lo;
    lo -= 1e-12 * span + 1e-300; hi += 1e-12 * span + 1e-300;
    theta.assign(want, 0.0);
    for (int i = 0; i < want; ++i) {
        // i-th largest: smallest x with count_below(x) >= n - i  -> eigenvalue index n-1-i (ascending)
        double l = lo, h = (i == 0) ? hi : theta[i - 1] + 1e-14 * tn;
        const int target = n - i;
        for (int it = 0; it < 200; ++it) {
            const double mid = 0.5 * (l + h);
            if (mid <= l || mid >= h) break;
            if (sturm_count_below(a, b, n, mid, pivmin) >= target) h = mid; else l = mid;
        }
        theta[i] = 0.5 * (l + h);
    }
}

// Solve (T - x I) y = rhs in place by Gaussian elimination with partial pivoting (tridiagonal).
static void tridiag_shift_solve(const std::vector<double>& a, const std::vector<double>& b, int n, double x,
                                double tiny, std::vector<double>& y, std::vector<double>& d, std::vector<double>& du,
                                std::vector<double>& du2, std::vector<double>& dl) {
    for (int i = 0; i < n; ++i) d[i] = a[i] - x;
    for (int i = 0; i < n - 1; ++i) { du[i] = b[i]; dl[i] = b[i]; }
    for (int i = 0; i < n; ++i) du2[i] = 0.0;
    for (int i = 0; i < n - 1; ++i) {
        if (fabs(d[i]) >= fabs(dl[i])) {
            if (d[i] == 0.0) d[i] = tiny;
            const double f = dl[i] / d[i];
            d[i + 1] -= f * du[i];
            y[i + 1] -= f * y[i];
        } else {                                    // swap rows i and i+1
            const double f = d[i] / dl[i];
            d[i] = dl[i];
            const double t = d[i + 1];
            d[i + 1] = du[i] - f * t;
            if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du2[i]; }
            du[i] = t;
            const double ty = y[i]; y[i] = y[i + 1]; y[i + 1] = ty - f * y[i + 1];
        }
    }
    if (d[n - 1] == 0.0) d[n - 1] = tiny;
    y[n - 1] /= d[n - 1];
    if (n > 1) y[n - 2] = (y[n - 2] - du[n - 2] * y[n - 1]) / d[n - 2];
    for (int i = n - 3; i >= 0; --i) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / d[i];
}

// Eigenvectors of T for the eigenvalues theta (descending) by inverse iteration; vectors of close
// eigenvalues are orthogonalised against each other.  S: [n][want] row-major.
static void tridiag_eigenvectors(const std::vector<double>& a, const std::vector<double>& b, int n,
                                 const std::vector<double>& theta, std::vector<double>& S) {
    const int want = (int)theta.size();
    double tn = 0.0;
    for (int i = 0; i < n; ++i) tn = std::max(tn, fabs(a[i]) + (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < n - 1 ? fabs(b[i]) : 0.0));
    const double tiny = std::max(tn * 2.2e-16, 1e-300);
    const double cluster_tol = 1e-3 * tn;
    S.assign((size_t)n * want, 0.0);
    std::vector<double> y(n), d(n), du(n), du2(n), dl(n);
    std::vector<std::vector<double>> vecs(want, std::vector<double>(n));
    int cluster_start = 0;
    for (int i = 0; i < want; ++i) {
        if (i > 0 && fabs(theta[i] - theta[i - 1]) > cluster_tol) cluster_start = i;
        // perturb repeated eigenvalues slightly so that the shifted systems differ
        double x = theta[i];
        if (i > cluster_start && fabs(x - theta[i - 1]) < 10.0 * tiny) x = theta[i - 1] - 10.0 * tiny * (i - cluster_start);
        uint32_t seed = 12345u + 7919u * (uint32_t)i;
        for (int r = 0; r < n; ++r) { seed = seed * 1664525u + 1013904223u; y[r] = ((double)(seed >> 8) / 16777216.0) - 0.5; }
        for (int it = 0; it < 6; ++it) {
            tridiag_shift_solve(a, b, n, x, tiny, y, d, du, du2, dl);
            for (int pass = 0; pass < 2; ++pass)
                for (int p = cluster_start; p < i; ++p) {
                    double dot = 0.0;
                    for (int r = 0; r < n; ++r) dot += vecs[p][r] * y[r];
                    for (int r = 0; r < n; ++r) y[r] -= dot * vecs[p][r];
                }
            double nrm = 0.0;
            for (int r = 0; r < n; ++r) nrm += y[r] * y[r];
            nrm = sqrt(nrm);
            if (!(nrm > 0.0) || !std::isfinite(nrm)) { for (int r = 0; r < n; ++r) y[r] = (r == (i % n)) ? 1.0 : 0.0; nrm = 1.0; }
            for (int r = 0; r < n; ++r) y[r] /= nrm;
        }
        vecs[i] = y;
        for (int r = 0; r < n; ++r) S[(size_t)r * want + i] = y[r];
    }
}

}  // namespace lanczos

// Residual estimates |beta_j s_{j,i}| of the leading Ritz pairs of T = tridiag(b, a, b); early exit.
bool ritz_converged(const std::vector<double>& a, const std::vector<double>& b, int n, int kw, int guard,
                    double tol, std::vector<double>& theta, std::vector<double>& S) {
    using namespace lanczos;
    const int want = std::min(kw + guard, n);
    tridiag_top_eigenvalues(a, b, n, want, theta);
    tridiag_eigenvectors(a, b, n, theta, S);
    const double scale = std::max(fabs(theta[0]), 1e-300);
    const double beta_last = b[n - 1];
    if (n < kw) return false;
    for (int i = want - 1; i >= 0; --i) {                 // the smallest wanted pair converges last
        const double res = fabs(beta_last * S[(size_t)(n - 1) * want + i]);
        if (!(res <= (i < kw ? tol : 1e-5) * scale)) return false;
    }
    return true;
}

static int lanczos_attempt(esper_ctx* c, int m, int k, double tol, int max_iter, bool use_pro, double* val_h, double* vec_h,
                           int* iters_out, bool* suspicious, bool plain, bool full = false) {
    using namespace lanczos;
    constexpr int KMAX = 64;
    const int kw = k - 1;                               // non-trivial pairs wanted
    int guard = (kw > 0) ? std::min(2, m - 1 - kw) : 0;
    if (guard < 0) guard = 0;
    const int cap = m - 1;                              // Krylov dimension of the deflated operator
    const double* Ms = c->markov.as<double>();
    cudaStream_t st = c->stream;
    *suspicious = false;

    // cooperative grid: one CTA per SM (all must be co-resident for the grid barriers)
    int coop = 0, per_sm = 0;
    ESPER_CUDA(c, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    const int ocap = max_iter + 4;
    const size_t pro_smem = sizeof(double) * ((size_t)m + 3 * (size_t)ocap);
    if (use_pro) {
        if (!c->attr_lanczos_pro) {
            ESPER_CUDA(c, cudaFuncSetAttribute(lanczos_pro_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            c->attr_lanczos_pro = true;
        }
        ESPER_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_pro_kernel, LAN_THREADS, pro_smem));
    }
    // Full reorthogonalisation: the resident kernel when the CTA's rows of Ms fit on chip, else the streaming one.
    int rows_cached = 0;
    size_t lan_smem = 0;
    bool resident = false, small = false;
    if (!use_pro) {
        const int chunk_rows = ceil_div(m, c->sm_count);
        const size_t budget = (size_t)(227 - 16) * 1024 - sizeof(double) * LAN_CH;  // per-CTA limit minus static arrays
        small = (m <= LAN_CH && chunk_rows <= 16);          // whole vector in shared memory, one warp per owned row
        if (small && (c->lanczos_mode == 0 || c->lanczos_mode == 3)) {
            rows_cached = (int)std::min<size_t>(std::min(chunk_rows, 14), budget / (sizeof(double) * (size_t)m));
            resident = chunk_rows <= rows_cached + 2;
        }
        if (resident) {
            lan_smem = sizeof(double) * ((size_t)LAN_CH + (size_t)rows_cached * m);
            // The opt-in is a property of the function in this process (not of the ctx): a smaller value set through
            // another ctx would LOWER the limit and make this launch fail.  Opt in to the device maximum, once per ctx.
            if (!c->lanczos_smem_set) {
                int optin = 0;
                cudaFuncAttributes fa;
                ESPER_CUDA(c, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
                ESPER_CUDA(c, cudaFuncGetAttributes(&fa, lanczos_resident_kernel<true>));
                ESPER_CUDA(c, cudaFuncSetAttribute(lanczos_resident_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   optin - (int)fa.sharedSizeBytes));
                c->lanczos_smem_set = optin - (long long)fa.sharedSizeBytes;
            }
            if ((long long)lan_smem > c->lanczos_smem_set)
                return fail(c, ESPER_E_STATE, "lanczos: %zu bytes of shared memory requested, %lld available", lan_smem, c->lanczos_smem_set);
            ESPER_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_resident_kernel<true>, LAN_THREADS, lan_smem));
        } else if (small) {
            rows_cached = 0;
            lan_smem = sizeof(double) * (size_t)LAN_CH;
            ESPER_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_resident_kernel<false>, LAN_THREADS, lan_smem));
        } else {
            rows_cached = 0;
            ESPER_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_steps_kernel, LAN_THREADS, 0));
        }
    }
    if (!coop || per_sm < 1) return fail(c, ESPER_E_CUDA, "cooperative launch unavailable on this device");
    const int grid = c->sm_count;

    const int stride = max_iter + 8;
    ESPER_RESERVE(c, c->lan_V, sizeof(double) * (size_t)(max_iter + 3) * m);
    ESPER_RESERVE(c, c->lan_w, sizeof(double) * (size_t)m * 2);
    ESPER_RESERVE(c, c->lan_h, sizeof(double) * ((size_t)stride * 3 + 6 * (size_t)grid + 2 * (size_t)ocap + 32));
    if (c->pin.reserve(sizeof(double) * ((size_t)stride * 2 * 2 + 4)) != 0) return fail(c, ESPER_E_NOMEM, "pinned alloc failed");
    double* V = c->lan_V.as<double>();
    double* w = c->lan_w.as<double>();
    double* w2 = w + m;
    double* h = c->lan_h.as<double>();
    double* alpha_d = h + stride;                       // alpha[stride], beta[stride] contiguous
    double* beta_d = alpha_d + stride;
    double* part_d = beta_d + stride;                   // [6 * grid]
    double* omega_d = part_d + 6 * (size_t)grid;        // [2 * ocap + 1]
    unsigned int* bar_d = reinterpret_cast<unsigned int*>(omega_d + 2 * (size_t)ocap + 2);
    int* stats_d = reinterpret_cast<int*>(bar_d + 2);
    long long* prof_d = getenv("ESPER_DEBUG_LANCZOS") ? reinterpret_cast<long long*>(stats_d + 2) : nullptr;   // 8 counters, zeroed below
    ESPER_CUDA(c, cudaMemsetAsync(alpha_d, 0, sizeof(double) * (size_t)stride * 2, st));
    ESPER_CUDA(c, cudaMemsetAsync(omega_d, 0, sizeof(double) * (2 * (size_t)ocap + 2) + 16 + 80, st));
    if (use_pro) {
        const double one = 1.0;
        ESPER_CUDA(c, cudaMemcpyAsync(omega_d + 1, &one, sizeof(double), cudaMemcpyHostToDevice, st));   // omega_{1,1} = 1
    }

    // locked vector v0 and the start vector q1 (orthogonal to v0): one single-block launch
    { LaunchScope ls(c, "lanczos"); start_vectors_kernel<<<1, 1024, 0, st>>>(c->degree.as<double>(), m, V, kw > 0 ? 1 : 0); }

    // Default schedule (esper_dmap_config 0): the barrier-free chain kernel with on-device convergence tests (k4_chain.cu).
    // It returns +1 for shapes it does not cover or when it did not converge within its shared-memory basis capacity;
    // the barrier kernels below then run the whole iteration from the same start vector.
    if (!use_pro && kw > 0 && c->lanczos_mode == 0 && !full && !getenv("ESPER_LANCZOS_NO_CHAIN")) {
        ESPER_RESERVE(c, c->lan_small, sizeof(double) * ((size_t)m * k + (size_t)(max_iter + 8) * KMAX + 2 * KMAX * KMAX + 4096 + (size_t)ceil_div(m, 8) * 2));
        double* Uc = c->lan_small.as<double>();
        int rcc = run_lanczos_chain(c, m, k, guard, tol, max_iter, V, Uc, !plain);
        if (rcc < 0) return rcc;
        if (rcc == 0) {
            if (vec_h) {
                if (plain) ESPER_CUDA(c, cudaMemcpy2DAsync(vec_h, sizeof(double) * kw, Uc + 1, sizeof(double) * k, sizeof(double) * kw, m, cudaMemcpyDeviceToHost, st));
                else ESPER_CUDA(c, cudaMemcpyAsync(vec_h, Uc, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, st));
            }
            ESPER_CUDA(c, cudaStreamSynchronize(st));
            std::vector<double> th((size_t)kw);
            rcc = chain_result(c, k, th.data(), iters_out);
            if (rcc < 0) return rcc;
            if (rcc == 0) {
                if (val_h) {
                    if (plain) for (int i = 0; i < kw; ++i) val_h[i] = th[i];
                    else { val_h[0] = 1.0; for (int i = 0; i < kw; ++i) val_h[1 + i] = th[i] * th[i]; }
                }
                return ESPER_OK;
            }
            c->lanczos_fallbacks++;
        }
    }

    // Batches of steps; the host tests batch n while batch n+1 already runs (speculatively).
    std::vector<double> a, b, theta, S;
    int steps_done = 0;          // steps whose alpha/beta have been checked on the host
    int steps_issued = 0;
    bool converged = (kw == 0);
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (!converged) { ESPER_CUDA(c, cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming)); ESPER_CUDA(c, cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming)); }
    double* pin = c->pin.as<double>();
    unsigned int* pin_flag = reinterpret_cast<unsigned int*>(pin + (size_t)stride * 4);
    int pending_steps[2] = {0, 0};
    int slot = 0;
    auto issue_batch = [&](int n_steps, int sl) -> int {
        cudaMemsetAsync(bar_d, 0, 2 * sizeof(unsigned int), st);
        cudaError_t e;
        {
            LaunchScope ls(c, "lanczos");
            if (use_pro) {
                ProParams sp{Ms, m, V, w, h, part_d, alpha_d, beta_d, omega_d, ocap, steps_issued + 1, n_steps, bar_d, stats_d};
                void* args[] = {&sp};
                e = cudaLaunchCooperativeKernel((void*)lanczos_pro_kernel, dim3(grid), dim3(LAN_THREADS), args, pro_smem, st);
            } else {
                StepParams sp{Ms, m, V, w, h, part_d, alpha_d, beta_d, steps_issued + 1, n_steps, bar_d, rows_cached, prof_d};
                void* args[] = {&sp};
                void* fn = resident ? (void*)lanczos_resident_kernel<true> : small ? (void*)lanczos_resident_kernel<false> : (void*)lanczos_steps_kernel;
                e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LAN_THREADS), args, lan_smem, st);
            }
        }
        if (e != cudaSuccess) { cudaGetLastError(); return fail(c, ESPER_E_CUDA, "cooperative launch: %s", cudaGetErrorString(e)); }
        steps_issued += n_steps;
        cudaMemcpyAsync(pin + (size_t)sl * stride * 2, alpha_d, sizeof(double) * (size_t)stride * 2, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(pin_flag + 2 * sl, bar_d, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
        cudaEventRecord(ev[sl], st);
        pending_steps[sl] = steps_issued;
        return ESPER_OK;
    };
    int rc = ESPER_OK;
    if (!converged) {
        const int first = full ? std::min(max_iter, 64) : std::min(max_iter, std::max(3 * (kw + guard) + 8, 32));
        const int batch = full ? 64 : (resident ? 24 : 16);      // resident rows are reloaded per launch: longer batches
        // Mode 2 (opt-in, not yet run on hardware): no speculative batch.  The batch launched while the host tests the
        // previous one runs to its end even when that test finds convergence (measured on the CPU for the benchmark
        // matrix: the test first passes at 161 steps, is seen at 169 and the GPU executes 185).  With several PDs in
        // flight the gap a non-speculative schedule leaves on this stream is filled by the other streams, so batches of
        // 8 steps issued only after a failed test save ~20 steps per PD.
        const bool speculate = (c->lanczos_mode != 2);
        rc = issue_batch(first, slot);
        while (rc == ESPER_OK) {
            const int cur_slot = slot;
            const bool more = steps_issued < max_iter;
            const bool ahead = more && speculate;
            if (ahead) { slot ^= 1; rc = issue_batch(std::min(batch, max_iter - steps_issued), slot); if (rc) break; }
            cudaError_t e = cudaEventSynchronize(ev[cur_slot]);
            if (e != cudaSuccess) { cudaGetLastError(); rc = fail(c, ESPER_E_CUDA, "lanczos: %s", cudaGetErrorString(e)); break; }
            if (pin_flag[2 * cur_slot + 1] != 0u) { rc = fail(c, ESPER_E_CUDA, "lanczos: grid barrier timed out (CTAs not co-resident?)"); break; }
            steps_done = pending_steps[cur_slot];
            const double* ab = pin + (size_t)cur_slot * stride * 2;
            a.assign(ab, ab + steps_done);
            b.assign(ab + stride, ab + stride + steps_done);
            if (getenv("ESPER_DEBUG_LANCZOS")) {
                fprintf(stderr, "[lanczos] pro=%d steps=%d issued=%d slot=%d\n", (int)use_pro, steps_done, steps_issued, cur_slot);
                if (atoi(getenv("ESPER_DEBUG_LANCZOS")) > 1) for (int i = 0; i < steps_done; ++i) fprintf(stderr, "  %3d alpha=%.6e beta=%.6e\n", i, a[i], b[i]);
            }
            const double scale0 = std::max(fabs(a[0]), 1e-300);
            bool breakdown = !(b[steps_done - 1] > 1e-14 * scale0);
            if (full) {                                          // the Krylov space may end inside a batch: keep the steps before it
                for (int i = 0; i < steps_done; ++i)
                    if (!(b[i] > 1e-14 * scale0)) { steps_done = i + 1; breakdown = true; break; }
                a.resize(steps_done); b.resize(steps_done);
            }
            // full spectrum: the iteration runs until the Krylov space is exhausted, no Ritz test in between
            if (full ? (breakdown || steps_done >= cap) : (ritz_converged(a, b, steps_done, kw, guard, tol, theta, S) || breakdown || steps_done >= cap)) { converged = true; break; }
            if (!more) break;
            if (!ahead) { slot ^= 1; rc = issue_batch(std::min(8, max_iter - steps_issued), slot); if (rc) break; }
        }
    }
    for (int i = 0; i < 2; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    if (rc) return rc;
    const int steps = steps_done;
    if (iters_out) *iters_out = steps;
    if (getenv("ESPER_DEBUG_LANCZOS")) {
        int nre = 0;
        cudaStreamSynchronize(st);
        cudaMemcpy(&nre, stats_d, sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[lanczos] pro=%d resident=%d rows_cached=%d converged=%d steps=%d issued=%d explicit_reorth=%d\n", (int)use_pro, (int)resident, rows_cached, (int)converged, steps, steps_issued, nre);
        if (prof_d) {
            long long pr[8];
            cudaMemcpy(pr, prof_d, sizeof(pr), cudaMemcpyDeviceToHost);
            const double n = steps_issued > 0 ? (double)steps_issued : 1.0;
            fprintf(stderr, "[lanczos] CTA0 cycles/step: S0 %.0f | bar1 %.0f | S1 %.0f | bar2 %.0f | S2 %.0f | bar3 %.0f | norm+next %.0f ; per-launch prologue total %lld\n",
                    pr[0] / n, pr[1] / n, pr[2] / n, pr[3] / n, pr[4] / n, pr[5] / n, pr[6] / n, pr[7]);
        }
    }
    if (!converged) { cudaStreamSynchronize(st); return fail(c, ESPER_E_NOCONV, "Lanczos did not converge in %d steps", steps); }
    if (full) {
        // every column on request (k4_full.cu): QL on the host records the rotations, the device applies them to the rows of V^T
        ESPER_CUDA(c, cudaStreamSynchronize(st));
        double* Uf = nullptr;
        std::vector<double> th;
        int rcf = finish_full_spectrum(c, m, k, steps, a, b, V, &Uf, th);
        if (rcf) return rcf;
        { LaunchScope ls(c, "lanczos"); fix_columns_kernel<<<k, 256, 0, st>>>(Uf, m, k); }
        ESPER_CUDA(c, cudaGetLastError());
        if (vec_h) ESPER_CUDA(c, cudaMemcpyAsync(vec_h, Uf, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, st));
        ESPER_CUDA(c, cudaStreamSynchronize(st));
        if (val_h) { val_h[0] = 1.0; for (int i = 0; i + 1 < k; ++i) val_h[1 + i] = th[(size_t)i] * th[(size_t)i]; }
        return ESPER_OK;
    }

    // Ritz vectors U = [v0, V_1..V_steps S]
    ESPER_RESERVE(c, c->lan_small, sizeof(double) * ((size_t)m * k + (size_t)(max_iter + 8) * KMAX + 2 * KMAX * KMAX + 4096 + (size_t)ceil_div(m, 8) * 2));
    double* U = c->lan_small.as<double>();
    double* S_d = U + (size_t)m * k;
    double* G_d = S_d + (size_t)(max_iter + 8) * KMAX;          // k x k
    double* C_d = G_d + KMAX * KMAX;                            // k x k
    double* aux_d = C_d + KMAX * KMAX;                          // theta[2], cols[2] (as ints), residual partials
    std::vector<double> Sk;
    if (kw > 0) {
        const int want = (int)theta.size();
        if (want < kw) return fail(c, ESPER_E_NOCONV, "Krylov space exhausted: %d Ritz pairs < %d wanted", want, kw);
        Sk.resize((size_t)steps * kw);
        for (int l = 0; l < steps; ++l)
            for (int i = 0; i < kw; ++i) Sk[(size_t)l * kw + i] = S[(size_t)l * want + i];
        ESPER_CUDA(c, cudaMemcpyAsync(S_d, Sk.data(), sizeof(double) * Sk.size(), cudaMemcpyHostToDevice, st));
    }
    { LaunchScope ls(c, "lanczos"); ritz_kernel<<<dim3(ceil_div(m, 128), k), 128, 0, st>>>(V, m, kw > 0 ? steps : 0, S_d, kw, U, k); }
    { LaunchScope ls(c, "lanczos"); fix_columns_kernel<<<k, 256, 0, st>>>(U, m, k); }
    if (use_pro && kw > 0) {
        // Independent check of the first and the last wanted pair: true residual |Ms u - theta u|.
        const int nblk = ceil_div(m, 8);
        double th2[2] = {theta[0], theta[kw - 1]};
        int cols2[2] = {1, kw};
        ESPER_CUDA(c, cudaMemcpyAsync(aux_d, th2, sizeof(th2), cudaMemcpyHostToDevice, st));
        ESPER_CUDA(c, cudaMemcpyAsync(aux_d + 2, cols2, sizeof(cols2), cudaMemcpyHostToDevice, st));
        { LaunchScope ls(c, "lanczos"); residual_kernel<<<dim3(nblk, 2), 256, 0, st>>>(Ms, m, U, k, reinterpret_cast<int*>(aux_d + 2), aux_d, aux_d + 4); }
        std::vector<double> part((size_t)2 * nblk);
        ESPER_CUDA(c, cudaMemcpyAsync(part.data(), aux_d + 4, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, st));
        ESPER_CUDA(c, cudaStreamSynchronize(st));
        for (int q = 0; q < 2; ++q) {
            double r2 = 0.0;
            for (int i = 0; i < nblk; ++i) r2 += part[(size_t)q * nblk + i];
            if (getenv("ESPER_DEBUG_LANCZOS")) fprintf(stderr, "[lanczos] true residual of pair %d: %.3e\n", q == 0 ? 1 : kw, sqrt(r2));
            if (!(sqrt(r2) <= 1e3 * tol * std::max(fabs(theta[0]), 1e-300) + 1e-13)) *suspicious = true;
        }
    }
    ESPER_CUDA(c, cudaGetLastError());
    if (plain) {
        // general symmetric matrix with a known (locked) eigenvector: return the kw Ritz pairs only
        if (vec_h && kw > 0)
            ESPER_CUDA(c, cudaMemcpy2DAsync(vec_h, sizeof(double) * kw, U + 1, sizeof(double) * k, sizeof(double) * kw, m,
                                            cudaMemcpyDeviceToHost, st));
        ESPER_CUDA(c, cudaStreamSynchronize(st));
        if (val_h) for (int i = 0; i < kw; ++i) val_h[i] = theta[i];
        return ESPER_OK;
    }
    if (vec_h) ESPER_CUDA(c, cudaMemcpyAsync(vec_h, U, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, st));
    ESPER_CUDA(c, cudaStreamSynchronize(st));
    if (val_h) {
        val_h[0] = 1.0;
        for (int i = 0; i < kw; ++i) val_h[1 + i] = theta[i] * theta[i];
    }
    return ESPER_OK;
}

int run_lanczos(esper_ctx* c, int m, int k, double tol, int max_iter, double* val_h, double* vec_h, int* iters_out, bool plain) {
    using namespace lanczos;
    constexpr int KMAX = 64;
    // plain: c->markov holds any symmetric matrix and sqrt(c->degree) a known eigenvector that is locked out;
    // k counts the pairs wanted besides it (internally the locked vector is column 0, as in the diffusion-map case)
    if (plain) {
        if (k < 1 || k > m - 1 || k + 1 > KMAX) return fail(c, ESPER_E_ARG, "k=%d out of range (1..min(m-1,%d))", k, KMAX - 1);
        k += 1;
    }
    if (k < 1 || k > m) return fail(c, ESPER_E_ARG, "k=%d out of range (1..m = %d)", k, m);
    if (k > KMAX) {
        // more than KMAX columns (up to the reference's full U, k = m): Lanczos to exhaustion + QL on the tridiagonal matrix
        if (m < 3) return fail(c, ESPER_E_ARG, "k=%d: the full-spectrum path needs m >= 3", k);
        bool susp = false;
        const int mode_saved = c->lanczos_mode;
        if (c->lanczos_mode == 0) c->lanczos_mode = 3;               // barrier kernels (the chain kernel's basis capacity is ~440 vectors)
        const int rcf = lanczos_attempt(c, m, k, tol > 0.0 ? tol : ESPER_DEFAULT_EIG_TOL, m - 1, false, val_h, vec_h, iters_out, &susp, false, true);
        c->lanczos_mode = mode_saved;
        return rcf;
    }
    if (tol <= 0.0) tol = ESPER_DEFAULT_EIG_TOL;
    if (max_iter <= 0) max_iter = 1024;
    if (max_iter > m - 1) max_iter = m - 1;
    // Default: full reorthogonalisation (three grid barriers per step).  The partial-reorthogonalisation kernel
    // (one barrier per step) is opt-in (ESPER_LANCZOS_PRO=1): measured at m = 2000 it needs an explicit pass on
    // 30 % of the steps and leaves true residuals of ~3e-9 (delta * beta * sqrt(j)) instead of 1e-12, for a
    // 20 % shorter solve -- not worth the accuracy.  A PRO run whose independent checks fail is repeated.
    bool use_pro = getenv("ESPER_LANCZOS_PRO") && (m <= PRO_MAX_M) && (max_iter + 4 <= 2048);
    bool suspicious = false;
    if (plain) use_pro = false;
    int rc = lanczos_attempt(c, m, k, tol, max_iter, use_pro, val_h, vec_h, iters_out, &suspicious, plain);
    if (use_pro && (suspicious || rc == ESPER_E_NOCONV)) {
        c->lanczos_fallbacks++;
        rc = lanczos_attempt(c, m, k, tol, max_iter, false, val_h, vec_h, iters_out, &suspicious, plain);
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Row-block (oversized PD) building blocks: the host layer owns the vectors (device pointers, e.g. torch
// tensors that NCCL all-gathers) and drives the iteration; every rank orthogonalises the same replicated
// vector with the same deterministic kernels, so the only collective per step is the all-gather of w.
// ------------------------------------------------------------------------------------------------
namespace lanczos {

// w[row] = Ms_rows[row, :] . q   (Ms block streamed from HBM; one warp per row, 8 x 16 B in flight per lane)
__global__ void __launch_bounds__(256) symv_rows_kernel(const double* __restrict__ Ms, int rows, int m,
                                                        const double* __restrict__ q, double* __restrict__ w) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* a = Ms + (size_t)row * m;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int cidx = lane;
    for (; cidx + 224 < m; cidx += 256) {
        const double x0 = __ldcs(a + cidx), x1 = __ldcs(a + cidx + 32), x2 = __ldcs(a + cidx + 64), x3 = __ldcs(a + cidx + 96),
                     x4 = __ldcs(a + cidx + 128), x5 = __ldcs(a + cidx + 160), x6 = __ldcs(a + cidx + 192), x7 = __ldcs(a + cidx + 224);
        a0 = fma(x0, __ldg(q + cidx), a0);       a1 = fma(x1, __ldg(q + cidx + 32), a1);
        a2 = fma(x2, __ldg(q + cidx + 64), a2);  a3 = fma(x3, __ldg(q + cidx + 96), a3);
        a0 = fma(x4, __ldg(q + cidx + 128), a0); a1 = fma(x5, __ldg(q + cidx + 160), a1);
        a2 = fma(x6, __ldg(q + cidx + 192), a2); a3 = fma(x7, __ldg(q + cidx + 224), a3);
    }
    for (; cidx < m; cidx += 32) a0 = fma(__ldcs(a + cidx), __ldg(q + cidx), a0);
    const double tot = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) w[row] = tot;
}

// ---- fused SYMV + all-gather over NVLink peer memory (row-block path, one process per GPU) -------------------------
// Every rank owns a communication buffer that its peers map through CUDA IPC:
//     ll[2][n_pad] 16-byte packets (two slots, step parity) | w[n_pad] doubles (the unpacked vector) | error u32
// symv_allgather_kernel computes the rank's rows of w = Ms q exactly like symv_rows_kernel and writes them in wire
// format, CTA by CTA as the rows finish, into slot (step & 1) of its OWN buffer: each double is one 16-byte packet
// {lo, step, hi, step} whose 8-byte halves carry their own flag (the low-latency protocol NCCL calls LL) -- no staging
// copy, no NCCL call, no memory fence, no remote store.  ll_gather_kernel on every rank (stream-ordered before the
// orthogonalisation) PULLS all n packets, its own from local memory and the others over NVLink, spinning per element
// until both flags show the step, so the transfer of the finished rows overlaps the rest of the peers' SYMV; a watchdog
// raises the error word instead of hanging.  Slot reuse two steps later is safe: a rank overwrites its slot for step
// j+2 only after its gather of step j+1 completed, i.e. after every peer produced step j+1, which each peer does
// (stream order) after its gather of step j.
// Three push-based versions were measured first (profiles/r1_fused_allgather.md): (1) every CTA stores its rows to all
// peers + a system-scope fence per CTA: SYMV bandwidth halved; (2) the last CTA pushes the block, one fence.sys: 6 %
// faster than NCCL at 2 GPUs, 11 % slower at 8 (a fence.sys behind stores to 7 peers costs tens of microseconds; with
// 16 segment pushes, 80 % slower); (3) LL packets stored to all peers by every CTA, no fence: bit-correct but 2.5x
// slower at N = 50 000 -- remote stores issued from all SMs while they stream Ms from HBM stall the SYMV.
constexpr int MAX_PEERS = 16;
struct PeerTable { uint4* ll[MAX_PEERS]; };

__device__ __forceinline__ void st_ll(uint4* p, double v, unsigned int flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)__double2loint(v)), "r"(flag),
                 "r"((unsigned int)__double2hiint(v)), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
symv_allgather_kernel(const double* __restrict__ Ms, int rows, int m, const double* __restrict__ q, uint4* __restrict__ mine_ll,
                      int n_pad, int row0, unsigned int step) {
    __shared__ double res[8];
    const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + wrp;
    if (row < rows) {
        const double* a = Ms + (size_t)row * m;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int cidx = lane;
        for (; cidx + 224 < m; cidx += 256) {
            const double x0 = __ldcs(a + cidx), x1 = __ldcs(a + cidx + 32), x2 = __ldcs(a + cidx + 64), x3 = __ldcs(a + cidx + 96),
                         x4 = __ldcs(a + cidx + 128), x5 = __ldcs(a + cidx + 160), x6 = __ldcs(a + cidx + 192), x7 = __ldcs(a + cidx + 224);
            a0 = fma(x0, __ldg(q + cidx), a0);       a1 = fma(x1, __ldg(q + cidx + 32), a1);
            a2 = fma(x2, __ldg(q + cidx + 64), a2);  a3 = fma(x3, __ldg(q + cidx + 96), a3);
            a0 = fma(x4, __ldg(q + cidx + 128), a0); a1 = fma(x5, __ldg(q + cidx + 160), a1);
            a2 = fma(x6, __ldg(q + cidx + 192), a2); a3 = fma(x7, __ldg(q + cidx + 224), a3);
        }
        for (; cidx < m; cidx += 32) a0 = fma(__ldcs(a + cidx), __ldg(q + cidx), a0);
        const double tot = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) res[wrp] = tot;
    }
    __syncthreads();
    // this CTA's 8 rows as 8 packets (128 contiguous bytes) in the rank's OWN buffer; the peers pull them over NVLink
    if (threadIdx.x < 8 && blockIdx.x * 8 + threadIdx.x < rows)
        st_ll(mine_ll + (size_t)(step & 1u) * n_pad + row0 + blockIdx.x * 8 + threadIdx.x, res[threadIdx.x], step);
}

// Receiver: element i lives in the buffer of rank i / rows_per (local memory for this rank's own rows, NVLink peer memory
// otherwise); w[i] = its double once both flags of the packet show `step`.  On timeout the error word is raised.
__global__ void __launch_bounds__(256)
ll_gather_kernel(PeerTable peers, int rows_per, int n_pad, int n, unsigned int step, long long timeout_cycles,
                 double* __restrict__ w, unsigned int* __restrict__ error) {
    const long long t0 = clock64();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4* src = peers.ll[i / rows_per] + (size_t)(step & 1u) * n_pad + i;
        uint4 v = ld_ll(src);
        while (v.y != step || v.w != step) {
            if (clock64() - t0 > timeout_cycles) { atomicExch(error, 0x100u); return; }
            __nanosleep(200);
            v = ld_ll(src);
        }
        w[i] = __hiloint2double((int)v.z, (int)v.x);
    }
}

}  // namespace lanczos

namespace {
inline size_t comm_ll_bytes(int n_pad) { return sizeof(uint4) * 2 * (size_t)n_pad; }
inline size_t comm_w_bytes(int n_pad) { return sizeof(double) * (size_t)n_pad; }
}

// Communication buffer of this rank (see above) and its IPC handle.
int run_comm_alloc(esper_ctx* c, int n_pad, int world, void* handle_out) {
    using namespace lanczos;
    if (n_pad < 1 || world < 1 || world > MAX_PEERS) return fail(c, ESPER_E_ARG, "comm_alloc: need n_pad >= 1 and 1 <= world <= %d", MAX_PEERS);
    const size_t bytes = comm_ll_bytes(n_pad) + comm_w_bytes(n_pad) + 128;
    c->comm_buf.release();                                   // a fresh allocation: its handle is exported below
    ESPER_RESERVE(c, c->comm_buf, bytes);
    ESPER_CUDA(c, cudaMemsetAsync(c->comm_buf.p, 0, bytes, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    c->comm_n_pad = n_pad; c->comm_world = world; c->comm_open = false;
    cudaIpcMemHandle_t h;
    ESPER_CUDA(c, cudaIpcGetMemHandle(&h, c->comm_buf.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, sizeof h);
    return ESPER_OK;
}

int run_comm_open(esper_ctx* c, const void* handles, int rank, int world) {
    using namespace lanczos;
    if (!c->comm_buf.p || world != c->comm_world || rank < 0 || rank >= world)
        return fail(c, ESPER_E_STATE, "comm_open: call comm_alloc with the same world size first");
    for (int p = 0; p < world; ++p) {
        void* base = nullptr;
        if (p == rank) base = c->comm_buf.p;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles + 64 * (size_t)p, sizeof h);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int q = 0; q < p; ++q) if (q != rank && c->comm_peer[q]) { cudaIpcCloseMemHandle(c->comm_peer[q]); c->comm_peer[q] = nullptr; }
                return fail(c, ESPER_E_CUDA, "comm_open: cudaIpcOpenMemHandle for rank %d: %s", p, cudaGetErrorString(e));
            }
        }
        c->comm_peer[p] = base;
    }
    c->comm_rank = rank; c->comm_open = true;
    return ESPER_OK;
}

int run_comm_close(esper_ctx* c) {
    if (c->comm_open)
        for (int p = 0; p < c->comm_world; ++p)
            if (p != c->comm_rank && c->comm_peer[p]) cudaIpcCloseMemHandle(c->comm_peer[p]);
    for (auto& p : c->comm_peer) p = nullptr;
    c->comm_open = false;
    return ESPER_OK;
}

int run_symv_allgather(esper_ctx* c, int rows, int m, int row0, const double* q_d, int step) {
    using namespace lanczos;
    if (!c->comm_open) return fail(c, ESPER_E_STATE, "symv_allgather_d: no open communicator (comm_alloc + comm_open)");
    if (step < 1 || row0 < 0 || row0 + rows > c->comm_n_pad) return fail(c, ESPER_E_ARG, "symv_allgather_d: step >= 1 and row0 + rows <= n_pad required");
    if (row0 != c->comm_rank * (c->comm_n_pad / c->comm_world))
        return fail(c, ESPER_E_ARG, "symv_allgather_d: row0 must be rank * rows_per_rank (%d)", c->comm_rank * (c->comm_n_pad / c->comm_world));
    {
        LaunchScope ls(c, "lanczos");
        symv_allgather_kernel<<<ceil_div(rows, 8), 256, 0, c->stream>>>(c->markov.as<double>(), rows, m, q_d, c->comm_buf.as<uint4>(),
                                                                        c->comm_n_pad, row0, (unsigned int)step);
    }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

// Enqueue the unpack of step `step` (waits per element for its flags); *w_full_d = the gathered dense vector on this rank.
int run_comm_wait(esper_ctx* c, int step, int n, double** w_full_d) {
    using namespace lanczos;
    if (!c->comm_open) return fail(c, ESPER_E_STATE, "comm_wait_d: no open communicator");
    if (n < 1 || n > c->comm_n_pad) return fail(c, ESPER_E_ARG, "comm_wait_d: need 1 <= n <= n_pad");
    PeerTable pt;
    for (int p = 0; p < MAX_PEERS; ++p) pt.ll[p] = reinterpret_cast<uint4*>(c->comm_peer[p < c->comm_world ? p : c->comm_rank]);
    double* w = reinterpret_cast<double*>(c->comm_buf.as<char>() + comm_ll_bytes(c->comm_n_pad));
    unsigned int* error = reinterpret_cast<unsigned int*>(c->comm_buf.as<char>() + comm_ll_bytes(c->comm_n_pad) + comm_w_bytes(c->comm_n_pad));
    {
        LaunchScope ls(c, "lanczos");
        const int grid = std::max(1, std::min(ceil_div(n, 256), c->sm_count));
        ll_gather_kernel<<<grid, 256, 0, c->stream>>>(pt, c->comm_n_pad / c->comm_world, c->comm_n_pad, n, (unsigned int)step,
                                                      20000000000LL, w, error);                                  // ~10 s watchdog
    }
    ESPER_CUDA(c, cudaGetLastError());
    if (w_full_d) *w_full_d = w;
    return ESPER_OK;
}

// 0 = healthy; non-zero = an unpack timed out (synchronises the stream).
int run_comm_status(esper_ctx* c, unsigned int* status) {
    using namespace lanczos;
    if (!c->comm_buf.p) return fail(c, ESPER_E_STATE, "comm_status: no communicator");
    const unsigned int* error = reinterpret_cast<unsigned int*>(c->comm_buf.as<char>() + comm_ll_bytes(c->comm_n_pad) + comm_w_bytes(c->comm_n_pad));
    ESPER_CUDA(c, cudaMemcpyAsync(status, error, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

int run_symv_rows(esper_ctx* c, int rows, int m, const double* q_d, double* w_rows_d) {
    using namespace lanczos;
    { LaunchScope ls(c, "lanczos"); symv_rows_kernel<<<ceil_div(rows, 8), 256, 0, c->stream>>>(c->markov.as<double>(), rows, m, q_d, w_rows_d); }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

// V_d[0] = v0 = sqrt(d)/|sqrt(d)|, V_d[1] = deterministic start vector orthogonal to v0 (unit norm).
int run_lanczos_start(esper_ctx* c, int m, const double* degree_all_h, double* V_d) {
    using namespace lanczos;
    cudaStream_t st = c->stream;
    ESPER_RESERVE(c, c->degree, sizeof(double) * (size_t)m);
    ESPER_RESERVE(c, c->lan_w, sizeof(double) * (size_t)m * 2);
    ESPER_RESERVE(c, c->lan_h, sizeof(double) * 4096);
    ESPER_CUDA(c, cudaMemcpyAsync(c->degree.p, degree_all_h, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, st));
    double* w = c->lan_w.as<double>(); double* w2 = w + m; double* h = c->lan_h.as<double>();
    { LaunchScope ls(c, "lanczos"); init_kernel<<<ceil_div(m, 256), 256, 0, st>>>(c->degree.as<double>(), m, w, w2); }
    { LaunchScope ls(c, "lanczos"); normalize_kernel<<<1, 1024, 0, st>>>(w, m, V_d, nullptr); }
    for (int pass = 0; pass < 2; ++pass) {
        { LaunchScope ls(c, "lanczos"); dots_kernel<<<1, 256, 0, st>>>(V_d, m, w2, h); }
        { LaunchScope ls(c, "lanczos"); update_kernel<<<ceil_div(m, 256), 256, 0, st>>>(V_d, m, 1, h, w2, nullptr); }
    }
    { LaunchScope ls(c, "lanczos"); normalize_kernel<<<1, 1024, 0, st>>>(w2, m, V_d + m, nullptr); }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

// One Lanczos orthogonalisation step on the replicated vectors: w = Ms q_j (gathered, length m) is
// orthogonalised against V[0..j] by classical Gram-Schmidt applied twice, ab_d[0] = alpha_j, ab_d[1] = beta_j,
// V[j+1] = w / beta_j.
int run_lanczos_ortho(esper_ctx* c, int m, int j, double* V_d, double* w_d, double* ab_d) {
    using namespace lanczos;
    cudaStream_t st = c->stream;
    if (j + 1 > 5000) return fail(c, ESPER_E_ARG, "lanczos_ortho_d: at most 5000 basis vectors");
    ESPER_RESERVE(c, c->lan_h, sizeof(double) * round_up((size_t)(j + 8) * DOT_SPLIT, 16384));   // grows in 128 kB steps
    double* h = c->lan_h.as<double>();
    ESPER_CUDA(c, cudaMemsetAsync(ab_d, 0, 2 * sizeof(double), st));
    for (int pass = 0; pass < 2; ++pass) {
        { LaunchScope ls(c, "lanczos"); dots_split_kernel<<<dim3(j + 1, DOT_SPLIT), 256, 0, st>>>(V_d, m, w_d, h); }
        { LaunchScope ls(c, "lanczos"); update_split_kernel<<<ceil_div(m, 256), 256, sizeof(double) * (size_t)(j + 1), st>>>(V_d, m, j + 1, h, w_d, ab_d); }
    }
    { LaunchScope ls(c, "lanczos"); normalize_kernel<<<1, 1024, 0, st>>>(w_d, m, V_d + (size_t)(j + 1) * m, ab_d + 1); }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

// U = [v0, V_1..V_steps S] -> vec_h [m, k] (unit columns, sign fixed).  S_h: [steps, k-1] row-major.
int run_ritz_vectors(esper_ctx* c, int m, int steps, int k, const double* V_d, const double* S_h, double* vec_h) {
    using namespace lanczos;
    cudaStream_t st = c->stream;
    const int kw = k - 1;
    ESPER_RESERVE(c, c->lan_small, sizeof(double) * ((size_t)m * k + (size_t)steps * (kw > 0 ? kw : 1) + 8));
    double* U = c->lan_small.as<double>();
    double* S_d = U + (size_t)m * k;
    if (kw > 0) ESPER_CUDA(c, cudaMemcpyAsync(S_d, S_h, sizeof(double) * (size_t)steps * kw, cudaMemcpyHostToDevice, st));
    { LaunchScope ls(c, "lanczos"); ritz_kernel<<<dim3(ceil_div(m, 128), k), 128, 0, st>>>(V_d, m, kw > 0 ? steps : 0, S_d, kw, U, k); }
    { LaunchScope ls(c, "lanczos"); fix_columns_kernel<<<k, 256, 0, st>>>(U, m, k); }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(vec_h, U, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, st));
    ESPER_CUDA(c, cudaStreamSynchronize(st));
    return ESPER_OK;
}

}  // namespace esper
