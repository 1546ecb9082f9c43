// k2_sweep.cu -- K2: the epsilon sweep of 1_Embedding/DM/GaussianBandwidth.py:36-44,
//     logSum[k] = log sum_{ij : t < thr} exp(-t),   t = coef[k] * D_ij^2,   k = 0 .. n_eps-1.
//
// The reference evaluates all m^2 * n_eps terms.  For a block of elements the sweep value is known
// in closed form for most k: if coef[k]*min_nonzero(D^2) >= thr every non-zero element is truncated
// (only exact zeros, e.g. the diagonal, contribute 1), and if coef[k]*max(D^2) < 2^-60 every term is
// exp(-t) == 1.0 in binary64.  Only the k in between (~220 of 1001 for a span of e^0.2 per step) need
// exp().  The inclusion test uses the same products t = coef*D2 as the reference, so it is bit-exact.
//
// One pass over D (coalesced, 8 elements per thread), per-k sums accumulated in shared memory per
// persistent block, fixed-order FP64 reductions (deterministic).  FP64-ALU bound on the active k.
#include "common.cuh"
#include "devutil.cuh"

namespace esper {
namespace sweep {

constexpr int THREADS = 256;
constexpr int EPT = 4;                       // elements per thread per tile
constexpr int KB = 8;                        // k values reduced per barrier pair
constexpr int WARPS = THREADS / 32;

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exp(-t) for 0 <= t: argument reduction by ln2 (hi/lo split), degree-13 Taylor polynomial on
// |f| <= ln2/2 (truncation error 4e-18), exponent patched in place.  <= 1 ulp.  Callers route t > 700
// (only reachable with thr = +inf) to the library exp, which handles gradual underflow.
__device__ __forceinline__ double exp_neg_fast(double t) {     // requires 0 <= t <= 700
    const double y = -t;
    const double n = rint(y * 1.4426950408889634074);
    double f = fma(n, -6.93147180369123816490e-01, y);
    f = fma(n, -1.90821492927058770002e-10, f);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, f, 2.08767569878681e-09);           // 1/12!
    p = fma(p, f, 2.505210838544172e-08);          // 1/11!
    p = fma(p, f, 2.755731922398589e-07);          // 1/10!
    p = fma(p, f, 2.7557319223985893e-06);         // 1/9!
    p = fma(p, f, 2.48015873015873e-05);           // 1/8!
    p = fma(p, f, 1.984126984126984e-04);          // 1/7!
    p = fma(p, f, 1.388888888888889e-03);          // 1/6!
    p = fma(p, f, 8.333333333333333e-03);          // 1/5!
    p = fma(p, f, 4.1666666666666664e-02);         // 1/4!
    p = fma(p, f, 1.6666666666666666e-01);         // 1/3!
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    return __longlong_as_double(__double_as_longlong(p) + ((long long)(int)n << 52));
}

constexpr int TR = 8;                        // tile rows (THREADS / TC row groups x EPT)
constexpr int TC = 128;                      // tile columns (a warp reads 32 consecutive doubles of a row)

// sym != 0: D is exactly symmetric; only elements j >= i are visited, off-diagonal ones weigh 2.
__global__ void __launch_bounds__(THREADS)
sweep_kernel(const double* __restrict__ D, int rows, int m, int sym, int mono, const double* __restrict__ coef, int n_eps,
             double thr, const int* __restrict__ row_start /*[tiles_r+1] prefix of valid tiles*/,
             unsigned int* __restrict__ next_tile, double* __restrict__ part /*[gridDim.x][n_eps]*/) {
    extern __shared__ double sm[];
    double* acc = sm;                         // [n_eps]
    double* cf = sm + n_eps;                  // [n_eps]
    double* red = cf + n_eps;                 // [WARPS][KB]
    double* stat = red + WARPS * KB;          // [4][WARPS] min, max, n_zero, n_valid
    double* below = stat + 4 * WARPS;         // [n_eps] difference array: closed-form A contributions
    double* above = below + n_eps;            // [n_eps] difference array: closed-form B contributions
    int* s_act = reinterpret_cast<int*>(above + n_eps);      // [n_eps] active k list (non-monotone coefficients)
    int* s_nact = s_act + n_eps;
    int* s_rng = s_nact + 1;                  // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double TINY = 8.673617379884035e-19;   // 2^-60: exp(-t) == 1.0 for 0 <= t < TINY
    const double SMALL = 7.450580596923828e-09;  // 2^-27: exp(-t) = 1 - t + t^2/2 to < 1e-25 relative

    for (int k = threadIdx.x; k < n_eps; k += THREADS) { acc[k] = 0.0; cf[k] = coef[k]; below[k] = 0.0; above[k] = 0.0; }
    __syncthreads();

    const int tiles_c = (m + TC - 1) / TC, tiles_r = (rows + TR - 1) / TR;
    const int n_valid = row_start[tiles_r];
    __shared__ int s_tile;
    for (;;) {
        // dynamic tile scheduler: valid tiles (on or above the diagonal when sym) are handed out in order
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(next_tile, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_valid) break;
        int lo_r = 0, hi_r = tiles_r;                             // tr = last row with row_start[tr] <= tile
        while (hi_r - lo_r > 1) { const int mid = (lo_r + hi_r) >> 1; if (row_start[mid] <= tile) lo_r = mid; else hi_r = mid; }
        const int tr = lo_r;
        const int tc = tiles_c - (row_start[tr + 1] - row_start[tr]) + (tile - row_start[tr]);
        const int row0 = tr * TR, col0 = tc * TC;
        double d2[EPT], wt[EPT];
        double mn = INFINITY, mx = 0.0, nz = 0.0, nv = 0.0;
        const int j = col0 + (threadIdx.x & (TC - 1));
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = row0 + (threadIdx.x >> 7) * EPT + e;
            d2[e] = 0.0;                     // excluded (below the diagonal, out of range, NaN): weight 0
            wt[e] = 0.0;
            if (i < rows && j < m && (!sym || j >= i)) {
                const double d = __ldg(D + (size_t)i * m + j);
                const double q = d * d;
                if (q == q) {
                    const double w = (sym && j > i) ? 2.0 : 1.0;
                    d2[e] = q; wt[e] = w;
                    nv += w;
                    if (q == 0.0) nz += w; else mn = fmin(mn, q);
                    mx = fmax(mx, q);
                }
            }
        }
        mn = warp_min(mn); mx = warp_max(mx); nz = warp_sum(nz); nv = warp_sum(nv);
        __syncthreads();
        if (lane == 0) { stat[warp] = mn; stat[WARPS + warp] = mx; stat[2 * WARPS + warp] = nz; stat[3 * WARPS + warp] = nv; }
        __syncthreads();
        mn = stat[0]; mx = stat[WARPS]; nz = 0.0; nv = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            mn = fmin(mn, stat[w]); mx = fmax(mx, stat[WARPS + w]);
            nz += stat[2 * WARPS + w]; nv += stat[3 * WARPS + w];
        }
        const double zero_term = (0.0 < thr) ? nz : 0.0;

        // Which k need evaluation for this tile?  (block-uniform facts)
        //   A: no non-zero element passes the truncation  -> contribution zero_term (exact zeros give 1)
        //   B: every element passes and exp(-t) == 1.0     -> contribution nv
        // With non-increasing coefficients (the reference's grid) A holds for k < ka and B for k >= kb, so the
        // closed forms go into two difference arrays (O(1) per tile) and [ka, kb) is evaluated; otherwise
        // every k is classified and the active ones are queued.
        int nact, ka = 0;
        if (mono) {
            if (threadIdx.x == 0) {
                int lo_k = 0, hi_k = n_eps;                        // ka = first k with cf[k]*mn < thr
                if (mn == INFINITY) lo_k = n_eps;
                else while (lo_k < hi_k) { const int mid = (lo_k + hi_k) >> 1; if (cf[mid] * mn < thr) hi_k = mid; else lo_k = mid + 1; }
                const int ka_ = lo_k;
                lo_k = 0; hi_k = n_eps;                            // kb = first k with cf[k]*mx < min(TINY, thr)
                while (lo_k < hi_k) { const int mid = (lo_k + hi_k) >> 1; if (cf[mid] * mx < TINY && cf[mid] * mx < thr) hi_k = mid; else lo_k = mid + 1; }
                int kb_ = lo_k < ka_ ? ka_ : lo_k;
                if (ka_ > 0) below[ka_ - 1] += zero_term;          // applies to all k <= ka-1 (suffix sum at the end)
                if (kb_ < n_eps) above[kb_] += nv;                 // applies to all k >= kb   (prefix sum at the end)
                s_rng[0] = ka_; s_rng[1] = kb_;
            }
            __syncthreads();
            ka = s_rng[0];
            nact = s_rng[1] - ka;
        } else {
            if (threadIdx.x == 0) *s_nact = 0;
            __syncthreads();
            for (int k = threadIdx.x; k < n_eps; k += THREADS) {
                const double c = cf[k];
                const bool shortcut = (c >= 0.0) && (c < INFINITY);      // monotone t = c*d2, no NaN products
                if (shortcut && (mn == INFINITY || !(c * mn < thr))) acc[k] += zero_term;
                else if (shortcut && c * mx < TINY && c * mx < thr) acc[k] += nv;
                else s_act[atomicAdd(s_nact, 1)] = k;
            }
            __syncthreads();
            nact = *s_nact;
        }

        for (int a0 = 0; a0 < nact; a0 += KB) {
#pragma unroll
            for (int jj = 0; jj < KB; ++jj) {
                double s = 0.0;
                if (a0 + jj < nact) {
                    const double c = cf[mono ? ka + a0 + jj : s_act[a0 + jj]];
                    const double tmax = c * mx;
                    if ((c >= 0.0) && tmax < SMALL && tmax < thr) {
                        // every element passes; the three-term series is exact to working precision
#pragma unroll
                        for (int e = 0; e < EPT; ++e) {
                            const double t = c * d2[e];
                            s = fma(wt[e], fma(t, fma(t, 0.5, -1.0), 1.0), s);
                        }
                    } else if ((c >= 0.0) && tmax < 700.0) {
                        // branch-free: all EPT exponentials are independent instruction streams
                        double ev[EPT];
#pragma unroll
                        for (int e = 0; e < EPT; ++e) ev[e] = exp_neg_fast(fmin(c * d2[e], 700.0));
#pragma unroll
                        for (int e = 0; e < EPT; ++e) {
                            const double t = c * d2[e];
                            const double term = (t < SMALL) ? fma(t, fma(t, 0.5, -1.0), 1.0) : ev[e];
                            s += (t < thr) ? wt[e] * term : 0.0;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) {
                            const double t = c * d2[e];
                            if (t < thr) s = fma(wt[e], (t < SMALL) ? fma(t, fma(t, 0.5, -1.0), 1.0) : exp(-t), s);
                        }
                    }
                    s = warp_sum(s);
                }
                if (lane == 0) red[warp * KB + jj] = s;
            }
            __syncthreads();
            if (threadIdx.x < KB && a0 + threadIdx.x < nact) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) t += red[w * KB + threadIdx.x];
                acc[mono ? ka + a0 + threadIdx.x : s_act[a0 + threadIdx.x]] += t;
            }
            __syncthreads();
        }
    }
    if (mono) {                                   // fold the difference arrays into acc (one thread; n_eps is small)
        __syncthreads();
        if (threadIdx.x == 0) {
            double run = 0.0;
            for (int k = n_eps - 1; k >= 0; --k) { run += below[k]; acc[k] += run; }
            run = 0.0;
            for (int k = 0; k < n_eps; ++k) { run += above[k]; acc[k] += run; }
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < n_eps; k += THREADS) part[(size_t)blockIdx.x * n_eps + k] = acc[k];
}

// flag = 0 if D[i][j] != D[j][i] anywhere (bitwise comparison of the values).
__global__ void __launch_bounds__(256) symmetry_kernel(const double* __restrict__ D, int m, int* __restrict__ flag) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * m) return;
    const int i = (int)(idx / m), j = (int)(idx % m);
    if (j > i) {
        const double a = D[idx], b = D[(size_t)j * m + i];
        if (!(a == b) && !(a != a && b != b)) *flag = 0;
    }
}

__global__ void __launch_bounds__(256)
sweep_final_kernel(const double* __restrict__ part, int n_blocks, int n_eps, int take_log, double* __restrict__ out) {
    __shared__ double sh[32];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) s += part[(size_t)b * n_eps + k];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[k] = take_log ? log(s) : s;
}

}  // namespace sweep

int check_symmetry(esper_ctx* c, int m) {
    ESPER_RESERVE(c, c->sweep_aux, sizeof(double) * 4096);
    int* flag = reinterpret_cast<int*>(c->sweep_aux.as<double>() + 4000);
    int one = 1;
    ESPER_CUDA(c, cudaMemcpyAsync(flag, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, "sweep");
        sweep::symmetry_kernel<<<ceil_div((long long)m * m, 256), 256, 0, c->stream>>>(c->dist.as<double>(), m, flag);
    }
    int res = 0;
    ESPER_CUDA(c, cudaMemcpyAsync(&res, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    c->dist_symmetric = res;
    return ESPER_OK;
}

int run_sweep(esper_ctx* c, int m, const double* coef_h, int n_eps, double thr, double* logsum_h) {
    return run_sweep_rows(c, m, m, coef_h, n_eps, thr, logsum_h, 1);
}

// rows x m block of D resident in c->dist; take_log = 0 returns the partial sums (row-block path).
int run_sweep_rows(esper_ctx* c, int rows, int m, const double* coef_h, int n_eps, double thr, double* logsum_h, int take_log) {
    using namespace sweep;
    int sym = 0;
    if (rows == m && c->dist_row0 < 0) {
        if (c->dist_symmetric < 0) { int rc = check_symmetry(c, m); if (rc) return rc; }
        sym = c->dist_symmetric > 0 ? 1 : 0;
    }
    const int tiles_r = ceil_div(rows, TR), tiles_c = ceil_div(m, TC);
    std::vector<int> row_start(tiles_r + 2, 0);                   // prefix count of valid tiles per tile row
    for (int tr = 0; tr < tiles_r; ++tr) {
        const int first = sym ? (tr * TR) / TC : 0;               // first tile column touching j >= i
        row_start[tr + 1] = row_start[tr] + (tiles_c - first);
    }
    const int n_valid = row_start[tiles_r];
    int grid = c->sm_count * 4;
    if (grid > n_valid) grid = n_valid;
    if (grid < 1) grid = 1;
    ESPER_RESERVE(c, c->sweep_part, sizeof(double) * (size_t)grid * n_eps);
    ESPER_RESERVE(c, c->sweep_aux, sizeof(double) * ((size_t)n_eps * 2 + 4096) + sizeof(int) * (size_t)(tiles_r + 8));
    double* coef_d = c->sweep_aux.as<double>() + 4096;
    double* logsum_d = coef_d + n_eps;
    int* row_start_d = reinterpret_cast<int*>(logsum_d + n_eps);
    unsigned int* next_tile_d = reinterpret_cast<unsigned int*>(row_start_d + tiles_r + 2);
    ESPER_CUDA(c, cudaMemcpyAsync(coef_d, coef_h, sizeof(double) * n_eps, cudaMemcpyHostToDevice, c->stream));
    ESPER_CUDA(c, cudaMemcpyAsync(row_start_d, row_start.data(), sizeof(int) * (size_t)(tiles_r + 1), cudaMemcpyHostToDevice, c->stream));
    ESPER_CUDA(c, cudaMemsetAsync(next_tile_d, 0, sizeof(unsigned int), c->stream));
    const size_t smem = sizeof(double) * ((size_t)4 * n_eps + WARPS * KB + 4 * WARPS) + sizeof(int) * ((size_t)n_eps + 4);
    int mono = 1;                                  // non-increasing, finite, non-negative coefficients?
    for (int k = 0; k < n_eps; ++k)
        if (!(coef_h[k] >= 0.0) || !(coef_h[k] < INFINITY) || (k > 0 && coef_h[k] > coef_h[k - 1])) { mono = 0; break; }
    if (smem > 200 * 1024) return fail(c, ESPER_E_ARG, "n_eps=%d too large for the sweep kernel", n_eps);
    if (!c->attr_sweep) {
        ESPER_CUDA(c, cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        c->attr_sweep = true;
    }
    {
        LaunchScope ls(c, "sweep");
        sweep_kernel<<<grid, THREADS, smem, c->stream>>>(c->dist.as<double>(), rows, m, sym, mono, coef_d, n_eps, thr,
                                                         row_start_d, next_tile_d, c->sweep_part.as<double>());
    }
    {
        LaunchScope ls(c, "sweep");
        sweep_final_kernel<<<n_eps, 256, 0, c->stream>>>(c->sweep_part.as<double>(), grid, n_eps, take_log, logsum_d);
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(logsum_h, logsum_d, sizeof(double) * n_eps, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

}  // namespace esper
