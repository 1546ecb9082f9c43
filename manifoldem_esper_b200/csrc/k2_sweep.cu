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


// ---------------------------------------------------------------------------------------------------------------------
// Moment sweep (opt-in, ESPER_SWEEP_MOMENTS=1): O(1) work per element instead of one exp() per (element, active k).
//
//   S_k = sum_{ij : c_k q_ij < thr} w_ij exp(-c_k q_ij),   q = D^2.
//
// The non-zero q are grouped into <= MB geometric bins [e_b, e_b rho), e_0 = min q, rho = 1/(1 - 2/thr) (1.25 for the
// reference's thr = 10).  For one bin and one k exactly one of three things holds (c_k >= 0 non-increasing in k, so
// each is a contiguous k range found by bisection; fl(c q) is monotone in q):
//   k <  ka_b : c_k e_b >= thr        -> no element of the bin passes the truncation
//   k >= kf_b : c_k e_{b+1} < thr     -> every element passes, and |c_k (q - mu_b)| <= thr (rho - 1)/(2 rho) = 1, so
//                                        exp(-c q) = exp(-c mu_b) sum_p (-c h_b)^p u^p / p!,  u = (q - mu_b)/h_b in [-1, 1],
//                                        truncated after p = MP = 16 (remainder <= e/17! = 7.6e-15 relative; all terms of
//                                        S_k are positive, so that is the relative error of the bin's sum): only the
//                                        moments M_p = sum w u^p are needed, accumulated in registers
//   otherwise : the truncation boundary thr/c_k falls inside the bin (at most 2 k for the reference's grid of e^0.2 per
//               step; more than MS -> the general kernel runs instead): evaluated per element with the same products
//               t = c_k q and the same exp as sweep_kernel, so the inclusion test stays bit-identical.
// Exact zeros (the diagonal, duplicate images) contribute 1 each for every k.  Launch chain: statistics -> plan (one
// block) -> moments (one pass over D per bin; one bin on high-noise stacks, where D^2 spans a few per cent) -> final.
// Every reduction has a fixed order (deterministic).  Numpy statement + error measurements: tools/sweep_moments_proto.py.
constexpr int MP = 16;                       // highest moment
constexpr int MS = 4;                        // straddling k per bin evaluated exactly
constexpr int MB = 32;                       // bins
constexpr int MV = MP + 1 + MS;              // values per (block, bin)
constexpr int PLAN_HDR = 8, PLAN_BIN = 6;    // plan layout (doubles): [status, B, zero_term, n_valid, mn, mx, rho, -] + B x [lo, hi, mu, 1/h, ka, kf]
constexpr int PLAN_LEN = PLAN_HDR + PLAN_BIN * MB;
constexpr int MTHREADS = 256;
constexpr int MWARPS = MTHREADS / 32;

// Rows p and rows-1-p of a [rows, m] block are handled together: m+1 elements of the upper triangle per pair (sym, square
// matrix only) or 2m (all elements: full matrix or the row block of the oversized-PD path).
struct PairRange { int i1, j1, len1, i2, j2, len2; };
__device__ __forceinline__ PairRange pair_range(int p, int rows, int m, int sym) {
    PairRange r;
    r.i1 = p; r.j1 = sym ? p : 0; r.len1 = m - r.j1;
    r.i2 = rows - 1 - p; r.j2 = sym ? r.i2 : 0; r.len2 = (r.i2 > r.i1) ? m - r.j2 : 0;
    return r;
}

__global__ void __launch_bounds__(MTHREADS)
moment_stats_kernel(const double* __restrict__ D, int rows, int m, int sym, double* __restrict__ stat_part /*[gridDim.x][4]*/) {
    __shared__ double sh[4 * MWARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mn = INFINITY, mx = 0.0, nz = 0.0, nv = 0.0;
    const int npairs = (rows + 1) >> 1;
    for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
        const PairRange r = pair_range(p, rows, m, sym);
        for (int e = threadIdx.x; e < r.len1 + r.len2; e += MTHREADS) {
            const int i = e < r.len1 ? r.i1 : r.i2;
            const int j = e < r.len1 ? r.j1 + e : r.j2 + (e - r.len1);
            const double d = __ldg(D + (size_t)i * m + j);
            const double q = d * d;
            if (q == q) {
                const double w = (sym && j > i) ? 2.0 : 1.0;
                nv += w;
                if (q == 0.0) nz += w; else mn = fmin(mn, q);
                mx = fmax(mx, q);
            }
        }
    }
    mn = warp_min(mn); mx = warp_max(mx); nz = warp_sum(nz); nv = warp_sum(nv);
    if (lane == 0) { sh[warp] = mn; sh[MWARPS + warp] = mx; sh[2 * MWARPS + warp] = nz; sh[3 * MWARPS + warp] = nv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        mn = sh[0]; mx = sh[MWARPS]; nz = 0.0; nv = 0.0;
        for (int w = 0; w < MWARPS; ++w) {
            mn = fmin(mn, sh[w]); mx = fmax(mx, sh[MWARPS + w]); nz += sh[2 * MWARPS + w]; nv += sh[3 * MWARPS + w];
        }
        double* o = stat_part + 4 * (size_t)blockIdx.x;
        o[0] = mn; o[1] = mx; o[2] = nz; o[3] = nv;
    }
}

// One block: global statistics, bin edges and the k ranges of every bin.  status != 0 -> the general kernel must run.
// Requires non-increasing finite coefficients >= 0 and 0 < thr < inf (checked by the host).
__global__ void __launch_bounds__(32)
moment_plan_kernel(const double* __restrict__ stat_part, int n_parts, const double* __restrict__ coef, int n_eps, double thr,
                   double* __restrict__ plan) {
    // the per-CTA statistics by all 32 lanes (min / max, and counts that are exact integers in FP64: any order gives the same bits);
    // one thread walking the ~300 partials with dependent loads took 29 us
    double mn = INFINITY, mx = 0.0, nz = 0.0, nv = 0.0;
    for (int b = threadIdx.x; b < n_parts; b += 32) {
        const double* s = stat_part + 4 * (size_t)b;
        mn = fmin(mn, s[0]); mx = fmax(mx, s[1]); nz += s[2]; nv += s[3];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nz += __shfl_xor_sync(0xffffffffu, nz, o); nv += __shfl_xor_sync(0xffffffffu, nv, o);
    }
    if (threadIdx.x != 0) return;
    double rho = (thr > 4.0) ? 1.0 / (1.0 - 2.0 / thr) : 2.0;    // thr (rho-1)/(2 rho) <= 1; rho <= 2 keeps q - mu exact-ish
    if (!(rho <= 2.0)) rho = 2.0;
    int status = 0, B = 0;
    if (mn < INFINITY) {                                          // at least one non-zero element
        double lo = mn;
        while (B < MB) {
            const double hi = lo * rho;
            double* pb = plan + PLAN_HDR + PLAN_BIN * B;
            const double mu = 0.5 * (lo + hi), h = hi - mu;
            int a = 0, z = n_eps;                                 // ka = first k with coef[k]*lo < thr
            while (a < z) { const int mid = (a + z) >> 1; if (coef[mid] * lo < thr) z = mid; else a = mid + 1; }
            const int ka = a;
            a = 0; z = n_eps;                                     // kf = first k with coef[k]*hi < thr   (kf >= ka)
            while (a < z) { const int mid = (a + z) >> 1; if (coef[mid] * hi < thr) z = mid; else a = mid + 1; }
            const int kf = a < ka ? ka : a;
            if (kf - ka > MS) status = 2;                         // too many k straddle this bin
            if (!(h > 0.0) || !(hi < INFINITY)) status = 3;       // degenerate edges (overflow / underflow)
            pb[0] = lo; pb[1] = hi; pb[2] = mu; pb[3] = (h > 0.0) ? 1.0 / h : 0.0; pb[4] = (double)ka; pb[5] = (double)kf;
            ++B;
            if (hi > mx) break;                                   // every q <= mx < hi is covered
            lo = hi;
            if (B == MB) status = 1;                              // D^2 spans more than rho^MB
        }
    }
    if (!(mx < INFINITY)) status = 4;
    plan[0] = (double)status; plan[1] = (double)B; plan[2] = nz; plan[3] = nv; plan[4] = mn; plan[5] = mx; plan[6] = rho; plan[7] = 0.0;
}

__global__ void __launch_bounds__(MTHREADS)
moment_accum_kernel(const double* __restrict__ D, int rows, int m, int sym, const double* __restrict__ coef, double thr,
                    const double* __restrict__ plan, double* __restrict__ part /*[gridDim.x][MB][MV]*/) {
    __shared__ double sh[MWARPS * MV];
    if (plan[0] != 0.0) return;                                   // the host falls back to sweep_kernel
    const int B = (int)plan[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npairs = (rows + 1) >> 1;
    const double SMALL = 7.450580596923828e-09;                   // as in sweep_kernel
    for (int b = 0; b < B; ++b) {
        const double* pb = plan + PLAN_HDR + PLAN_BIN * b;
        const double lo = pb[0], hi = pb[1], mu = pb[2], sc = pb[3];
        const int ka = (int)pb[4], ns = (int)pb[5] - ka;
        double cs[MS];
#pragma unroll
        for (int s = 0; s < MS; ++s) cs[s] = (s < ns) ? coef[ka + s] : 0.0;
        double acc[MP + 1], st[MS];
#pragma unroll
        for (int p = 0; p <= MP; ++p) acc[p] = 0.0;
#pragma unroll
        for (int s = 0; s < MS; ++s) st[s] = 0.0;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
            const PairRange r = pair_range(p, rows, m, sym);
#pragma unroll 2
            for (int e = threadIdx.x; e < r.len1 + r.len2; e += MTHREADS) {
                const int i = e < r.len1 ? r.i1 : r.i2;
                const int j = e < r.len1 ? r.j1 + e : r.j2 + (e - r.len1);
                const double d = __ldg(D + (size_t)i * m + j);
                const double q = d * d;
                // membership: lo <= q < hi (the last bin has hi > max q); zeros and NaN belong to no bin
                if (!(q >= lo && q < hi) || q == 0.0) continue;
                const double w = (sym && j > i) ? 2.0 : 1.0;
                const double u = (q - mu) * sc;
                double pw = w;
                acc[0] += pw;
#pragma unroll
                for (int pp = 1; pp <= MP; ++pp) { pw *= u; acc[pp] += pw; }
                if (ns > 0) {
#pragma unroll
                    for (int s = 0; s < MS; ++s) {
                        if (s < ns) {
                            const double t = cs[s] * q;
                            if (t < thr) {
                                const double term = (t < SMALL) ? fma(t, fma(t, 0.5, -1.0), 1.0) : exp_neg_fast(fmin(t, 700.0));
                                st[s] = fma(w, term, st[s]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();                                          // sh is reused from the previous bin
#pragma unroll
        for (int p = 0; p <= MP; ++p) { const double v = warp_sum(acc[p]); if (lane == 0) sh[warp * MV + p] = v; }
#pragma unroll
        for (int s = 0; s < MS; ++s) { const double v = warp_sum(st[s]); if (lane == 0) sh[warp * MV + MP + 1 + s] = v; }
        __syncthreads();
        if (threadIdx.x < MV) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < MWARPS; ++w) t += sh[w * MV + threadIdx.x];
            part[((size_t)blockIdx.x * MB + b) * MV + threadIdx.x] = t;
        }
    }
}

__global__ void __launch_bounds__(1024)
moment_final_kernel(const double* __restrict__ part, int n_blocks, const double* __restrict__ coef, int n_eps, double thr,
                    const double* __restrict__ plan, double* __restrict__ out) {
    __shared__ double tot[MB * MV];
    __shared__ double pl[PLAN_LEN];
    if (plan[0] != 0.0) return;
    const int B = (int)plan[1];
    for (int i = threadIdx.x; i < PLAN_HDR + PLAN_BIN * B; i += blockDim.x) pl[i] = plan[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int v = warp; v < B * MV; v += nwarps) {                 // v = b*MV + value
        const int b = v / MV, x = v - b * MV;
        double s = 0.0;
        for (int blk = lane; blk < n_blocks; blk += 32) s += part[((size_t)blk * MB + b) * MV + x];
        s = warp_sum(s);
        if (lane == 0) tot[v] = s;
    }
    __syncthreads();
    const double zero_term = (0.0 < thr) ? pl[2] : 0.0;
    for (int k = threadIdx.x; k < n_eps; k += blockDim.x) {
        const double c = coef[k];
        double S = zero_term;
        for (int b = 0; b < B; ++b) {
            const double* pb = pl + PLAN_HDR + PLAN_BIN * b;
            const int ka = (int)pb[4], kf = (int)pb[5];
            const double* M = tot + b * MV;
            if (k < ka) continue;
            if (k < kf) { S += M[MP + 1 + (k - ka)]; continue; }
            const double h = pb[1] - pb[2];                       // hi - mu, the h the moments were scaled with (sc = 1/h)
            const double x = -(c * h);
            double poly = M[MP];
#pragma unroll
            for (int p = MP - 1; p >= 0; --p) poly = fma(poly, x / (double)(p + 1), M[p]);
            S += exp(-(c * pb[2])) * poly;
        }
        out[k] = log(S);
    }
}


// Row-block path: tot[b][v] = sum over blocks of part[blk][b][v] (fixed order), the table a rank contributes to the all-reduce.
__global__ void __launch_bounds__(1024)
moment_table_kernel(const double* __restrict__ part, int n_blocks, int B, double* __restrict__ tot /*[MB][MV]*/) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int v = warp; v < MB * MV; v += nwarps) {
        const int b = v / MV, x = v - b * MV;
        double s = 0.0;
        if (b < B) {
            for (int blk = lane; blk < n_blocks; blk += 32) s += part[((size_t)blk * MB + b) * MV + x];
            s = warp_sum(s);
        }
        if (lane == 0) tot[v] = s;
    }
}

}  // namespace sweep

// Oversized-PD path, step 1 of the moment sweep: {min non-zero q, max q, weight of exact zeros, weight of valid elements} of the
// resident [rows, m] block (every element weighs 1).  The ranks all-reduce these (min, max, sum, sum).
int run_moment_stats_rows(esper_ctx* c, int rows, int m, double* out4_h) {
    using namespace sweep;
    int grid = c->sm_count * 2;
    if (grid > 480) grid = 480;
    if (grid > (rows + 1) / 2) grid = (rows + 1) / 2;
    if (grid < 1) grid = 1;
    ESPER_RESERVE(c, c->sweep_aux, sizeof(double) * 8192);
    double* stat_d = c->sweep_aux.as<double>();
    {
        LaunchScope ls(c, "sweep");
        moment_stats_kernel<<<grid, MTHREADS, 0, c->stream>>>(c->dist.as<double>(), rows, m, 0, stat_d);
    }
    ESPER_CUDA(c, cudaGetLastError());
    std::vector<double> st((size_t)4 * grid);
    ESPER_CUDA(c, cudaMemcpyAsync(st.data(), stat_d, sizeof(double) * st.size(), cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    double mn = INFINITY, mx = 0.0, nz = 0.0, nv = 0.0;
    for (int b = 0; b < grid; ++b) { mn = fmin(mn, st[4 * b]); mx = fmax(mx, st[4 * b + 1]); nz += st[4 * b + 2]; nv += st[4 * b + 3]; }
    out4_h[0] = mn; out4_h[1] = mx; out4_h[2] = nz; out4_h[3] = nv;
    return ESPER_OK;
}

// Step 2: moments and boundary sums of the resident block against the caller's plan (manifoldem_esper_b200/moment_sweep.py:
// made from the all-reduced statistics, so every rank uses the same bins).  table_h [MB][MV].
int run_moment_accum_rows(esper_ctx* c, int rows, int m, const double* coef_h, int n_eps, double thr, const double* plan_h,
                          double* table_h) {
    using namespace sweep;
    if (plan_h[0] != 0.0) return fail(c, ESPER_E_ARG, "moment_accum_rows: the plan is not usable (status %d)", (int)plan_h[0]);
    const int B = (int)plan_h[1];
    if (B < 0 || B > MB) return fail(c, ESPER_E_ARG, "moment_accum_rows: %d bins", B);
    for (int b = 0; b < B; ++b) {
        const int ka = (int)plan_h[PLAN_HDR + PLAN_BIN * b + 4], kf = (int)plan_h[PLAN_HDR + PLAN_BIN * b + 5];
        if (ka < 0 || kf < ka || kf - ka > MS || kf > n_eps) return fail(c, ESPER_E_ARG, "moment_accum_rows: bad k range in bin %d", b);
    }
    int grid = c->sm_count * 2;
    if (grid > 480) grid = 480;
    if (grid > (rows + 1) / 2) grid = (rows + 1) / 2;
    if (grid < 1) grid = 1;
    ESPER_RESERVE(c, c->sweep_part, sizeof(double) * (size_t)grid * MB * MV);
    ESPER_RESERVE(c, c->sweep_aux, sizeof(double) * ((size_t)n_eps + 8192));
    double* plan_d = c->sweep_aux.as<double>() + 2048;
    double* table_d = c->sweep_aux.as<double>() + 4096;            // [MB * MV] = 672 doubles
    double* coef_d = c->sweep_aux.as<double>() + 8192;
    static_assert(MB * MV <= 4000, "sweep_aux layout");
    ESPER_CUDA(c, cudaMemcpyAsync(coef_d, coef_h, sizeof(double) * n_eps, cudaMemcpyHostToDevice, c->stream));
    ESPER_CUDA(c, cudaMemcpyAsync(plan_d, plan_h, sizeof(double) * PLAN_LEN, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, "sweep");
        moment_accum_kernel<<<grid, MTHREADS, 0, c->stream>>>(c->dist.as<double>(), rows, m, 0, coef_d, thr, plan_d,
                                                              c->sweep_part.as<double>());
    }
    {
        LaunchScope ls(c, "sweep");
        moment_table_kernel<<<1, 1024, 0, c->stream>>>(c->sweep_part.as<double>(), grid, B, table_d);
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(table_h, table_d, sizeof(double) * MB * MV, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

void moment_layout_host(int* out5) {
    out5[0] = sweep::MP; out5[1] = sweep::MS; out5[2] = sweep::MB; out5[3] = sweep::PLAN_HDR; out5[4] = sweep::PLAN_BIN;
}

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

// Moment sweep of the full resident matrix (see the comment above moment_stats_kernel).  Returns ESPER_OK with
// *done = 1 when logsum_h holds the result, *done = 0 when the data need the general kernel (D^2 too spread out, or
// more than MS coefficients per bin boundary).
static int run_sweep_moments(esper_ctx* c, int m, int sym, const double* coef_h, int n_eps, double thr, double* logsum_h, int* done) {
    using namespace sweep;
    *done = 0;
    const int npairs = (m + 1) / 2;
    int grid = c->sm_count * 2;
    if (grid > 480) grid = 480;
    if (grid > npairs) grid = npairs;
    if (grid < 1) grid = 1;
    ESPER_RESERVE(c, c->sweep_part, sizeof(double) * (size_t)grid * MB * MV);
    ESPER_RESERVE(c, c->sweep_aux, sizeof(double) * ((size_t)n_eps * 2 + 4096) + sizeof(int) * 64);
    double* stat_d = c->sweep_aux.as<double>();                   // [grid][4], grid <= 480
    double* plan_d = c->sweep_aux.as<double>() + 2048;            // [PLAN_LEN] (the symmetry flag lives at 4000)
    double* coef_d = c->sweep_aux.as<double>() + 4096;
    double* logsum_d = coef_d + n_eps;
    static_assert(2048 + PLAN_LEN <= 4000 && 480 * 4 <= 2048, "sweep_aux layout");
    ESPER_CUDA(c, cudaMemcpyAsync(coef_d, coef_h, sizeof(double) * n_eps, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, "sweep");
        moment_stats_kernel<<<grid, MTHREADS, 0, c->stream>>>(c->dist.as<double>(), m, m, sym, stat_d);
    }
    {
        LaunchScope ls(c, "sweep");
        moment_plan_kernel<<<1, 32, 0, c->stream>>>(stat_d, grid, coef_d, n_eps, thr, plan_d);
    }
    {
        LaunchScope ls(c, "sweep");
        moment_accum_kernel<<<grid, MTHREADS, 0, c->stream>>>(c->dist.as<double>(), m, m, sym, coef_d, thr, plan_d,
                                                              c->sweep_part.as<double>());
    }
    {
        LaunchScope ls(c, "sweep");
        moment_final_kernel<<<1, 1024, 0, c->stream>>>(c->sweep_part.as<double>(), grid, coef_d, n_eps, thr, plan_d, logsum_d);
    }
    ESPER_CUDA(c, cudaGetLastError());
    double hdr[PLAN_HDR];
    ESPER_CUDA(c, cudaMemcpyAsync(hdr, plan_d, sizeof hdr, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaMemcpyAsync(logsum_h, logsum_d, sizeof(double) * n_eps, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    if (getenv("ESPER_DEBUG_SWEEP"))
        fprintf(stderr, "[sweep moments] status %d bins %d zeros %.0f valid %.0f min %.6e max %.6e rho %.6f\n", (int)hdr[0], (int)hdr[1],
                hdr[2], hdr[3], hdr[4], hdr[5], hdr[6]);
    *done = (hdr[0] == 0.0) ? 1 : 0;
    if (*done) { c->sweep_last_path = 1; c->sweep_last_bins = (int)hdr[1]; }
    return ESPER_OK;
}

// rows x m block of D resident in c->dist; take_log = 0 returns the partial sums (row-block path).
int run_sweep_rows(esper_ctx* c, int rows, int m, const double* coef_h, int n_eps, double thr, double* logsum_h, int take_log) {
    using namespace sweep;
    int sym = 0;
    if (rows == m && c->dist_row0 < 0) {
        if (c->dist_symmetric < 0) { int rc = check_symmetry(c, m); if (rc) return rc; }
        sym = c->dist_symmetric > 0 ? 1 : 0;
    }
    int mono = 1;                                  // non-increasing, finite, non-negative coefficients?
    for (int k = 0; k < n_eps; ++k)
        if (!(coef_h[k] >= 0.0) || !(coef_h[k] < INFINITY) || (k > 0 && coef_h[k] > coef_h[k - 1])) { mono = 0; break; }
    c->sweep_last_path = 0; c->sweep_last_bins = 0;
    if (c->sweep_mode == 1 && rows == m && c->dist_row0 < 0 && take_log && mono && thr > 0.0 && thr < INFINITY) {
        int done = 0;
        const int rc = run_sweep_moments(c, m, sym, coef_h, n_eps, thr, logsum_h, &done);
        if (rc) return rc;
        if (done) return ESPER_OK;
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
