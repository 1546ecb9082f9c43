// k7_conic.cu -- stage-4 preparation on the device:
//   * manifold normalisation, 2_Manifold_Rotations/Rotation_Histograms.py:100-112 (per-column to [0,1], then the
//     whole matrix to [-1,1]) -- IEEE basic operations only, so the result is bit-identical to numpy's;
//   * the conic fits of findParabolas (:153-212): 3_Manifold_Binning/ConicFit.py:205-264 (`fit2`, general conic
//     x^2 + Bxy + Cy^2 + Dx + Ey + F = 0) and :267-283 (`fit3`, parabola y = a x^2 + b x + c) for EVERY column pair
//     v1 < v2 of every PD in one launch.  The reference solves the 5x5 (3x3) normal equations with sympy's
//     `Matrix.rref`, which for floating-point entries is a fraction-free Gauss-Jordan elimination with
//     largest-magnitude pivots; `rref_solve` below restates it operation for operation (no FMA contraction).
//     Moment sums are FP64 tree reductions (the reference: OpenBLAS dgemm order), so R^2 agrees to rounding
//     (~1e-13), not bit for bit; the decisions taken from it (best partner per eigenfunction) are what the tests pin.
#include "common.cuh"
#include "devutil.cuh"

namespace esper {
namespace conic {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

// ---- normalisation ------------------------------------------------------------------------------
// Pass A: one CTA per (column, PD): d2 = (x - min) / max(x - min); records min/max of d2 (0 and 1 unless the
// column is constant, where the reference produces NaN as well).
__global__ void __launch_bounds__(THREADS)
normalize_columns_kernel(const double* __restrict__ U0, int m, int ncol, double* __restrict__ out, double* __restrict__ colstat) {
    const int col = blockIdx.x, pd = blockIdx.y;
    const double* src = U0 + (size_t)pd * m * ncol + col;
    double* dst = out + (size_t)pd * m * ncol + col;
    __shared__ double red[WARPS];
    __shared__ double bc;
    auto block_reduce = [&](double v, bool want_max) -> double {
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = want_max ? fmax(v, w) : fmin(v, w);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = red[0];
            for (int w = 1; w < WARPS; ++w) r = want_max ? fmax(r, red[w]) : fmin(r, red[w]);
            bc = r;
        }
        __syncthreads();
        const double r = bc;
        __syncthreads();
        return r;
    };
    double mn = INFINITY;
    for (int i = threadIdx.x; i < m; i += THREADS) mn = fmin(mn, src[(size_t)i * ncol]);
    mn = block_reduce(mn, false);
    double mx = -INFINITY;
    for (int i = threadIdx.x; i < m; i += THREADS) mx = fmax(mx, __dsub_rn(src[(size_t)i * ncol], mn));
    mx = block_reduce(mx, true);                   // np.ptp of the shifted column (its minimum is exactly 0)
    double lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < m; i += THREADS) {
        const double d = __ddiv_rn(__dsub_rn(src[(size_t)i * ncol], mn), mx);
        dst[(size_t)i * ncol] = d;
        lo = fmin(lo, d); hi = fmax(hi, d);
    }
    lo = block_reduce(lo, false);
    hi = block_reduce(hi, true);
    if (threadIdx.x == 0) { colstat[((size_t)pd * ncol + col) * 2] = lo; colstat[((size_t)pd * ncol + col) * 2 + 1] = hi; }
}

// Pass B: d = 2 (d - gmin) / gptp - 1 over the whole matrix of a PD; only columns [first, first + dim) are kept.
__global__ void __launch_bounds__(THREADS)
normalize_global_kernel(const double* __restrict__ tmp, const double* __restrict__ colstat, int m, int ncol, int first, int dim,
                        double* __restrict__ Uinit) {
    const int pd = blockIdx.y;
    double gmin = INFINITY, gmax = -INFINITY;
    for (int c = 0; c < ncol; ++c) {
        gmin = fmin(gmin, colstat[((size_t)pd * ncol + c) * 2]);
        gmax = fmax(gmax, colstat[((size_t)pd * ncol + c) * 2 + 1]);
    }
    const double gptp = __dsub_rn(gmax, gmin);
    const long long tot = (long long)m * dim;
    for (long long e = (long long)blockIdx.x * THREADS + threadIdx.x; e < tot; e += (long long)gridDim.x * THREADS) {
        const int i = (int)(e / dim), c = (int)(e % dim);
        const double d = tmp[((size_t)pd * m + i) * ncol + first + c];
        Uinit[((size_t)pd * m + i) * dim + c] = __dsub_rn(__ddiv_rn(__dmul_rn(2.0, __dsub_rn(d, gmin)), gptp), 1.0);
    }
}

// ---- conic fits -----------------------------------------------------------------------------------
// Last column of the reduced row-echelon form of the augmented normal equations [AtA | Atb] (NC x (NC+1)).
template <int NC>
__device__ void rref_solve(double (&M)[NC][NC + 1], double (&sol)[NC]) {
    int piv_row = 0, n_piv = 0;
    int pivots[NC];
    for (int pc = 0; pc < NC + 1 && piv_row < NC; ++pc) {
        int best = piv_row; double bv = fabs(M[piv_row][pc]);
        for (int r = piv_row + 1; r < NC; ++r) if (fabs(M[r][pc]) > bv) { bv = fabs(M[r][pc]); best = r; }
        if (bv == 0.0) continue;
        if (best != piv_row)
            for (int c = 0; c < NC + 1; ++c) { const double t = M[piv_row][c]; M[piv_row][c] = M[best][c]; M[best][c] = t; }
        pivots[n_piv++] = pc;
        const double pv = M[piv_row][pc];
        for (int r = 0; r < NC; ++r) {
            if (r == piv_row) continue;
            const double val = M[r][pc];
            if (val == 0.0) continue;
            for (int c = 0; c < NC + 1; ++c) M[r][c] = __dsub_rn(__dmul_rn(pv, M[r][c]), __dmul_rn(val, M[piv_row][c]));
        }
        ++piv_row;
    }
    for (int i = 0; i < n_piv; ++i) {
        const int j = pivots[i];
        const double pv = M[i][j];
        for (int c = j + 1; c < NC + 1; ++c) M[i][c] = __ddiv_rn(M[i][c], pv);
        M[i][j] = 1.0;
    }
    for (int i = 0; i < NC; ++i) sol[i] = M[i][NC];
}

template <int NC> struct Design;
template <> struct Design<5> {           // fit2: columns (xy, y^2, x, y, 1), right-hand side -(x^2)
    __device__ static void row(double x, double y, double (&a)[5], double& b) {
        a[0] = __dmul_rn(x, y); a[1] = __dmul_rn(y, y); a[2] = x; a[3] = y; a[4] = 1.0; b = -__dmul_rn(x, x);
    }
};
template <> struct Design<3> {           // fit3: columns (x^2, x, 1), right-hand side y
    __device__ static void row(double x, double y, double (&a)[3], double& b) {
        a[0] = __dmul_rn(x, x); a[1] = x; a[2] = 1.0; b = y;
    }
};

// One CTA per (pair, PD).  R2 [n_pd, n_pairs]; coef [n_pd, n_pairs, 5] (fit3: first three entries).
template <int NC>
__global__ void __launch_bounds__(THREADS)
conic_fit_kernel(const double* __restrict__ U, int m, int dim, double* __restrict__ R2, double* __restrict__ coef) {
    constexpr int NS = NC * (NC + 1) / 2 + NC;
    const int pair = blockIdx.x, pd = blockIdx.y, n_pairs = gridDim.x;
    int v1 = 0, rem = pair;
    while (rem >= dim - 1 - v1) { rem -= dim - 1 - v1; ++v1; }
    const int v2 = v1 + 1 + rem;
    const double* base = U + (size_t)pd * m * dim;
    __shared__ double red[WARPS][NS];
    __shared__ double cf[NC + 1];                 // coefficients, then mean(b)

    double s[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) s[q] = 0.0;
    for (int i = threadIdx.x; i < m; i += THREADS) {
        double a[NC], b;
        Design<NC>::row(base[(size_t)i * dim + v1], base[(size_t)i * dim + v2], a, b);
        int q = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k)
#pragma unroll
            for (int l = k; l < NC; ++l) s[q++] += a[k] * a[l];
#pragma unroll
        for (int k = 0; k < NC; ++k) s[q++] += a[k] * b;
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        double v = s[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[NS];
        for (int q = 0; q < NS; ++q) { double v = red[0][q]; for (int w = 1; w < WARPS; ++w) v += red[w][q]; tot[q] = v; }
        double M[NC][NC + 1], sol[NC];
        int q = 0;
        for (int k = 0; k < NC; ++k)
            for (int l = k; l < NC; ++l) { M[k][l] = tot[q]; M[l][k] = tot[q]; ++q; }
        for (int k = 0; k < NC; ++k) M[k][NC] = tot[q++];
        const double sum_b = M[NC - 1][NC];       // the last design column is the constant 1
        rref_solve<NC>(M, sol);
        for (int k = 0; k < NC; ++k) cf[k] = sol[k];
        cf[NC] = sum_b / (double)m;
    }
    __syncthreads();
    double c[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) c[k] = cf[k];
    const double mean_b = cf[NC];
    double ssr = 0.0, tss = 0.0;
    for (int i = threadIdx.x; i < m; i += THREADS) {
        double a[NC], b;
        Design<NC>::row(base[(size_t)i * dim + v1], base[(size_t)i * dim + v2], a, b);
        double fit = __dmul_rn(a[0], c[0]);
#pragma unroll
        for (int k = 1; k < NC; ++k) fit = __dadd_rn(fit, __dmul_rn(a[k], c[k]));
        const double r = b - fit, t = b - mean_b;
        ssr += r * r; tss += t * t;
    }
    for (int o = 16; o > 0; o >>= 1) { ssr += __shfl_xor_sync(0xffffffffu, ssr, o); tss += __shfl_xor_sync(0xffffffffu, tss, o); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = ssr; red[threadIdx.x >> 5][1] = tss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = red[0][0], t = red[0][1];
        for (int w = 1; w < WARPS; ++w) { a += red[w][0]; t += red[w][1]; }
        const size_t o = (size_t)pd * n_pairs + pair;
        R2[o] = 1.0 - a / t;
        if (coef) {
            for (int k = 0; k < 5; ++k) coef[o * 5 + k] = k < NC ? c[k] : 0.0;
        }
    }
}

}  // namespace conic

// U0 [n_pd, m, ncol] (host) -> Uinit [n_pd, m, dim] (host, optional) and device-resident in c->rot_U for the fits.
int run_manifold_prepare(esper_ctx* c, const double* U0, int n_pd, int m, int ncol, int first, int dim, int kind,
                         double* Uinit_h, double* R2_h, double* coef_h) {
    using namespace conic;
    if (n_pd < 1 || m < 1 || ncol < 1 || dim < 2 || first < 0 || first + dim > ncol)
        return fail(c, ESPER_E_ARG, "manifold_prepare: need n_pd, m >= 1, dim >= 2 and first + dim <= ncol (got ncol=%d, first=%d, dim=%d)",
                    ncol, first, dim);
    if (n_pd > 65535) return fail(c, ESPER_E_ARG, "manifold_prepare: at most 65535 PDs per call");
    if (kind != 0 && kind != 2 && kind != 3) return fail(c, ESPER_E_ARG, "manifold_prepare: kind must be 0 (no fits), 2 (fit2) or 3 (fit3)");
    const size_t n_in = (size_t)n_pd * m * ncol, n_out = (size_t)n_pd * m * dim;
    ESPER_RESERVE(c, c->rot_R, sizeof(double) * n_in);                     // raw manifolds
    ESPER_RESERVE(c, c->rot_H, sizeof(double) * (n_in + (size_t)n_pd * ncol * 2));   // column-normalised copy + column stats
    ESPER_RESERVE(c, c->rot_U, sizeof(double) * n_out);
    ESPER_CUDA(c, cudaMemcpyAsync(c->rot_R.p, U0, sizeof(double) * n_in, cudaMemcpyHostToDevice, c->stream));
    double* tmp = c->rot_H.as<double>();
    double* colstat = tmp + n_in;
    {
        LaunchScope ls(c, "conic");
        normalize_columns_kernel<<<dim3(ncol, n_pd), THREADS, 0, c->stream>>>(c->rot_R.as<double>(), m, ncol, tmp, colstat);
    }
    {
        LaunchScope ls(c, "conic");
        const int gx = ceil_div((long long)m * dim, THREADS * 4);
        normalize_global_kernel<<<dim3(gx < 1 ? 1 : gx, n_pd), THREADS, 0, c->stream>>>(tmp, colstat, m, ncol, first, dim, c->rot_U.as<double>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    if (Uinit_h) ESPER_CUDA(c, cudaMemcpyAsync(Uinit_h, c->rot_U.p, sizeof(double) * n_out, cudaMemcpyDeviceToHost, c->stream));
    if (kind != 0) {
        int rc = run_conic_fits(c, nullptr, n_pd, m, dim, kind, R2_h, coef_h);
        if (rc) return rc;
    }
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

// Uinit == nullptr: the manifolds already resident in c->rot_U (after run_manifold_prepare).
int run_conic_fits(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, int kind, double* R2_h, double* coef_h) {
    using namespace conic;
    if (n_pd < 1 || n_pd > 65535 || m < 1 || dim < 2) return fail(c, ESPER_E_ARG, "conic_fit_batch: need 1 <= n_pd <= 65535, m >= 1, dim >= 2");
    if (kind != 2 && kind != 3) return fail(c, ESPER_E_ARG, "conic_fit_batch: kind must be 2 (ConicFit.fit2) or 3 (ConicFit.fit3)");
    if (!R2_h) return fail(c, ESPER_E_ARG, "conic_fit_batch: R2 is NULL");
    const int n_pairs = dim * (dim - 1) / 2;
    const size_t n_out = (size_t)n_pd * m * dim;
    if (Uinit) {
        ESPER_RESERVE(c, c->rot_U, sizeof(double) * n_out);
        ESPER_CUDA(c, cudaMemcpyAsync(c->rot_U.p, Uinit, sizeof(double) * n_out, cudaMemcpyHostToDevice, c->stream));
    } else if (c->rot_U.cap < sizeof(double) * n_out) {
        return fail(c, ESPER_E_STATE, "conic_fit_batch: Uinit == NULL but no resident manifolds");
    }
    const size_t n_r2 = (size_t)n_pd * n_pairs;
    ESPER_RESERVE(c, c->rot_aux, sizeof(double) * n_r2 * 6);
    double* R2_d = c->rot_aux.as<double>();
    double* coef_d = R2_d + n_r2;
    {
        LaunchScope ls(c, "conic");
        if (kind == 2) conic_fit_kernel<5><<<dim3(n_pairs, n_pd), THREADS, 0, c->stream>>>(c->rot_U.as<double>(), m, dim, R2_d, coef_d);
        else           conic_fit_kernel<3><<<dim3(n_pairs, n_pd), THREADS, 0, c->stream>>>(c->rot_U.as<double>(), m, dim, R2_d, coef_d);
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(R2_h, R2_d, sizeof(double) * n_r2, cudaMemcpyDeviceToHost, c->stream));
    if (coef_h) ESPER_CUDA(c, cudaMemcpyAsync(coef_h, coef_d, sizeof(double) * n_r2 * 5, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

}  // namespace esper
