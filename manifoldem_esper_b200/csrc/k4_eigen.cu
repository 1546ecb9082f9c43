// k4_eigen.cu -- K4: leading eigenpairs of the symmetric Markov matrix Ms (m x m, FP64), replacing the
// full SVD of 1_Embedding/DM/DiffusionMaps.py:166-169 (Ms is symmetric PSD, so singular values =
// eigenvalues and U = eigenvectors; the reference saves s**2 and U).
//
// Method: Lanczos with full reorthogonalisation (classical Gram-Schmidt applied twice) on the
// operator deflated by the analytically known top pair (lambda_0 = 1, v_0 = sqrt(d)/|sqrt(d)|),
// which is kept as a locked vector in the orthogonalisation basis.  Every kernel is a streaming
// pass: SYMV reads Ms once per step (m^2 x 8 B, L2-resident at m = 2000), the orthogonalisation
// reads the basis twice.  The small tridiagonal eigenproblem (bisection + inverse iteration) runs
// on the host every few steps to test convergence; Ritz vectors are assembled on the device.
#include "common.cuh"
#include "devutil.cuh"
#include <algorithm>

namespace esper {
namespace lanczos {

// w = Ms q, one warp per row.
__global__ void __launch_bounds__(256) symv_kernel(const double* __restrict__ Ms, int m,
                                                   const double* __restrict__ q, double* __restrict__ w) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const double* a = Ms + (size_t)row * m;
    double s = 0.0;
    if ((m & 1) == 0) {
        const double2* a2 = reinterpret_cast<const double2*>(a);
        const double2* q2 = reinterpret_cast<const double2*>(q);
        for (int j = lane; j < (m >> 1); j += 32) {
            const double2 x = __ldg(a2 + j), y = __ldg(q2 + j);
            s = fma(x.x, y.x, s);
            s = fma(x.y, y.y, s);
        }
    } else {
        for (int j = lane; j < m; j += 32) s = fma(__ldg(a + j), __ldg(q + j), s);
    }
    s = warp_sum(s);
    if (lane == 0) w[row] = s;
}

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

// U[r][1+i] = sum_l V_{1+l}[r] S[l][i]; U[r][0] = V_0[r].  One thread per row, S staged in shared memory.
template <int KMAX>
__global__ void __launch_bounds__(128) ritz_kernel(const double* __restrict__ V, int m, int steps,
                                                   const double* __restrict__ S, int kw, double* __restrict__ U, int k) {
    __shared__ double s_tile[32 * KMAX];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) acc[i] = 0.0;
    for (int l0 = 0; l0 < steps; l0 += 32) {
        const int nl = min(32, steps - l0);
        __syncthreads();
        for (int t = threadIdx.x; t < nl * kw; t += blockDim.x) s_tile[(t / kw) * KMAX + (t % kw)] = S[(size_t)(l0 + t / kw) * kw + (t % kw)];
        __syncthreads();
        if (r < m) {
            for (int l = 0; l < nl; ++l) {
                const double v = __ldg(V + (size_t)(1 + l0 + l) * m + r);
#pragma unroll
                for (int i = 0; i < KMAX; ++i) if (i < kw) acc[i] = fma(v, s_tile[l * KMAX + i], acc[i]);
            }
        }
    }
    if (r < m) {
        U[(size_t)r * k] = V[r];
        for (int i = 0; i < kw; ++i) U[(size_t)r * k + 1 + i] = acc[i];
    }
}

// Normalise each column to unit length and make its largest-magnitude entry positive (first on ties).
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

int run_lanczos(esper_ctx* c, int m, int k, double tol, int max_iter, double* val_h, double* vec_h, int* iters_out) {
    using namespace lanczos;
    constexpr int KMAX = 64;
    if (k < 1 || k > m || k > KMAX) return fail(c, ESPER_E_ARG, "k=%d out of range (1..min(m,%d))", k, KMAX);
    if (tol <= 0.0) tol = 1e-11;
    const int kw = k - 1;                               // non-trivial pairs wanted
    int guard = (kw > 0) ? std::min(2, m - 1 - kw) : 0;
    if (guard < 0) guard = 0;
    int cap = m - 1;                                    // Krylov dimension of the deflated operator
    if (max_iter <= 0) max_iter = 1024;
    if (max_iter > cap) max_iter = cap;
    const double* Ms = c->markov.as<double>();
    ESPER_RESERVE(c, c->lan_V, sizeof(double) * (size_t)(max_iter + 3) * m);
    ESPER_RESERVE(c, c->lan_w, sizeof(double) * (size_t)m * 2);
    ESPER_RESERVE(c, c->lan_h, sizeof(double) * (size_t)(max_iter + 4) * 3);
    double* V = c->lan_V.as<double>();
    double* w = c->lan_w.as<double>();
    double* w2 = w + m;
    double* h = c->lan_h.as<double>();
    double* alpha_d = h + (max_iter + 4);
    double* beta_d = alpha_d + (max_iter + 4);
    cudaStream_t st = c->stream;
    ESPER_CUDA(c, cudaMemsetAsync(alpha_d, 0, sizeof(double) * (size_t)(max_iter + 4) * 2, st));

    // locked vector v0 and the start vector q1 (orthogonal to v0)
    { LaunchScope ls(c, "lanczos"); init_kernel<<<ceil_div(m, 256), 256, 0, st>>>(c->degree.as<double>(), m, w, w2); }
    { LaunchScope ls(c, "lanczos"); normalize_kernel<<<1, 1024, 0, st>>>(w, m, V, nullptr); }
    if (kw > 0) {
        for (int pass = 0; pass < 2; ++pass) {
            { LaunchScope ls(c, "lanczos"); dots_kernel<<<1, 256, 0, st>>>(V, m, w2, h); }
            { LaunchScope ls(c, "lanczos"); update_kernel<<<ceil_div(m, 256), 256, 0, st>>>(V, m, 1, h, w2, nullptr); }
        }
        { LaunchScope ls(c, "lanczos"); normalize_kernel<<<1, 1024, 0, st>>>(w2, m, V + m, nullptr); }
    }

    std::vector<double> a, b, theta, S;
    int steps = 0;
    bool converged = (kw == 0);
    int next_check = std::min(max_iter, std::max(2 * (kw + guard) + 8, 32));
    std::vector<double> ab((size_t)(max_iter + 4) * 2);
    while (!converged && steps < max_iter) {
        const int j = steps + 1;                        // current Lanczos vector is V[j]
        { LaunchScope ls(c, "lanczos"); symv_kernel<<<ceil_div(m, 8), 256, 0, st>>>(Ms, m, V + (size_t)j * m, w); }
        for (int pass = 0; pass < 2; ++pass) {
            { LaunchScope ls(c, "lanczos"); dots_kernel<<<j + 1, 256, 0, st>>>(V, m, w, h); }
            { LaunchScope ls(c, "lanczos"); update_kernel<<<ceil_div(m, 256), 256, 0, st>>>(V, m, j + 1, h, w, alpha_d + steps); }
        }
        { LaunchScope ls(c, "lanczos"); normalize_kernel<<<1, 1024, 0, st>>>(w, m, V + (size_t)(j + 1) * m, beta_d + steps); }
        ++steps;
        if (steps >= next_check || steps == max_iter) {
            ESPER_CUDA(c, cudaMemcpyAsync(ab.data(), alpha_d, sizeof(double) * (size_t)(max_iter + 4) * 2,
                                          cudaMemcpyDeviceToHost, st));
            ESPER_CUDA(c, cudaStreamSynchronize(st));
            a.assign(ab.begin(), ab.begin() + steps);
            b.assign(ab.begin() + (max_iter + 4), ab.begin() + (max_iter + 4) + steps);
            const int want = std::min(kw + guard, steps);
            tridiag_top_eigenvalues(a, b, steps, want, theta);
            tridiag_eigenvectors(a, b, steps, theta, S);
            const double scale = std::max(fabs(theta[0]), 1e-300);
            bool ok = (steps >= kw);
            const double beta_last = b[steps - 1];
            bool breakdown = !(beta_last > 1e-14 * scale);
            for (int i = 0; i < want && ok; ++i) {
                const double res = fabs(beta_last * S[(size_t)(steps - 1) * want + i]);
                const double lim = (i < kw ? tol : 1e-5) * scale;
                if (!(res <= lim)) ok = false;
            }
            if (ok || breakdown || steps >= cap) converged = true;
            next_check = steps + 16;
        }
    }
    if (iters_out) *iters_out = steps;
    if (!converged) return fail(c, ESPER_E_NOCONV, "Lanczos did not converge in %d steps", steps);

    // Ritz vectors U = [v0, V_1..V_steps S]
    ESPER_RESERVE(c, c->lan_small, sizeof(double) * ((size_t)m * k + (size_t)(max_iter + 4) * KMAX));
    double* U = c->lan_small.as<double>();
    double* S_d = U + (size_t)m * k;
    std::vector<double> Sk;
    if (kw > 0) {
        const int want = (int)theta.size();
        if (want < kw) return fail(c, ESPER_E_NOCONV, "Krylov space exhausted: %d Ritz pairs < %d wanted", want, kw);
        Sk.resize((size_t)steps * kw);
        for (int l = 0; l < steps; ++l)
            for (int i = 0; i < kw; ++i) Sk[(size_t)l * kw + i] = S[(size_t)l * want + i];
        ESPER_CUDA(c, cudaMemcpyAsync(S_d, Sk.data(), sizeof(double) * Sk.size(), cudaMemcpyHostToDevice, st));
    }
    { LaunchScope ls(c, "lanczos"); ritz_kernel<KMAX><<<ceil_div(m, 128), 128, 0, st>>>(V, m, kw > 0 ? steps : 0, S_d, kw, U, k); }
    { LaunchScope ls(c, "lanczos"); fix_columns_kernel<<<k, 256, 0, st>>>(U, m, k); }
    ESPER_CUDA(c, cudaGetLastError());
    if (vec_h) ESPER_CUDA(c, cudaMemcpyAsync(vec_h, U, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost, st));
    ESPER_CUDA(c, cudaStreamSynchronize(st));
    if (val_h) {
        val_h[0] = 1.0;
        for (int i = 0; i < kw; ++i) val_h[1 + i] = theta[i] * theta[i];
    }
    return ESPER_OK;
}

}  // namespace esper
