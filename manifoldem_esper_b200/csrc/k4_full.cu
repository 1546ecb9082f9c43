// k4_full.cu -- the WHOLE spectrum on request (k > 64, up to k = m): the reference saves the full U of its SVD
// (1_Embedding/DM/DiffusionMaps.py:166-169); its consumers read at most 16 columns, which is what the Lanczos solver of
// k4_eigen.cu / k4_chain.cu serves.  For callers that want every column:
//   1. Lanczos with full reorthogonalisation is run to exhaustion (m - 1 steps of the operator deflated by the locked top
//      pair): the basis V is then an orthonormal basis of the whole space and T = V^T Ms V holds the whole spectrum.
//   2. The host diagonalises the tridiagonal T with the implicit QL algorithm (EISPACK tql2 / Numerical Recipes tqli as
//      published), but instead of accumulating the eigenvector matrix it RECORDS the plane rotations (c, s) of every sweep:
//      O(n^2) scalar work.
//   3. The device applies the recorded rotations to the rows of V^T (row r of [q_1 .. q_n], kept in shared memory, one
//      thread per row, the rows of one CTA filling its SM): O(n^3) work in parallel, and the rotated rows ARE the rows of U.
// Output convention as esper_dmap_embed: val = lambda^2 descending, column 0 the steady state, unit columns, sign fixed.
#include "common.cuh"
#include "devutil.cuh"
#include <algorithm>
#include <numeric>

namespace esper {
namespace full {

// Zt[r][i] = V[1 + i][r]  (row-major [m][n], n = steps)
__global__ void __launch_bounds__(256) transpose_basis_kernel(const double* __restrict__ V, int m, int n, double* __restrict__ Zt) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int i = i0 + y, r = r0 + threadIdx.x;
        tile[y][threadIdx.x] = (i < n && r < m) ? V[(size_t)(1 + i) * m + r] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, i = i0 + threadIdx.x;
        if (r < m && i < n) Zt[(size_t)r * n + i] = tile[threadIdx.x][y];
    }
}

// One thread per row of Zt; the rows of a CTA live in shared memory for the whole rotation sequence.
// Segment q: rotations on the column pairs (hi, hi+1), (hi-1, hi), ... (hi-len+1, hi-len+2) in that order, with
//     f = z[i+1];  z[i+1] = s z[i] + c f;  z[i] = c z[i] - s f          (tql2's accumulation step)
__global__ void __launch_bounds__(32) apply_rotations_kernel(double* __restrict__ Zt, int m, int n, int rows_per_cta,
                                                             const int2* __restrict__ segs, int n_segs,
                                                             const long long* __restrict__ seg_off, const double2* __restrict__ cs) {
    extern __shared__ double s_rows[];                    // [rows_per_cta][n]
    const int row0 = blockIdx.x * rows_per_cta;
    const int n_rows = min(rows_per_cta, m - row0);
    for (int idx = threadIdx.x; idx < n_rows * n; idx += blockDim.x) s_rows[idx] = Zt[(size_t)row0 * n + idx];
    __syncwarp();
    if ((int)threadIdx.x < n_rows) {
        double* z = s_rows + (size_t)threadIdx.x * n;
        for (int q = 0; q < n_segs; ++q) {
            const int2 sg = __ldg(segs + q);
            const double2* rot = cs + __ldg(seg_off + q);
            int i = sg.x;
            double zi1 = z[i + 1];
            int t = 0;
            for (; t + 4 <= sg.y; t += 4) {               // four rotations' coefficients in flight
                const double2 r0 = __ldg(rot + t), r1 = __ldg(rot + t + 1), r2 = __ldg(rot + t + 2), r3 = __ldg(rot + t + 3);
                double zi = z[i];     z[i + 1] = fma(r0.y, zi, r0.x * zi1); zi1 = fma(r0.x, zi, -(r0.y * zi1)); --i;
                zi = z[i];            z[i + 1] = fma(r1.y, zi, r1.x * zi1); zi1 = fma(r1.x, zi, -(r1.y * zi1)); --i;
                zi = z[i];            z[i + 1] = fma(r2.y, zi, r2.x * zi1); zi1 = fma(r2.x, zi, -(r2.y * zi1)); --i;
                zi = z[i];            z[i + 1] = fma(r3.y, zi, r3.x * zi1); zi1 = fma(r3.x, zi, -(r3.y * zi1)); --i;
            }
            for (; t < sg.y; ++t) {
                const double2 r0 = __ldg(rot + t);
                const double zi = z[i];
                z[i + 1] = fma(r0.y, zi, r0.x * zi1); zi1 = fma(r0.x, zi, -(r0.y * zi1)); --i;
            }
            z[i + 1] = zi1;
        }
    }
    __syncwarp();
    for (int idx = threadIdx.x; idx < n_rows * n; idx += blockDim.x) Zt[(size_t)row0 * n + idx] = s_rows[idx];
}

// U[r][0] = V_0[r];  U[r][1 + q] = Zt[r][perm[q]]   for q < k - 1
__global__ void __launch_bounds__(256) gather_columns_kernel(const double* __restrict__ V, const double* __restrict__ Zt, int m, int n,
                                                             const int* __restrict__ perm, int k, double* __restrict__ U) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * k) return;
    const int r = (int)(idx / k), q = (int)(idx % k);
    U[idx] = (q == 0) ? V[r] : Zt[(size_t)r * n + perm[q - 1]];
}

}  // namespace full

// Implicit QL on T = tridiag(e, d, e) (d[0..n), e[0..n-1) sub-diagonal; both overwritten): d becomes the eigenvalues (unsorted);
// every sweep's rotations are appended to (segs, cs) in the order tql2 would apply them to its eigenvector matrix.
// Returns 0, or the index of the eigenvalue that did not converge in 60 sweeps + 1.
int tridiag_ql_rotations(std::vector<double>& d, std::vector<double>& e, int n, std::vector<int2>& segs, std::vector<long long>& seg_off,
                         std::vector<double2>& cs) {
    segs.clear(); seg_off.clear(); cs.clear();
    if (n < 2) return 0;
    e.resize(n);
    e[n - 1] = 0.0;
    auto sign = [](double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); };
    for (int l = 0; l < n; ++l) {
        int iter = 0, mm;
        do {
            for (mm = l; mm < n - 1; ++mm) {
                const double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= 2.220446049250313e-16 * dd) break;
            }
            if (mm != l) {
                if (iter++ == 60) return l + 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + sign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                const long long off = (long long)cs.size();
                int i;
                for (i = mm - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[mm] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    cs.push_back(make_double2(c, s));
                }
                const int len = (int)((long long)cs.size() - off);
                if (len > 0) { segs.push_back(make_int2(mm - 1, len)); seg_off.push_back(off); }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[mm] = 0.0;
            }
        } while (mm != l);
    }
    return 0;
}

// Host-only reference application of the recorded rotations to the identity: Z [n][n] row-major = eigenvectors of T in the
// columns (used by esper_tridiag_eigen_all and its CPU test; the device applies the same list to the rows of V^T).
void apply_rotations_host(int n, const std::vector<int2>& segs, const std::vector<long long>& seg_off, const std::vector<double2>& cs,
                          double* Z) {
    for (size_t q = 0; q < segs.size(); ++q) {
        const double2* rot = cs.data() + seg_off[q];
        for (int k = 0; k < n; ++k) {
            double* z = Z + (size_t)k * n;
            int i = segs[q].x;
            for (int t = 0; t < segs[q].y; ++t, --i) {
                const double f = z[i + 1];
                z[i + 1] = rot[t].y * z[i] + rot[t].x * f;
                z[i] = rot[t].x * z[i] - rot[t].y * f;
            }
        }
    }
}

// Host only: all eigenpairs of T = tridiag(beta, alpha, beta) of order n through the recorded-rotation path: theta [n] descending,
// Z [n][n] row-major with the eigenvectors in the columns (same order).  Returns 0 or ESPER_E_NOCONV.
int tridiag_eigen_all_host(const double* alpha, const double* beta, int n, double* theta, double* Z) {
    std::vector<double> d(alpha, alpha + n), e(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) e[i] = beta[i];
    std::vector<int2> segs; std::vector<long long> seg_off; std::vector<double2> cs;
    if (tridiag_ql_rotations(d, e, n, segs, seg_off, cs)) return ESPER_E_NOCONV;
    std::vector<double> I((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) I[(size_t)i * n + i] = 1.0;
    apply_rotations_host(n, segs, seg_off, cs, I.data());
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return d[x] > d[y]; });
    for (int q = 0; q < n; ++q) {
        theta[q] = d[perm[q]];
        for (int r = 0; r < n; ++r) Z[(size_t)r * n + q] = I[(size_t)r * n + perm[q]];
    }
    return ESPER_OK;
}

// After a Lanczos run to exhaustion (`steps` vectors q_1..q_steps in V[1..], alpha / beta on the host): all eigenpairs.
// U_out: device [m][k] (column 0 = V[0], column 1 + q = eigenvector of the q-th largest eigenvalue of T, not yet normalised);
// theta_out: the k - 1 leading eigenvalues of T, descending.
int finish_full_spectrum(esper_ctx* c, int m, int k, int steps, const std::vector<double>& a, const std::vector<double>& b, double* V,
                         double** U_out, std::vector<double>& theta_out) {
    using namespace full;
    cudaStream_t st = c->stream;
    const int n = steps;
    if (n + 1 < k) return fail(c, ESPER_E_NOCONV, "full spectrum: the Krylov space ended after %d of %d steps (invariant subspace); %d pairs available, %d wanted", n, m - 1, n + 1, k);
    std::vector<double> d(a.begin(), a.begin() + n), e(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) e[i] = b[i];
    std::vector<int2> segs; std::vector<long long> seg_off; std::vector<double2> cs;
    const int bad = tridiag_ql_rotations(d, e, n, segs, seg_off, cs);
    if (bad) return fail(c, ESPER_E_NOCONV, "full spectrum: QL iteration did not converge for eigenvalue %d", bad - 1);
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return d[x] > d[y]; });
    // device buffers: Zt [m][n] | U [m][k] | rotations | segments
    const size_t zt_b = sizeof(double) * (size_t)m * n, u_b = sizeof(double) * (size_t)m * k;
    const size_t cs_b = sizeof(double2) * std::max<size_t>(cs.size(), 1), sg_b = sizeof(int2) * std::max<size_t>(segs.size(), 1);
    const size_t so_b = sizeof(long long) * std::max<size_t>(segs.size(), 1), pm_b = sizeof(int) * (size_t)n;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    ESPER_RESERVE(c, c->lan_small, up(zt_b) + up(u_b) + up(cs_b) + up(sg_b) + up(so_b) + up(pm_b) + 256);
    char* base = c->lan_small.as<char>();
    double* Zt = reinterpret_cast<double*>(base);
    double* U = reinterpret_cast<double*>(base + up(zt_b));
    double2* cs_d = reinterpret_cast<double2*>(base + up(zt_b) + up(u_b));
    int2* sg_d = reinterpret_cast<int2*>(base + up(zt_b) + up(u_b) + up(cs_b));
    long long* so_d = reinterpret_cast<long long*>(base + up(zt_b) + up(u_b) + up(cs_b) + up(sg_b));
    int* pm_d = reinterpret_cast<int*>(base + up(zt_b) + up(u_b) + up(cs_b) + up(sg_b) + up(so_b));
    if (!cs.empty()) {
        ESPER_CUDA(c, cudaMemcpyAsync(cs_d, cs.data(), sizeof(double2) * cs.size(), cudaMemcpyHostToDevice, st));
        ESPER_CUDA(c, cudaMemcpyAsync(sg_d, segs.data(), sizeof(int2) * segs.size(), cudaMemcpyHostToDevice, st));
        ESPER_CUDA(c, cudaMemcpyAsync(so_d, seg_off.data(), sizeof(long long) * segs.size(), cudaMemcpyHostToDevice, st));
    }
    ESPER_CUDA(c, cudaMemcpyAsync(pm_d, perm.data(), pm_b, cudaMemcpyHostToDevice, st));
    { LaunchScope ls(c, "lanczos"); transpose_basis_kernel<<<dim3(ceil_div(n, 32), ceil_div(m, 32)), dim3(32, 8), 0, st>>>(V, m, n, Zt); }
    if (!cs.empty()) {
        int optin = 0;
        ESPER_CUDA(c, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        int rows = std::max(1, std::min(32, std::min(ceil_div(m, c->sm_count), (int)((size_t)(optin - 1024) / (sizeof(double) * (size_t)n)))));
        const size_t smem = sizeof(double) * (size_t)rows * n;
        if (smem > (size_t)optin) return fail(c, ESPER_E_ARG, "full spectrum: a row of %d doubles does not fit in shared memory", n);
        ESPER_CUDA(c, cudaFuncSetAttribute(apply_rotations_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        LaunchScope ls(c, "lanczos");
        apply_rotations_kernel<<<ceil_div(m, rows), 32, smem, st>>>(Zt, m, n, rows, sg_d, (int)segs.size(), so_d, cs_d);
    }
    { LaunchScope ls(c, "lanczos"); gather_columns_kernel<<<ceil_div((long long)m * k, 256), 256, 0, st>>>(V, Zt, m, n, pm_d, k, U); }
    ESPER_CUDA(c, cudaGetLastError());
    *U_out = U;
    theta_out.resize((size_t)k - 1);
    for (int q = 1; q < k; ++q) theta_out[(size_t)q - 1] = d[perm[q - 1]];
    return ESPER_OK;
}

}  // namespace esper
