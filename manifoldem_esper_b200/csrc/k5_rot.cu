// k5_rot.cu -- K5: the Nd eigen-rotation + 2-D histogram sweep of
// 2_Manifold_Rotations/Rotation_Histograms.py:283-365 with Generate_Nd_Rot.py:4-45.
//
// One CTA per evaluation (PD, subspace, theta): two threads build rows v1, v2 of the composite
// rotation R = G_0 G_1 ... G_{T-1} by T Givens column updates, left to right like the reference's
// chain of `dot`s and with the accumulation order of a dgemm micro-kernel (first product rounded,
// later terms fused), which reproduces numpy/OpenBLAS bit for bit; every thread then rotates points
// (x, y) = (R[v1,:].u, R[v2,:].u) in FP64, bins them with numpy.histogramdd's rule (edges from
// linspace, right edge inclusive, outliers dropped) into a shared-memory privatised 51x51 histogram,
// and the block counts the non-zero bins.  The sequential part of the search (each Rij keeps the
// previous arg-max) runs as n_rij dependent launch pairs (evaluate all, select first arg-max).
#include "common.cuh"
#include "devutil.cuh"

namespace esper {

void hist_edges_host(int bins, double lo, double hi, double* edges) {
    // numpy.linspace(lo, hi, bins+1): step = (hi-lo)/bins; y = arange(0, bins+1)*step + lo; y[-1] = hi
    const double step = (hi - lo) / (double)bins;
    for (int i = 0; i <= bins; ++i) {
        volatile double t = (double)i * step;      // keep the product rounded before the add (no FMA)
        edges[i] = t + lo;
    }
    edges[bins] = hi;
}

void nd_rotation_host(int n, const double* theta, double* R) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) R[i * n + j] = (i == j) ? 1.0 : 0.0;
    int r = 0;
    for (int k = 0; k < n - 1; ++k)
        for (int j = k + 1; j < n; ++j, ++r) {
            const double cs = cos(theta[r]), sn = sin(theta[r]);
            const double nsn = -1.0 * sn;
            for (int i = 0; i < n; ++i) {
                // dot-product order of a dgemm micro-kernel: k ascending, one accumulator, fused multiply-add
                const double a = R[i * n + k], b = R[i * n + j];
                volatile double p1 = a * cs, p3 = a * nsn;
                R[i * n + k] = fma(b, sn, p1);
                R[i * n + j] = fma(b, cs, p3);
            }
        }
}

namespace rot {

constexpr int MAX_DIM = 16;
constexpr int MAX_BINS = 64;

__device__ __forceinline__ int find_bin(double x, const double* __restrict__ edges, int bins, double inv_w, double lo) {
    // numpy: idx = searchsorted(edges, x, 'right'); x == edges[-1] -> last bin; idx in [1, bins] kept.
    if (!(x >= edges[0] && x <= edges[bins])) return -1;       // outlier or NaN
    int b = (int)((x - lo) * inv_w);
    b = b < 0 ? 0 : (b > bins - 1 ? bins - 1 : b);
    while (b > 0 && x < edges[b]) --b;
    while (b < bins - 1 && x >= edges[b + 1]) ++b;
    return b;
}

// Rotate all m points of one PD by rows (rx, ry) of R, histogram, count non-zero bins.
__device__ __forceinline__ int rotate_hist_count(const double* __restrict__ U, int m, int dim,
                                                 const double* rx, const double* ry, const double* edges, int bins,
                                                 double inv_w, double lo, unsigned int* hist, int* red,
                                                 int* __restrict__ H_out) {
    const int nb2 = bins * bins;
    for (int t = threadIdx.x; t < nb2; t += blockDim.x) hist[t] = 0u;
    __syncthreads();
    for (int p = threadIdx.x; p < m; p += blockDim.x) {
        const double* u = U + (size_t)p * dim;
        double x = __dmul_rn(rx[0], u[0]), y = __dmul_rn(ry[0], u[0]);
        for (int cidx = 1; cidx < dim; ++cidx) {         // same accumulation order as the BLAS product
            const double uc = u[cidx];
            x = __fma_rn(rx[cidx], uc, x);
            y = __fma_rn(ry[cidx], uc, y);
        }
        const int bx = find_bin(x, edges, bins, inv_w, lo);
        const int by = find_bin(y, edges, bins, inv_w, lo);
        if (bx >= 0 && by >= 0) atomicAdd(&hist[bx * bins + by], 1u);
    }
    __syncthreads();
    int cnt = 0;
    for (int t = threadIdx.x; t < nb2; t += blockDim.x) {
        const unsigned int hcount = hist[t];
        cnt += (hcount != 0u);
        if (H_out) H_out[t] = (int)hcount;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    int total = 0;
    if (threadIdx.x == 0) for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) total += red[w2];
    return total;   // valid in thread 0
}

// Batch of independent evaluations with caller-supplied rotation matrices.
__global__ void __launch_bounds__(256)
eval_kernel(const double* __restrict__ Uall, int m, int dim, const double* __restrict__ R,
            const int* __restrict__ pd_of_eval, const int* __restrict__ v1, const int* __restrict__ v2,
            const double* __restrict__ edges_g, int bins, double lo, double hi,
            int* __restrict__ nonzero, int* __restrict__ H) {
    __shared__ unsigned int hist[MAX_BINS * MAX_BINS];
    __shared__ double edges[MAX_BINS + 1];
    __shared__ double rx[MAX_DIM], ry[MAX_DIM];
    __shared__ int red[8];
    const int e = blockIdx.x;
    if (threadIdx.x <= bins) edges[threadIdx.x] = edges_g[threadIdx.x];
    if (threadIdx.x < dim) {
        rx[threadIdx.x] = R[(size_t)e * dim * dim + (size_t)v1[e] * dim + threadIdx.x];
        ry[threadIdx.x] = R[(size_t)e * dim * dim + (size_t)v2[e] * dim + threadIdx.x];
    }
    __syncthreads();
    const double inv_w = (double)bins / (hi - lo);
    const int cnt = rotate_hist_count(Uall + (size_t)pd_of_eval[e] * m * dim, m, dim, rx, ry, edges, bins, inv_w, lo,
                                      hist, red, H ? H + (size_t)e * bins * bins : nullptr);
    if (threadIdx.x == 0) nonzero[e] = cnt;
}

// One stage of the sequential search: grid (n_theta, max_sub, n_pd).
//   sel: [n_pd, max_sub, T] index into the theta grid chosen so far for each plane, -1 = angle 0.
__global__ void __launch_bounds__(256)
sweep_eval_kernel(const double* __restrict__ Uall, int m, int dim, int T, const int* __restrict__ combo,
                  const int* __restrict__ n_sub, int max_sub, const int* __restrict__ rij, int n_rij, int stage,
                  const double* __restrict__ cos_g, const double* __restrict__ sin_g, int n_theta,
                  const int* __restrict__ sel, const double* __restrict__ edges_g, int bins, double lo, double hi,
                  int* __restrict__ counts /*[n_pd, max_sub, n_rij, n_theta]*/) {
    __shared__ unsigned int hist[MAX_BINS * MAX_BINS];
    __shared__ double edges[MAX_BINS + 1];
    __shared__ double rows[2][MAX_DIM];
    __shared__ int red[8];
    const int ti = blockIdx.x, sub = blockIdx.y, pd = blockIdx.z;
    if (sub >= n_sub[pd]) return;
    const int va = combo[((size_t)pd * max_sub + sub) * 2], vb = combo[((size_t)pd * max_sub + sub) * 2 + 1];
    const int plane_now = rij[(size_t)pd * n_rij + stage];
    if (threadIdx.x <= bins) edges[threadIdx.x] = edges_g[threadIdx.x];
    if (threadIdx.x < 2) {
        // Row `v` of R: start from e_v and apply the T Givens column updates in lexicographic plane order.
        const int v = threadIdx.x == 0 ? va : vb;
        double row[MAX_DIM];
        for (int q = 0; q < dim; ++q) row[q] = (q == v) ? 1.0 : 0.0;
        const int* s = sel + ((size_t)pd * max_sub + sub) * T;
        int r = 0;
        for (int a = 0; a < dim - 1; ++a)
            for (int b = a + 1; b < dim; ++b, ++r) {
                const int gi = (r == plane_now) ? ti : s[r];
                if (gi < 0) continue;                                   // angle 0: G = I, exact no-op
                const double cs = cos_g[gi], sn = sin_g[gi];
                const double nsn = __dmul_rn(-1.0, sn);
                const double ra = row[a], rb = row[b];
                row[a] = __fma_rn(rb, sn, __dmul_rn(ra, cs));
                row[b] = __fma_rn(rb, cs, __dmul_rn(ra, nsn));
            }
        for (int q = 0; q < dim; ++q) rows[threadIdx.x][q] = row[q];
    }
    __syncthreads();
    const double inv_w = (double)bins / (hi - lo);
    const int cnt = rotate_hist_count(Uall + (size_t)pd * m * dim, m, dim, rows[0], rows[1], edges, bins, inv_w, lo,
                                      hist, red, nullptr);
    if (threadIdx.x == 0) counts[(((size_t)pd * max_sub + sub) * n_rij + stage) * n_theta + ti] = cnt;
}

// First arg-max of the empty-bin score == first arg-min of the non-zero-bin count.
__global__ void sweep_select_kernel(const int* __restrict__ counts, const int* __restrict__ n_sub, int max_sub,
                                    const int* __restrict__ rij, int n_rij, int stage, int n_theta, int T,
                                    int* __restrict__ sel) {
    const int sub = blockIdx.x, pd = blockIdx.y;
    if (sub >= n_sub[pd] || threadIdx.x != 0) return;
    const int* cnt = counts + (((size_t)pd * max_sub + sub) * n_rij + stage) * n_theta;
    int best = 0;
    for (int t = 1; t < n_theta; ++t) if (cnt[t] < cnt[best]) best = t;
    sel[((size_t)pd * max_sub + sub) * T + rij[(size_t)pd * n_rij + stage]] = best;
}

__global__ void sweep_output_kernel(const int* __restrict__ sel, const double* __restrict__ theta_grid, int total,
                                    double* __restrict__ thetas) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) thetas[i] = sel[i] >= 0 ? theta_grid[sel[i]] : 0.0;
}

}  // namespace rot

static int upload(esper_ctx* c, DevBuf& buf, const void* src, size_t bytes) {
    ESPER_RESERVE(c, buf, bytes);
    ESPER_CUDA(c, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return ESPER_OK;
}

int run_rot_eval(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, const double* R,
                 const int* pd_of_eval, const int* v1, const int* v2, int n_eval, int bins,
                 double lo, double hi, int* nonzero, int* H) {
    using namespace rot;
    if (dim < 2 || dim > MAX_DIM || bins < 1 || bins > MAX_BINS || !(hi > lo))
        return fail(c, ESPER_E_ARG, "rot_hist_eval: need 2<=dim<=%d, 1<=bins<=%d, hi>lo", MAX_DIM, MAX_BINS);
    for (int e = 0; e < n_eval; ++e)
        if (pd_of_eval[e] < 0 || pd_of_eval[e] >= n_pd || v1[e] < 0 || v1[e] >= dim || v2[e] < 0 || v2[e] >= dim)
            return fail(c, ESPER_E_ARG, "rot_hist_eval: evaluation %d indexes out of range", e);
    std::vector<double> edges(bins + 1);
    hist_edges_host(bins, lo, hi, edges.data());
    int rc;
    c->rot_n_pd = 0;                                   // rot_U no longer holds esper_manifold_prepare's output
    if ((rc = upload(c, c->rot_U, Uinit, sizeof(double) * (size_t)n_pd * m * dim))) return rc;
    if ((rc = upload(c, c->rot_R, R, sizeof(double) * (size_t)n_eval * dim * dim))) return rc;
    std::vector<int> idx((size_t)3 * n_eval);
    memcpy(idx.data(), pd_of_eval, sizeof(int) * n_eval);
    memcpy(idx.data() + n_eval, v1, sizeof(int) * n_eval);
    memcpy(idx.data() + 2 * (size_t)n_eval, v2, sizeof(int) * n_eval);
    if ((rc = upload(c, c->rot_idx, idx.data(), sizeof(int) * idx.size()))) return rc;
    if ((rc = upload(c, c->rot_theta, edges.data(), sizeof(double) * edges.size()))) return rc;
    ESPER_RESERVE(c, c->rot_out, sizeof(int) * (size_t)n_eval);
    int* H_d = nullptr;
    if (H) { ESPER_RESERVE(c, c->rot_H, sizeof(int) * (size_t)n_eval * bins * bins); H_d = c->rot_H.as<int>(); }
    const int* idx_d = c->rot_idx.as<int>();
    {
        LaunchScope ls(c, "rot");
        eval_kernel<<<n_eval, 256, 0, c->stream>>>(c->rot_U.as<double>(), m, dim, c->rot_R.as<double>(), idx_d,
                                                   idx_d + n_eval, idx_d + 2 * (size_t)n_eval, c->rot_theta.as<double>(),
                                                   bins, lo, hi, c->rot_out.as<int>(), H_d);
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(nonzero, c->rot_out.p, sizeof(int) * (size_t)n_eval, cudaMemcpyDeviceToHost, c->stream));
    if (H) ESPER_CUDA(c, cudaMemcpyAsync(H, H_d, sizeof(int) * (size_t)n_eval * bins * bins, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

int run_rot_sweep(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, const int* combo,
                  const int* n_sub, int max_sub, const int* rij, int n_rij, const double* theta_grid,
                  int n_theta, int bins, double lo, double hi, double* thetas_out, int* counts_out) {
    using namespace rot;
    if (dim < 2 || dim > MAX_DIM || bins < 1 || bins > MAX_BINS || !(hi > lo) || n_theta < 1 || max_sub < 1 || n_rij < 0)
        return fail(c, ESPER_E_ARG, "rot_sweep: need 2<=dim<=%d, 1<=bins<=%d, hi>lo, n_theta>=1", MAX_DIM, MAX_BINS);
    if (n_pd > 65535 || max_sub > 65535) return fail(c, ESPER_E_ARG, "rot_sweep: too many PDs / subspaces per call");
    const int T = dim * (dim - 1) / 2;
    for (int p = 0; p < n_pd; ++p) {
        if (n_sub[p] < 0 || n_sub[p] > max_sub) return fail(c, ESPER_E_ARG, "rot_sweep: n_sub[%d] out of range", p);
        for (int s = 0; s < n_sub[p]; ++s)
            for (int q = 0; q < 2; ++q) {
                const int v = combo[((size_t)p * max_sub + s) * 2 + q];
                if (v < 0 || v >= dim) return fail(c, ESPER_E_ARG, "rot_sweep: combo[%d,%d] out of range", p, s);
            }
        for (int s = 0; s < n_rij; ++s)
            if (rij[(size_t)p * n_rij + s] < 0 || rij[(size_t)p * n_rij + s] >= T)
                return fail(c, ESPER_E_ARG, "rot_sweep: rij[%d,%d] out of range", p, s);
    }
    // host tables: histogram edges, cos/sin of the grid with the C library (as the reference's math.cos/sin)
    std::vector<double> tab((size_t)(bins + 1) + 3 * (size_t)n_theta);
    hist_edges_host(bins, lo, hi, tab.data());
    double* cg = tab.data() + bins + 1; double* sg = cg + n_theta; double* tg = sg + n_theta;
    for (int t = 0; t < n_theta; ++t) { cg[t] = cos(theta_grid[t]); sg[t] = sin(theta_grid[t]); tg[t] = theta_grid[t]; }
    int rc;
    if (Uinit) {                                       // NULL: manifolds left resident by esper_manifold_prepare
        c->rot_n_pd = 0;
        if ((rc = upload(c, c->rot_U, Uinit, sizeof(double) * (size_t)n_pd * m * dim))) return rc;
        c->rot_n_pd = n_pd; c->rot_m = m; c->rot_dim = dim;
    }
    if ((rc = upload(c, c->rot_theta, tab.data(), sizeof(double) * tab.size()))) return rc;
    std::vector<int> idx((size_t)n_pd * max_sub * 2 + n_pd + (size_t)n_pd * (n_rij > 0 ? n_rij : 1));
    memcpy(idx.data(), combo, sizeof(int) * (size_t)n_pd * max_sub * 2);
    memcpy(idx.data() + (size_t)n_pd * max_sub * 2, n_sub, sizeof(int) * n_pd);
    if (n_rij > 0) memcpy(idx.data() + (size_t)n_pd * max_sub * 2 + n_pd, rij, sizeof(int) * (size_t)n_pd * n_rij);
    if ((rc = upload(c, c->rot_idx, idx.data(), sizeof(int) * idx.size()))) return rc;
    const int* combo_d = c->rot_idx.as<int>();
    const int* nsub_d = combo_d + (size_t)n_pd * max_sub * 2;
    const int* rij_d = nsub_d + n_pd;
    const size_t n_sel = (size_t)n_pd * max_sub * T;
    const size_t n_cnt = (size_t)n_pd * max_sub * (n_rij > 0 ? n_rij : 1) * n_theta;
    ESPER_RESERVE(c, c->rot_out, sizeof(int) * (n_sel + n_cnt));
    ESPER_RESERVE(c, c->rot_aux, sizeof(double) * n_sel);
    int* sel_d = c->rot_out.as<int>();
    int* cnt_d = sel_d + n_sel;
    ESPER_CUDA(c, cudaMemsetAsync(sel_d, 0xFF, sizeof(int) * n_sel, c->stream));        // -1 = angle 0
    ESPER_CUDA(c, cudaMemsetAsync(cnt_d, 0, sizeof(int) * n_cnt, c->stream));
    const double* edges_d = c->rot_theta.as<double>();
    const double* cos_d = edges_d + bins + 1; const double* sin_d = cos_d + n_theta; const double* grid_d = sin_d + n_theta;
    for (int stage = 0; stage < n_rij; ++stage) {
        {
            LaunchScope ls(c, "rot");
            dim3 grid(n_theta, max_sub, n_pd);
            sweep_eval_kernel<<<grid, 256, 0, c->stream>>>(c->rot_U.as<double>(), m, dim, T, combo_d, nsub_d, max_sub,
                                                           rij_d, n_rij, stage, cos_d, sin_d, n_theta, sel_d, edges_d,
                                                           bins, lo, hi, cnt_d);
        }
        {
            LaunchScope ls(c, "rot");
            dim3 grid(max_sub, n_pd);
            sweep_select_kernel<<<grid, 32, 0, c->stream>>>(cnt_d, nsub_d, max_sub, rij_d, n_rij, stage, n_theta, T, sel_d);
        }
    }
    {
        LaunchScope ls(c, "rot");
        sweep_output_kernel<<<ceil_div((long long)n_sel, 256), 256, 0, c->stream>>>(sel_d, grid_d, (int)n_sel,
                                                                                    c->rot_aux.as<double>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(thetas_out, c->rot_aux.p, sizeof(double) * n_sel, cudaMemcpyDeviceToHost, c->stream));
    if (counts_out && n_rij > 0)
        ESPER_CUDA(c, cudaMemcpyAsync(counts_out, cnt_d, sizeof(int) * n_cnt, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    // rows of unused subspaces are zero (sel = -1 -> 0.0) as documented
    return ESPER_OK;
}

}  // namespace esper
