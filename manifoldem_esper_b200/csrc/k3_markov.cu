// k3_markov.cu -- K3: Gaussian kernel, degree and symmetric normalisation,
//     A = exp(-D^2/(2 eps))                       1_Embedding/DM/DiffusionMaps.py:109
//     d_i = sum_j A_ij ; Ms = D^-1/2 (A D^-1) D^1/2  1_Embedding/DM/DiffusionMaps.py:131-152   (alpha = 0)
// The reference multiplies by dense diagonal matrices; each such product has one non-zero term, so
//     Ms_ij = (sqrt(1/d_i) * (A_ij * (1/d_j))) * sqrt(d_j)
// with every operation rounded to nearest reproduces its roundings.  Two passes over D (the second
// recomputes A instead of storing it): 2 reads + 1 write of m^2 x 8 B; HBM/L2-bandwidth bound.
#include "common.cuh"
#include "devutil.cuh"

namespace esper {
namespace markov {

__device__ __forceinline__ double kernel_value(double d, double two_eps) {
    return exp(-1.0 * __ddiv_rn(__dmul_rn(d, d), two_eps));
}

// One block per row: degree d_i = sum_j A_ij.
__global__ void __launch_bounds__(256) degree_kernel(const double* __restrict__ D, int m, double two_eps,
                                                     double* __restrict__ deg) {
    __shared__ double sh[32];
    const int i = blockIdx.x;
    const double* row = D + (size_t)i * m;
    double s = 0.0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) s += kernel_value(__ldg(row + j), two_eps);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) deg[i] = s;
}

// Rows [row0, row0 + gridDim.y) of Ms; D and Ms hold those rows only (row0 = 0 for the full matrix).
__global__ void __launch_bounds__(256) normalise_kernel(const double* __restrict__ D, int m, int row0, double two_eps,
                                                        const double* __restrict__ deg, double* __restrict__ Ms) {
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double di = deg[row0 + i], dj = __ldg(deg + j);
    const double a = kernel_value(__ldg(D + (size_t)i * m + j), two_eps);
    const double mij = __dmul_rn(a, __ddiv_rn(1.0, dj));                       // M = A @ inv(D)
    const double left = __dmul_rn(sqrt(__ddiv_rn(1.0, di)), mij);              // sqrtm(inv(D)) @ M
    Ms[(size_t)i * m + j] = __dmul_rn(left, sqrt(dj));                         // ... @ sqrtm(D)
}

}  // namespace markov

int run_markov(esper_ctx* c, int m, double eps) {
    using namespace markov;
    ESPER_RESERVE(c, c->markov, sizeof(double) * (size_t)m * m);
    ESPER_RESERVE(c, c->degree, sizeof(double) * (size_t)m);
    const double two_eps = 2.0 * eps;
    {
        LaunchScope ls(c, "markov");
        degree_kernel<<<m, 256, 0, c->stream>>>(c->dist.as<double>(), m, two_eps, c->degree.as<double>());
    }
    {
        LaunchScope ls(c, "markov");
        dim3 grid(ceil_div(m, 256), m);
        normalise_kernel<<<grid, 256, 0, c->stream>>>(c->dist.as<double>(), m, 0, two_eps, c->degree.as<double>(),
                                                      c->markov.as<double>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

// Row-block path.  phase 0: degrees of the owned rows -> degree_rows_h.  phase 1: Ms block from the degrees of
// ALL rows (gathered by the caller) -> resident c->markov [rows, m]; c->degree holds all m degrees.
int run_markov_rows(esper_ctx* c, int rows, int m, int row0, double eps, const double* degree_all_h, double* degree_rows_h, int phase) {
    using namespace markov;
    const double two_eps = 2.0 * eps;
    ESPER_RESERVE(c, c->degree, sizeof(double) * (size_t)(m + rows));
    double* deg_all = c->degree.as<double>();
    double* deg_rows = deg_all + m;
    if (phase == 0) {
        {
            LaunchScope ls(c, "markov");
            degree_kernel<<<rows, 256, 0, c->stream>>>(c->dist.as<double>(), m, two_eps, deg_rows);
        }
        ESPER_CUDA(c, cudaGetLastError());
        ESPER_CUDA(c, cudaMemcpyAsync(degree_rows_h, deg_rows, sizeof(double) * (size_t)rows, cudaMemcpyDeviceToHost, c->stream));
        ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
        return ESPER_OK;
    }
    ESPER_RESERVE(c, c->markov, sizeof(double) * (size_t)rows * m);
    ESPER_CUDA(c, cudaMemcpyAsync(deg_all, degree_all_h, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, "markov");
        dim3 grid(ceil_div(m, 256), rows);
        normalise_kernel<<<grid, 256, 0, c->stream>>>(c->dist.as<double>(), m, row0, two_eps, deg_all, c->markov.as<double>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

}  // namespace esper
