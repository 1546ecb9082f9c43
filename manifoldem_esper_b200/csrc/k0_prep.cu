// k0_prep.cu -- K0: per-image standardisation (Distances.py:88-92), PD mean-image centring and the
// TF32 hi/lo operand split that feeds the tcgen05 Gram kernel.  HBM-bound streaming passes.
//
//   pass 1  row_stats+col_sum : on a SUBSET of <= 256 images: mean32/std32, then the column mean of their
//                               standardised pixels -> xbar32[P]  (ESPER_DIST_CENTER).  Distances are
//                               invariant to subtracting ANY common image, so a subset mean serves.
//   pass 2  stats_split       : per image (one CTA): sum / sum-of-squares (FP64) -> mean32, std32, then a
//                               second sweep of the same row (L2-resident, 4P bytes per CTA):
//                               c = fl32(z - xbar); hi = tf32(c); lo = tf32(c - hi); |c|^2 in FP64
//
// Algorithmic bytes per PD (n images, P pixels): read 4nP once from HBM (+4nP from L2), write 8nP.
#include "common.cuh"
#include "devutil.cuh"

namespace esper {

__device__ __forceinline__ void row_stats(const float* __restrict__ row, int P, double* sh, float& mean32, float& std32) {
    double s = 0.0, s2 = 0.0;
    if ((P & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        int n4 = P >> 2;
        for (int q = threadIdx.x; q < n4; q += blockDim.x) {
            float4 v = __ldg(r4 + q);
            double a = v.x, b = v.y, c = v.z, d = v.w;
            s += (a + b) + (c + d);
            s2 += (a * a + b * b) + (c * c + d * d);
        }
    } else {
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            double a = __ldg(row + p);
            s += a; s2 += a * a;
        }
    }
    __shared__ double bc[2];
    s = block_sum(s, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        double mean = s / P;
        double var = s2 / P - mean * mean;
        if (var < 0.0) var = 0.0;
        bc[0] = (double)(float)mean;
        bc[1] = (double)(float)sqrt(var);
    }
    __syncthreads();
    mean32 = (float)bc[0];
    std32 = (float)bc[1];
    __syncthreads();
}

// Statistics of the images i = blockIdx.x * stride (the centring subset).
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ raw, int n, int P, int stride,
                                                        int standardize, RowStat* __restrict__ st) {
    __shared__ double sh[32];
    int i = blockIdx.x * stride;
    if (i >= n) return;
    if (!standardize) {
        if (threadIdx.x == 0) { st[i].mean = 0.f; st[i].stdv = 1.f; st[i].norm2 = 0.0; }
        return;
    }
    float m32, s32;
    row_stats(raw + (size_t)i * P, P, sh, m32, s32);
    if (threadIdx.x == 0) { st[i].mean = m32; st[i].stdv = s32; st[i].norm2 = 0.0; }
}

// Partial column sums of the standardised images: part[g][p] = sum_{i in group g} z_ip  (FP64).
__global__ void __launch_bounds__(256) col_sum_kernel(const float* __restrict__ raw, int n_sub, int stride, int P,
                                                      const RowStat* __restrict__ st, int groups,
                                                      double* __restrict__ part) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int g = blockIdx.y;
    int k0 = (int)((long long)n_sub * g / groups), k1 = (int)((long long)n_sub * (g + 1) / groups);
    if (p >= P) return;
    double acc = 0.0;
    for (int k = k0; k < k1; ++k) {
        const int i = k * stride;
        float m = st[i].mean, sd = st[i].stdv;
        acc += (double)zscore(__ldg(raw + (size_t)i * P + p), m, sd);
    }
    part[(size_t)g * P + p] = acc;
}

__global__ void __launch_bounds__(256) col_mean_kernel(const double* __restrict__ part, int groups, int P,
                                                       int n, float* __restrict__ xbar) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double acc = 0.0;
    for (int g = 0; g < groups; ++g) acc += part[(size_t)g * P + p];
    xbar[p] = (float)(acc / n);
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// One block per image: write the hi / lo TF32 planes (row stride ld, zero padded) and |c|^2.
__global__ void __launch_bounds__(512) stats_split_kernel(const float* __restrict__ raw, int n, int P, int ld,
                                                          int standardize, RowStat* __restrict__ st,
                                                          const float* __restrict__ xbar, int center,
                                                          float* __restrict__ hi, float* __restrict__ lo) {
    __shared__ double sh[32];
    int i = blockIdx.x;
    const float* row = raw + (size_t)i * P;
    float4* hrow = reinterpret_cast<float4*>(hi + (size_t)i * ld);
    float4* lrow = reinterpret_cast<float4*>(lo + (size_t)i * ld);
    float m = 0.f, sd = 1.f;
    if (standardize) row_stats(row, P, sh, m, sd);
    const bool vec = ((P & 3) == 0);
    double acc = 0.0;
    int nq = ld >> 2;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        int p0 = q << 2;
        float v[4];
        if (vec && p0 + 3 < P) {
            float4 t = __ldg(reinterpret_cast<const float4*>(row) + q);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (p0 + e < P) ? __ldg(row + p0 + e) : 0.f;
        }
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float c = 0.f;
            if (p0 + e < P) {
                c = zscore(v[e], m, sd);
                if (center) c = __fsub_rn(c, __ldg(xbar + p0 + e));
            }
            h[e] = tf32_rna(c);
            l[e] = tf32_rna(__fsub_rn(c, h[e]));
            acc += (double)c * (double)c;
        }
        hrow[q] = make_float4(h[0], h[1], h[2], h[3]);
        lrow[q] = make_float4(l[0], l[1], l[2], l[3]);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) { st[i].mean = m; st[i].stdv = sd; st[i].norm2 = acc; }
}

int run_prep(esper_ctx* c, int n, int P, unsigned flags, int rows_pad, int ld, bool exact_center) {
    const float* raw = c->imgs.as<float>();
    ESPER_RESERVE(c, c->row_stats, sizeof(RowStat) * (size_t)rows_pad);
    ESPER_RESERVE(c, c->plane_hi, sizeof(float) * (size_t)rows_pad * ld);
    ESPER_RESERVE(c, c->plane_lo, sizeof(float) * (size_t)rows_pad * ld);
    RowStat* st = c->row_stats.as<RowStat>();
    const int standardize = (flags & ESPER_DIST_STANDARDIZE) ? 1 : 0;
    const int center = (flags & ESPER_DIST_CENTER) ? 1 : 0;
    float* xbar = nullptr;
    if (center) {
        // distances: any common image may be subtracted, a subset mean (images 0, stride, 2*stride, ...) serves;
        // PCA (PCA.py:53-55) needs the mean over all images
        const int stride = (!exact_center && n > 256) ? n / 256 : 1;
        const int n_sub = (n + stride - 1) / stride;
        const int groups = n_sub >= 16 ? 16 : 1;
        ESPER_RESERVE(c, c->col_part, sizeof(double) * (size_t)(groups + 1) * P);
        double* part = c->col_part.as<double>();
        xbar = reinterpret_cast<float*>(part + (size_t)groups * P);
        {
            LaunchScope ls(c, "prep");
            row_stats_kernel<<<n_sub, 256, 0, c->stream>>>(raw, n, P, stride, standardize, st);
        }
        {
            LaunchScope ls(c, "prep");
            dim3 grid(ceil_div(P, 256), groups);
            col_sum_kernel<<<grid, 256, 0, c->stream>>>(raw, n_sub, stride, P, st, groups, part);
        }
        {
            LaunchScope ls(c, "prep");
            col_mean_kernel<<<ceil_div(P, 256), 256, 0, c->stream>>>(part, groups, P, n_sub, xbar);
        }
    }
    if (rows_pad > n) {
        ESPER_CUDA(c, cudaMemsetAsync(c->plane_hi.as<float>() + (size_t)n * ld, 0,
                                      sizeof(float) * (size_t)(rows_pad - n) * ld, c->stream));
        ESPER_CUDA(c, cudaMemsetAsync(c->plane_lo.as<float>() + (size_t)n * ld, 0,
                                      sizeof(float) * (size_t)(rows_pad - n) * ld, c->stream));
    }
    {
        LaunchScope ls(c, "prep");
        stats_split_kernel<<<n, 512, 0, c->stream>>>(raw, n, P, ld, standardize, st, xbar, center,
                                                     c->plane_hi.as<float>(), c->plane_lo.as<float>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

}  // namespace esper
