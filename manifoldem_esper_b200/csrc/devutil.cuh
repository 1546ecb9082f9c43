// devutil.cuh -- small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace esper {

struct RowStat { float mean; float stdv; double norm2; };   // per image: z-score parameters, |c|^2

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of one double per thread (blockDim.x <= 1024, sh >= 32 doubles); result in thread 0.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = (l < nw) ? sh[l] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

__device__ __forceinline__ float zscore(float raw, float mean, float stdv) {
    // float32 subtract then float32 divide, as numpy does for `image -= mean; image /= std`
    return __fdiv_rn(__fsub_rn(raw, mean), stdv);
}

}  // namespace esper
