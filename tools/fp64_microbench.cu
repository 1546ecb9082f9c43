// fp64_microbench.cu -- what the FP64 pipe, warp shuffles and shared-memory loads of one B200 SM deliver (cycles per phase of
// a 512-thread CTA), to size the matrix-vector phase of the chain eigensolver.  nvcc -arch=sm_100a -O3 -o fp64_microbench ...
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) bench(double* out, long long* cyc, int iters) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 20000; i += 512) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = tid * 1e-3 + q;
    const double x = 1.0000001, y = 1e-9;
    long long t0, t1;
    // 1. independent DFMA: 16 chains per thread
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] = fma(a[q], x, y);
    }
    __syncthreads(); t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    // 2. dependent DFMA chain (latency)
    double d = a[0];
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) d = fma(d, x, y);
    }
    __syncthreads(); t1 = clock64();
    if (tid == 0) cyc[1] = t1 - t0;
    a[1] += d;
    // 3. LDS.64, consecutive lanes consecutive doubles, 16 independent loads per iteration (aligned to 128 B)
    double s = 0.0;
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) s += sm[q * 1024 + tid + (it & 1) * 512];
    }
    __syncthreads(); t1 = clock64();
    if (tid == 0) cyc[2] = t1 - t0;
    // 4. the same, misaligned by 48 bytes
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) s += sm[q * 1024 + tid + (it & 1) * 512 + 6];
    }
    __syncthreads(); t1 = clock64();
    if (tid == 0) cyc[3] = t1 - t0;
    // 5. shuffles of doubles: 16 per iteration
    double h = s;
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) h += __shfl_xor_sync(0xffffffffu, h, 1 + (q & 15));
    }
    __syncthreads(); t1 = clock64();
    if (tid == 0) cyc[4] = t1 - t0;
    // 6. LDS + DFMA as in the matrix-vector phase: 10 loads, 17 fma per column, 4 columns
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double qv = sm[tid + 512 * u + 16000];
#pragma unroll
            for (int q = 0; q < 10; ++q) a[q] = fma(sm[q * 1600 + tid + 512 * u - (u == 3 ? 500 : 0)], qv, a[q]);
#pragma unroll
            for (int q = 10; q < 16; ++q) a[q] = fma(a[(q + 1) & 15], qv, a[q]);
        }
    }
    __syncthreads(); t1 = clock64();
    if (tid == 0) cyc[5] = t1 - t0;
    double r = h;
#pragma unroll
    for (int q = 0; q < 16; ++q) r += a[q];
    out[blockIdx.x * 512 + tid] = r;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    const int iters = 64;
    for (int rep = 0; rep < 2; ++rep) bench<<<1, 512, 200000>>>(out, cyc, iters);
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    const double n = iters;
    printf("512 threads on one SM, cycles per iteration:\n");
    printf(" 16 independent DFMA per thread      %.0f  (%.2f cycles per warp instruction per SM sub-partition)\n", h[0] / n, h[0] / n / 16 / 4);
    printf(" 16 dependent DFMA per thread        %.0f  (latency ~%.1f cycles with 4 warps per scheduler)\n", h[1] / n, h[1] / n / 16);
    printf(" 16 LDS.64 per thread, aligned       %.0f  (%.2f cycles per warp load)\n", h[2] / n, h[2] / n / 16 / 16);
    printf(" 16 LDS.64 per thread, +48 bytes     %.0f  (%.2f cycles per warp load)\n", h[3] / n, h[3] / n / 16 / 16);
    printf(" 16 dependent shfl.f64 per thread    %.0f  (%.1f cycles each)\n", h[4] / n, h[4] / n / 16);
    printf(" matvec-like: 44 LDS.64 + 64 DFMA    %.0f\n", h[5] / n);
    return 0;
}
