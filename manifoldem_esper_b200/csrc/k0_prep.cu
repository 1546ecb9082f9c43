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
#include <cuda_fp16.h>
#include "common.cuh"
#include "devutil.cuh"

namespace esper {

// L2 cache-policy hints for the two sweeps of a row (stats_split_kernel): the statistics sweep asks the L2 to KEEP the row
// (evict_last), the split sweep reads it for the last time (evict_first) and streams the operand planes out (evict_first),
// so that ~300 rows in flight plus 0.5 GB of planes passing through do not push the rows out before their second read
// (round 1 / round 2 ncu: 944 MB read from DRAM for a 500 MB stack).
#ifndef ESPER_PREP_EMU
__device__ __forceinline__ float4 ld_keep(const float4* p) {
    uint64_t pol; float4 v;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_last_use(const float4* p) {
    uint64_t pol; float4 v;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream(uint2* p, uint2 v) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
#endif

__device__ __forceinline__ void row_stats(const float* __restrict__ row, int P, double* sh, float& mean32, float& std32) {
    double s = 0.0, s2 = 0.0;
    if ((P & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        int n4 = P >> 2;
        for (int q = threadIdx.x; q < n4; q += blockDim.x) {
            float4 v = ld_keep(r4 + q);
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
    if (standardize == 2) return;                        // statistics already in st (np_stats_kernel)
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

// One block per image: write the hi / lo operand planes (row stride ld elements, zero padded) and |c|^2.
//   HALF = false: TF32 values in float32 storage (3xTF32 Gram, tcgen05 kind::tf32)
//   HALF = true : FP16 values, c scaled by the exact power of two `scale` first (3xFP16 Gram, kind::f16, opt-in).  An FP16
//                 significand has the 11 bits of a TF32 one, so hi + lo carries the same 22 bits; the scale keeps lo out
//                 of the FP16 subnormal range, |c| <= 2 sqrt(P) for z-scores bounds the top (host: gram_f16_scale_log2).
template <bool HALF>
__global__ void __launch_bounds__(512) stats_split_kernel(const float* __restrict__ raw, int n, int P, int ld,
                                                          int standardize, RowStat* __restrict__ st,
                                                          const float* __restrict__ xbar, int center, float scale,
                                                          void* __restrict__ hi_v, void* __restrict__ lo_v) {
    __shared__ double sh[32];
    int i = blockIdx.x;
    const float* row = raw + (size_t)i * P;
    float m = 0.f, sd = 1.f;
    if (standardize == 1) row_stats(row, P, sh, m, sd);
    else if (standardize == 2) { m = st[i].mean; sd = st[i].stdv; }      // numpy's own float32 statistics (np_stats_kernel)
    const bool vec = ((P & 3) == 0);
    double acc = 0.0;
    int nq = ld >> 2;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        int p0 = q << 2;
        float v[4];
        if (vec && p0 + 3 < P) {
            float4 t = ld_last_use(reinterpret_cast<const float4*>(row) + q);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (p0 + e < P) ? __ldg(row + p0 + e) : 0.f;
        }
        float cv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float c = 0.f;
            if (p0 + e < P) {
                c = zscore(v[e], m, sd);
                if (center) c = __fsub_rn(c, __ldg(xbar + p0 + e));
            }
            cv[e] = c;
            acc += (double)c * (double)c;
        }
        if (!HALF) {
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                h[e] = tf32_rna(cv[e]);
                l[e] = tf32_rna(__fsub_rn(cv[e], h[e]));
            }
            st_stream(reinterpret_cast<float4*>(reinterpret_cast<float*>(hi_v) + (size_t)i * ld) + q, make_float4(h[0], h[1], h[2], h[3]));
            st_stream(reinterpret_cast<float4*>(reinterpret_cast<float*>(lo_v) + (size_t)i * ld) + q, make_float4(l[0], l[1], l[2], l[3]));
        } else {
            __half h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float cs = __fmul_rn(cv[e], scale);             // exact: scale is a power of two
                h[e] = __float2half_rn(cs);
                l[e] = __float2half_rn(__fsub_rn(cs, __half2float(h[e])));
            }
            uint2 hv, lv;
            hv.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
            hv.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
            lv.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
            lv.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
            st_stream(reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi_v) + (size_t)i * ld) + q, hv);
            st_stream(reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo_v) + (size_t)i * ld) + q, lv);
        }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) { st[i].mean = m; st[i].stdv = sd; st[i].norm2 = acc; }
}

// ------------------------------------------------------------------------------------------------
// numpy-exact float32 statistics.  On the CTF branch the reference standardises with `(tmp - tmp.mean()) / tmp.std()`
// on a float32 array (Distances.py:114-115), i.e. numpy's float32 pairwise summation: blocks of <= 128 elements summed
// with 8 interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), halves split at n/2 rounded down
// to a multiple of 8 (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum), mean = sum / n, and the two-pass
// variance of numpy/_core/_methods.py:_var (x = arr - mean; sum(x*x) / n; sqrt), all in float32.  Reproducing the
// summation tree makes the z-scores bit-identical to the reference's, which matters for near-duplicate images on
// that branch (a 1-ulp difference of the float32 std is a 6e-8 rescaling of one image, and the defocus-tolerant
// distance is not invariant to it the way the plain Euclidean one is).
// The tree (leaves and combine nodes by level) depends on P only and is built once on the host.
// ------------------------------------------------------------------------------------------------
struct PwNode { int left, right, out; };

// kind 0: x;  1: (x - mean)^2;  2: x - mean;  3: ((x - mean) - mean2)^2   (every operation rounded to float32)
__device__ __forceinline__ float pw_leaf(const float* __restrict__ a, int len, int lane8, float mean, float mean2, int kind) {
    auto elem = [&](int i) -> float {
        const float x = __ldg(a + i);
        if (kind == 0) return x;
        float d = __fsub_rn(x, mean);
        if (kind == 2) return d;
        if (kind == 3) d = __fsub_rn(d, mean2);
        return __fmul_rn(d, d);
    };
    if (len < 8) {                                   // only a whole array shorter than 8 elements
        float res = 0.f;
        for (int i = 0; i < len; ++i) res = __fadd_rn(res, elem(i));
        return res;
    }
    const int n8 = len - (len % 8);
    float r = elem(lane8);
    for (int i = 8; i < n8; i += 8) r = __fadd_rn(r, elem(i + lane8));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    for (int i = n8; i < len; ++i) r = __fadd_rn(r, elem(i));
    return r;
}

// centered_std = 0: mean() and std() of the image as it is (`(tmp - tmp.mean()) / tmp.std()`, Distances.py:114-115).
// centered_std = 1: the non-CTF branch (Distances.py:89-91) subtracts the mean in place first and takes std() of the centred
//                   float32 image, whose own float32 mean is a tiny non-zero number that numpy subtracts again inside std().
__global__ void __launch_bounds__(256) np_stats_kernel(const float* __restrict__ raw, int P, const int2* __restrict__ leaves,
                                                       int L, const PwNode* __restrict__ nodes, const int* __restrict__ level_off,
                                                       int n_levels, int root, RowStat* __restrict__ st, int centered_std) {
    extern __shared__ float slot[];
    const int img = blockIdx.x;
    const float* a = raw + (size_t)img * P;
    const int group = threadIdx.x >> 3, lane8 = threadIdx.x & 7, n_groups = blockDim.x >> 3;
    float mean = 0.f, mean2 = 0.f, stdv = 1.f;
    const int n_pass = centered_std ? 3 : 2;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int kind = centered_std ? (pass == 0 ? 0 : pass == 1 ? 2 : 3) : pass;
        for (int lf = group; lf < ((L + n_groups - 1) / n_groups) * n_groups; lf += n_groups) {   // whole warps stay converged
            const bool live = lf < L;
            const int2 ol = live ? leaves[lf] : make_int2(0, 8);
            const float v = pw_leaf(a + ol.x, live ? ol.y : (P < 8 ? P : 8), lane8, mean, mean2, kind);
            if (live && lane8 == 0) slot[lf] = v;
        }
        __syncthreads();
        for (int lv = 0; lv < n_levels; ++lv) {
            for (int q = level_off[lv] + threadIdx.x; q < level_off[lv + 1]; q += blockDim.x) {
                const PwNode nd = nodes[q];
                slot[nd.out] = __fadd_rn(slot[nd.left], slot[nd.right]);
            }
            __syncthreads();
        }
        const float total = slot[root];
        __syncthreads();
        if (pass == 0) mean = __fdiv_rn(total, (float)P);
        else if (centered_std && pass == 1) mean2 = __fdiv_rn(total, (float)P);
        else stdv = __fsqrt_rn(__fdiv_rn(total, (float)P));
    }
    if (threadIdx.x == 0) { st[img].mean = mean; st[img].stdv = stdv; st[img].norm2 = 0.0; }
}

namespace {
struct PwBuild {
    std::vector<int2> leaves;
    struct N { int left, right, height, slot; };
    std::vector<N> nodes;                 // every tree node; leaves have left = -1
    int build(int off, int n) {
        if (n <= 128) {
            leaves.push_back(make_int2(off, n));
            nodes.push_back({-1, -1, 0, (int)leaves.size() - 1});
            return (int)nodes.size() - 1;
        }
        int n2 = n / 2;
        n2 -= n2 % 8;
        const int l = build(off, n2), r = build(off + n2, n - n2);
        nodes.push_back({l, r, std::max(nodes[l].height, nodes[r].height) + 1, -1});
        return (int)nodes.size() - 1;
    }
};
}  // namespace

// float32 mean / population std of every image of the resident stack, bit-identical to numpy's (see above)
int row_stats_numpy(esper_ctx* c, int n, int P, RowStat* st, int centered_std) {
    if (c->pw_P != P) {
        PwBuild b;
        const int root = b.build(0, P);
        const int L = (int)b.leaves.size();
        int max_h = b.nodes[root].height, next = L;
        std::vector<PwNode> prog;
        std::vector<int> level_off(1, 0);
        for (int h = 1; h <= max_h; ++h) {
            for (auto& nd : b.nodes) if (nd.left >= 0 && nd.height == h) nd.slot = next++;
            for (auto& nd : b.nodes) if (nd.left >= 0 && nd.height == h) prog.push_back({b.nodes[nd.left].slot, b.nodes[nd.right].slot, nd.slot});
            level_off.push_back((int)prog.size());
        }
        const size_t bytes = sizeof(int2) * (size_t)L + sizeof(PwNode) * (prog.size() + 1) + sizeof(int) * level_off.size();
        ESPER_RESERVE(c, c->pw_plan, bytes + 64);
        char* base = c->pw_plan.as<char>();
        ESPER_CUDA(c, cudaMemcpyAsync(base, b.leaves.data(), sizeof(int2) * (size_t)L, cudaMemcpyHostToDevice, c->stream));
        if (!prog.empty())
            ESPER_CUDA(c, cudaMemcpyAsync(base + sizeof(int2) * (size_t)L, prog.data(), sizeof(PwNode) * prog.size(), cudaMemcpyHostToDevice, c->stream));
        ESPER_CUDA(c, cudaMemcpyAsync(base + sizeof(int2) * (size_t)L + sizeof(PwNode) * (prog.size() + 1), level_off.data(),
                                      sizeof(int) * level_off.size(), cudaMemcpyHostToDevice, c->stream));
        ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
        c->pw_P = P; c->pw_L = L; c->pw_nodes = (int)prog.size(); c->pw_levels = max_h; c->pw_root = b.nodes[root].slot;
    }
    char* base = c->pw_plan.as<char>();
    const int2* leaves = reinterpret_cast<const int2*>(base);
    const PwNode* nodes = reinterpret_cast<const PwNode*>(base + sizeof(int2) * (size_t)c->pw_L);
    const int* level_off = reinterpret_cast<const int*>(base + sizeof(int2) * (size_t)c->pw_L + sizeof(PwNode) * ((size_t)c->pw_nodes + 1));
    const size_t smem = sizeof(float) * (size_t)(c->pw_L + c->pw_nodes + 1);
    if (smem > 48 * 1024) return fail(c, ESPER_E_ARG, "row statistics: %d pixels per image need %zu bytes of shared memory", P, smem);
    np_stats_kernel<<<n, 256, smem, c->stream>>>(c->imgs.as<float>(), P, leaves, c->pw_L, nodes, level_off, c->pw_levels, c->pw_root, st, centered_std);
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

// half_scale_log2 < 0: TF32 planes (float32 storage, ld a multiple of 32); >= 0: FP16 planes scaled by 2^half_scale_log2
// (ld a multiple of 64).
int run_prep(esper_ctx* c, int n, int P, unsigned flags, int rows_pad, int ld, bool exact_center, int half_scale_log2, bool numpy_stats) {
    const float* raw = c->imgs.as<float>();
    ESPER_RESERVE(c, c->row_stats, sizeof(RowStat) * (size_t)rows_pad);
    const bool half = half_scale_log2 >= 0;
    const size_t esz = half ? sizeof(__half) : sizeof(float);
    ESPER_RESERVE(c, c->plane_hi, esz * (size_t)rows_pad * ld);
    ESPER_RESERVE(c, c->plane_lo, esz * (size_t)rows_pad * ld);
    RowStat* st = c->row_stats.as<RowStat>();
    int standardize = (flags & ESPER_DIST_STANDARDIZE) ? 1 : 0;
    const int center = (flags & ESPER_DIST_CENTER) ? 1 : 0;
    if (standardize && numpy_stats && (size_t)P <= (size_t)1 << 19) {
        // small stacks: mean32 / std32 exactly as numpy computes them for `image -= image.mean(); image /= image.std()`
        // (Distances.py:89-91), so that the z-scores -- and with them the pair-by-pair distances -- carry the reference's bits
        LaunchScope ls(c, "prep");
        int rc = row_stats_numpy(c, n, P, st, 1);
        if (rc) return rc;
        standardize = 2;
    }
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
        ESPER_CUDA(c, cudaMemsetAsync(c->plane_hi.as<char>() + esz * (size_t)n * ld, 0, esz * (size_t)(rows_pad - n) * ld, c->stream));
        ESPER_CUDA(c, cudaMemsetAsync(c->plane_lo.as<char>() + esz * (size_t)n * ld, 0, esz * (size_t)(rows_pad - n) * ld, c->stream));
    }
    {
        LaunchScope ls(c, "prep");
        const float scale = half ? ldexpf(1.0f, half_scale_log2) : 1.0f;
        if (half)
            stats_split_kernel<true><<<n, 512, 0, c->stream>>>(raw, n, P, ld, standardize, st, xbar, center, scale, c->plane_hi.p, c->plane_lo.p);
        else
            stats_split_kernel<false><<<n, 512, 0, c->stream>>>(raw, n, P, ld, standardize, st, xbar, center, scale, c->plane_hi.p, c->plane_lo.p);
    }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

}  // namespace esper
