// k8_dataprep.cu -- data formats either side of the hot path:
//   * tau noisy duplicates per pristine state at a target SNR, standardised on the background ring:
//     0_Data_Inputs/Pristine_AddTauSNR/AddNoise.py:17-86 (find_SNR, add_noise, standardize) driven by
//     Generate_TauSNR.py:56-64 (image index = state * tau + copy).  The stack is generated where stage 1 reads it.
//     The reference draws its noise from numpy's global Mersenne Twister (unseeded); here it is a counter-based
//     Philox4x32-10 stream keyed by (seed; image, pixel pair) with a Box-Muller transform in FP64, so any image can
//     be regenerated independently (restated in numpy by synth.philox_normal for the tests).
//   * MRC2014 mode-2 stacks streamed from disk through two page-locked buffers (the reference: mrcfile.mmap,
//     1_Embedding/DM/Distances.py:75-76), so that the read of chunk i+1 overlaps the H2D copy of chunk i.
//   * bin-average accumulation of the consumer, 3_Manifold_Binning/Manifold_Binning.py:940-974: float64 sums of the
//     float32 images of every bin, added in the caller's order (bit-identical to the reference's `+=` loop).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.cuh"
#include "devutil.cuh"

namespace esper {
namespace dataprep {

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;

// ---- Philox4x32-10 (Salmon et al., SC'11) ----------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two independent N(0,1) values for pixel pair `pair` of image `image`.
__device__ __forceinline__ void normal_pair(unsigned long long seed, uint32_t image, uint32_t pair, double& z0, double& z1) {
    uint32_t x[4];
    philox4x32_10(pair, image, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
    const double u0 = ((double)(x[0] >> 5) * 67108864.0 + (double)(x[1] >> 6)) * (1.0 / 9007199254740992.0);   // [0, 1)
    const double u1 = ((double)(x[2] >> 5) * 67108864.0 + (double)(x[3] >> 6)) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(1.0 - u0));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    z0 = r * c; z1 = r * s;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < WARPS; ++w) t += red[w];
    return t;
}

// AddNoise.find_SNR (:17-39): signal = pixels outside +-0.25 std of the image; noise mean = signal mean,
// noise std = sqrt(var(signal) / SNR).  One CTA per pristine image; out: (sig_mean, noise_std) per image.
__global__ void __launch_bounds__(THREADS)
snr_params_kernel(const float* __restrict__ pristine, int P, double snr, double* __restrict__ out) {
    __shared__ double red[WARPS];
    const float* img = pristine + (size_t)blockIdx.x * P;
    double s = 0.0;
    for (int i = threadIdx.x; i < P; i += THREADS) s += (double)img[i];
    const double mean = block_sum(s, red) / P;
    double v = 0.0;
    for (int i = threadIdx.x; i < P; i += THREADS) { const double d = (double)img[i] - mean; v += d * d; }
    const double thr = 0.25 * sqrt(block_sum(v, red) / P);
    double cnt = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < P; i += THREADS) {
        const double x = (double)img[i];
        if (!(-thr < x && x < thr)) { cnt += 1.0; ss += x; }
    }
    cnt = block_sum(cnt, red);
    const double sig_mean = block_sum(ss, red) / cnt;
    double sv = 0.0;
    for (int i = threadIdx.x; i < P; i += THREADS) {
        const double x = (double)img[i];
        if (!(-thr < x && x < thr)) { const double d = x - sig_mean; sv += d * d; }
    }
    const double sig_var = block_sum(sv, red) / cnt;
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = sig_mean; out[2 * blockIdx.x + 1] = sqrt(sig_var / snr); }
}

// AddNoise.add_noise + standardize (:43-75) for output image s * tau + t.  Background = pixels farther than
// int(box/2) from (int(box/2), int(box/2)).  Two sweeps regenerate the same noise (statistics, then the output).
__global__ void __launch_bounds__(THREADS)
tau_snr_kernel(const float* __restrict__ pristine, int box, int tau, const double* __restrict__ params, unsigned long long seed,
               float* __restrict__ stack) {
    __shared__ double red[WARPS];
    const int image = blockIdx.x, state = image / tau, P = box * box;
    const float* src = pristine + (size_t)state * P;
    float* dst = stack + (size_t)image * P;
    const double mu = params[2 * state], sd = params[2 * state + 1];
    const int cx = box / 2, r2 = (box / 2) * (box / 2), pairs = (P + 1) / 2;
    double s = 0.0, q = 0.0, nb = 0.0;
    for (int p = threadIdx.x; p < pairs; p += THREADS) {
        double z[2];
        normal_pair(seed, (uint32_t)image, (uint32_t)p, z[0], z[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * p + h;
            if (i >= P) break;
            const int y = i / box - cx, x = i % box - cx;
            if (x * x + y * y > r2) {
                const double d = ((double)src[i] + (mu + sd * z[h])) - mu;     // shifted by the noise mean
                s += d; q += d * d; nb += 1.0;
            }
        }
    }
    nb = block_sum(nb, red);
    s = block_sum(s, red);
    q = block_sum(q, red);
    const double bg_mean = mu + s / nb;
    const double var = q / nb - (s / nb) * (s / nb);
    const double bg_std = sqrt(var > 0.0 ? var : 0.0);
    for (int p = threadIdx.x; p < pairs; p += THREADS) {
        double z[2];
        normal_pair(seed, (uint32_t)image, (uint32_t)p, z[0], z[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * p + h;
            if (i >= P) break;
            const double noisy = (double)src[i] + (mu + sd * z[h]);
            dst[i] = (nb > 0.0 && bg_std > 0.0) ? (float)((noisy - bg_mean) / bg_std) : 0.f;   // :68-69: img_norm = 0
        }
    }
}

// apply_noise = False (Generate_TauSNR.py:62-63): tau exact copies of every state.
__global__ void duplicate_kernel(const float4* __restrict__ pristine, long long P4, int tau, float4* __restrict__ stack, long long total4) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (long long)gridDim.x * blockDim.x) {
        const long long image = e / P4, off = e % P4;
        stack[e] = pristine[(image / tau) * P4 + off];
    }
}
__global__ void duplicate_scalar_kernel(const float* __restrict__ pristine, long long P, int tau, float* __restrict__ stack, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long image = e / P, off = e % P;
        stack[e] = pristine[(image / tau) * P + off];
    }
}

// Manifold_Binning.py:940-974: finalStackBin[b] += image, float64, in the order of `order` within each bin.
// Thread -> one pixel of one bin; the image loop is sequential so the float64 sums equal the reference's bit for bit.
__global__ void __launch_bounds__(256)
bin_accumulate_kernel(const float* __restrict__ imgs, int P, const int* __restrict__ order, const int* __restrict__ offsets,
                      double* __restrict__ sums) {
    const int b = blockIdx.y;
    const int beg = offsets[b], end = offsets[b + 1];
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        double acc = 0.0;
        int k = beg;
        for (; k + 4 <= end; k += 4) {                      // four loads in flight, adds still in order
            const float a0 = imgs[(size_t)order[k] * P + pix], a1 = imgs[(size_t)order[k + 1] * P + pix];
            const float a2 = imgs[(size_t)order[k + 2] * P + pix], a3 = imgs[(size_t)order[k + 3] * P + pix];
            acc += (double)a0; acc += (double)a1; acc += (double)a2; acc += (double)a3;
        }
        for (; k < end; ++k) acc += (double)imgs[(size_t)order[k] * P + pix];
        sums[(size_t)b * P + pix] = acc;
    }
}

}  // namespace dataprep

int run_tau_snr_stack(esper_ctx* c, const float* pristine_h, int ss, int box, int tau, double snr, unsigned long long seed,
                      double* params_h, float* stack_h) {
    using namespace dataprep;
    const long long P = (long long)box * box;
    const long long n = (long long)ss * tau;
    if (ss < 1 || box < 1 || tau < 1 || !(snr >= 0.0) || n > 2147483647LL || P > 2147483647LL)
        return fail(c, ESPER_E_ARG, "tau_snr_stack: need ss, box, tau >= 1 and snr >= 0 (0 = duplicates without noise)");
    // pristine images, then the (sig_mean, noise_std) table, 8-byte aligned
    const size_t par_off = round_up(sizeof(float) * (size_t)ss * P, 8);
    ESPER_RESERVE(c, c->prep_src, par_off + sizeof(double) * 2 * (size_t)ss);
    c->img_n = 0; c->img_P = 0;
    ESPER_RESERVE(c, c->imgs, sizeof(float) * (size_t)n * P);
    float* src = c->prep_src.as<float>();
    double* par = reinterpret_cast<double*>(c->prep_src.as<char>() + par_off);
    ESPER_CUDA(c, cudaMemcpyAsync(src, pristine_h, sizeof(float) * (size_t)ss * P, cudaMemcpyHostToDevice, c->stream));
    if (snr > 0.0) {
        { LaunchScope ls(c, "dataprep"); snr_params_kernel<<<ss, THREADS, 0, c->stream>>>(src, (int)P, snr, par); }
        { LaunchScope ls(c, "dataprep"); tau_snr_kernel<<<(int)n, THREADS, 0, c->stream>>>(src, box, tau, par, seed, c->imgs.as<float>()); }
        if (params_h) ESPER_CUDA(c, cudaMemcpyAsync(params_h, par, sizeof(double) * 2 * (size_t)ss, cudaMemcpyDeviceToHost, c->stream));
    } else {
        LaunchScope ls(c, "dataprep");
        const long long total = n * P;
        if (P % 4 == 0) duplicate_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(src), P / 4, tau,
                                                                                  reinterpret_cast<float4*>(c->imgs.p), total / 4);
        else duplicate_scalar_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, P, tau, c->imgs.as<float>(), total);
        if (params_h) memset(params_h, 0, sizeof(double) * 2 * (size_t)ss);
    }
    ESPER_CUDA(c, cudaGetLastError());
    c->img_n = (int)n; c->img_P = (int)P;
    if (stack_h) ESPER_CUDA(c, cudaMemcpyAsync(stack_h, c->imgs.p, sizeof(float) * (size_t)n * P, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

// ---- MRC2014 ---------------------------------------------------------------------------------------
// Header words (little-endian int32): 0-2 nx, ny, nz; 3 mode; 23 nsymbt (extended-header bytes); data at 1024 + nsymbt.
int mrc_info_host(const char* path, int* dims, int* mode, long long* data_offset, char* err, size_t err_len) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) { snprintf(err, err_len, "cannot open %s", path); return ESPER_E_ARG; }
    int32_t hdr[256];
    const ssize_t got = pread(fd, hdr, sizeof hdr, 0);
    struct stat st;
    const int frc = fstat(fd, &st);
    close(fd);
    if (got != (ssize_t)sizeof hdr) { snprintf(err, err_len, "%s: shorter than an MRC header", path); return ESPER_E_ARG; }
    const int nx = hdr[0], ny = hdr[1], nz = hdr[2], md = hdr[3], nsymbt = hdr[23];
    if (nx < 1 || ny < 1 || nz < 1 || nsymbt < 0) { snprintf(err, err_len, "%s: invalid MRC dimensions (byte-swapped file?)", path); return ESPER_E_ARG; }
    const int bytes_per = md == 0 ? 1 : md == 1 ? 2 : md == 2 ? 4 : md == 6 ? 2 : md == 12 ? 2 : 0;
    if (!bytes_per) { snprintf(err, err_len, "%s: unsupported MRC mode %d", path, md); return ESPER_E_ARG; }
    const long long off = 1024LL + nsymbt;
    if (frc == 0 && (long long)st.st_size < off + (long long)nx * ny * nz * bytes_per) {
        snprintf(err, err_len, "%s: file is shorter than its header says", path);
        return ESPER_E_ARG;
    }
    dims[0] = nx; dims[1] = ny; dims[2] = nz;
    *mode = md; *data_offset = off;
    return ESPER_OK;
}

int run_upload_mrc(esper_ctx* c, const char* path, int first, int count, int* dims_out) {
    char err[400];
    int dims[3], mode; long long off;
    int rc = mrc_info_host(path, dims, &mode, &off, err, sizeof err);
    if (rc) return fail(c, rc, "upload_mrc: %s", err);
    if (mode != 2) return fail(c, ESPER_E_ARG, "upload_mrc: %s is MRC mode %d; the pipeline reads mode 2 (float32) stacks", path, mode);
    const long long P = (long long)dims[0] * dims[1];
    if (first < 0 || first >= dims[2]) return fail(c, ESPER_E_ARG, "upload_mrc: first image %d outside [0, %d)", first, dims[2]);
    const int n = (count <= 0 || first + count > dims[2]) ? dims[2] - first : count;
    if (P > 2147483647LL) return fail(c, ESPER_E_ARG, "upload_mrc: image too large");
    c->img_n = 0; c->img_P = 0;
    ESPER_RESERVE(c, c->imgs, sizeof(float) * (size_t)n * P);
    const size_t chunk = (size_t)32 << 20;
    if (c->pin_stream[0].reserve(chunk) || c->pin_stream[1].reserve(chunk)) return fail(c, ESPER_E_NOMEM, "upload_mrc: pinned staging buffers");
    cudaEvent_t ev[2];
    ESPER_CUDA(c, cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    ESPER_CUDA(c, cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int fd = open(path, O_RDONLY);
    if (fd < 0) { cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); return fail(c, ESPER_E_ARG, "upload_mrc: cannot open %s", path); }
    const size_t total = sizeof(float) * (size_t)n * P;
    const long long base = off + (long long)first * P * (long long)sizeof(float);
    size_t done = 0; int b = 0; bool used[2] = {false, false};
    rc = ESPER_OK;
    while (done < total) {
        const size_t len = total - done < chunk ? total - done : chunk;
        if (used[b]) cudaEventSynchronize(ev[b]);               // the H2D copy that last read this buffer has finished
        size_t got = 0;
        while (got < len) {
            const ssize_t r = pread(fd, c->pin_stream[b].as<char>() + got, len - got, base + (long long)(done + got));
            if (r <= 0) { rc = fail(c, ESPER_E_ARG, "upload_mrc: read error in %s at byte %lld", path, base + (long long)(done + got)); break; }
            got += (size_t)r;
        }
        if (rc) break;
        if (cudaMemcpyAsync(c->imgs.as<char>() + done, c->pin_stream[b].p, len, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
            cudaGetLastError(); rc = fail(c, ESPER_E_CUDA, "upload_mrc: cudaMemcpyAsync failed"); break;
        }
        cudaEventRecord(ev[b], c->stream); used[b] = true;
        done += len; b ^= 1;
    }
    close(fd);
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    if (rc) return rc;
    c->img_n = n; c->img_P = (int)P;
    if (dims_out) { dims_out[0] = n; dims_out[1] = dims[1]; dims_out[2] = dims[0]; }     // (images, ny, nx) like mrcfile's .data.shape
    return ESPER_OK;
}

int run_bin_accumulate(esper_ctx* c, int n, int P, const int* order_h, const int* offsets_h, int n_bins, double* sums_h) {
    using namespace dataprep;
    if (n_bins < 1 || n_bins > 65535) return fail(c, ESPER_E_ARG, "bin_accumulate: need 1 <= n_bins <= 65535");
    if (offsets_h[0] != 0) return fail(c, ESPER_E_ARG, "bin_accumulate: offsets[0] must be 0");
    for (int b = 0; b < n_bins; ++b)
        if (offsets_h[b + 1] < offsets_h[b]) return fail(c, ESPER_E_ARG, "bin_accumulate: offsets must be non-decreasing");
    const int n_sel = offsets_h[n_bins];
    for (int k = 0; k < n_sel; ++k)
        if (order_h[k] < 0 || order_h[k] >= n) return fail(c, ESPER_E_ARG, "bin_accumulate: image index %d out of range at position %d", order_h[k], k);
    ESPER_RESERVE(c, c->rot_idx, sizeof(int) * ((size_t)n_sel + n_bins + 2));
    ESPER_RESERVE(c, c->prep_out, sizeof(double) * (size_t)n_bins * P);
    int* order_d = c->rot_idx.as<int>();
    int* off_d = order_d + n_sel;
    if (n_sel) ESPER_CUDA(c, cudaMemcpyAsync(order_d, order_h, sizeof(int) * (size_t)n_sel, cudaMemcpyHostToDevice, c->stream));
    ESPER_CUDA(c, cudaMemcpyAsync(off_d, offsets_h, sizeof(int) * ((size_t)n_bins + 1), cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, "dataprep");
        const int gx = std::max(1, std::min(ceil_div(P, 256), ceil_div(c->sm_count * 8, n_bins)));
        bin_accumulate_kernel<<<dim3(gx, n_bins), 256, 0, c->stream>>>(c->imgs.as<float>(), P, order_d, off_d, c->prep_out.as<double>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    ESPER_CUDA(c, cudaMemcpyAsync(sums_h, c->prep_out.p, sizeof(double) * (size_t)n_bins * P, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

}  // namespace esper
