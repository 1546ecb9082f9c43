// k6_ctf.cu -- CTF branch of stage 1: defocus-tolerant distances and the Wiener-filtered stack,
// 1_Embedding/DM/Distances.py:107-145 with the filters of Generate_Filters.py.
//
//   reference                                            here
//   z-score (float32), Butterworth low-pass G in          fy_i = G .* FFT2(z_i)   (G real and even, so the reference's
//   Fourier space, fy_i = fft2 of the filtered image      ifft2(...).real -> fft2 round trip is the identity up to 1e-16)
//   CTF_i from (px, Cs, df, kV, ampc), B = inf            evaluated per pixel in FP64 (ctf_gen_kernel)
//   wiener_i = ifft2(fft2(img_i) CTF_i / -(sum CTF^2+.2)) same, inverse transform fused behind the forward one
//   D^2 = |CTF|^2 (|fy|^2)^T + (..)^T - 2 Re(Cf Cf^H)      two products on the tensor-core Gram kernel (k1_gram.cu):
//                                                         T1 = (w C^2) (|fy|^2)^T  (general [n,n]),  T2 = Q Q^T (symmetric),
//                                                         Q_i = sqrt(w) C_i [Re fy_i | Im fy_i]
// Everything is evaluated on the half spectrum kx in [0, N/2] of the real images (weights w = 1 on the self-conjugate
// columns, 2 elsewhere): all summands are even in k.  The K index of the operand planes is k = kx*N + ky.
//
// FFT: hand-written mixed-radix Stockham transform in FP64, one line of N complex values in shared memory per warp, radices 4/2/3/5/7 in registers, any other prime factor by a direct O(r^2) butterfly.  Two real image rows are
// packed into one complex transform.  Passes:
//   rows  : raw rows -> z-score -> FFT along x -> T[img][kx][y]           (ctf_rows_kernel)
//   cols  : T lines -> FFT along y -> filters, operand planes, fy (FP64) -> Wiener multiply -> inverse FFT along y -> T
//   irows : T -> inverse FFT along x -> float32 Wiener image                (ctf_irows_kernel)
// HBM traffic per PD (n images, P pixels): ~ 4nP read + 2 x 8nP (T write/read) + 8nP..(planes) -- a few GB, < 1 ms of
// bandwidth at C2; the two products dominate (about 1.9x the non-CTF Gram work).
#include "common.cuh"
#include "devutil.cuh"

namespace esper {
namespace ctf {

constexpr int MAXF = 16;
constexpr int LINES = 8;             // complex lines per CTA (= 16 image rows in the row passes)
constexpr int THREADS = 256;
struct FftPlan { int N; int nfac; int radix[MAXF]; };
static_assert(THREADS == 32 * LINES, "one warp per line");

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 tw(const double2* W, int e, bool inv) {
    double2 w = W[e];
    if (inv) w.y = -w.y;
    return w;
}

// One radix-R Stockham butterfly j of the stage with sub-transform length Ns (Govindaraju et al. formulation):
//   v_p = in[j + p N/R] W_N^{p (j mod Ns) N/(Ns R)},  out[(j / Ns) Ns R + (j mod Ns) + q Ns] = sum_p v_p W_R^{pq}
// k = j mod Ns, o = (j / Ns) Ns R + k and step = k N/(Ns R) are supplied by the caller.
template <int R>
__device__ __forceinline__ void butterfly(const double2* __restrict__ in, double2* __restrict__ out, int j, int k, int o,
                                          int step, int Ns, int N, const double2* __restrict__ W, bool inv) {
    const int stride = N / R;
    double2 v[R];
#pragma unroll
    for (int p = 0; p < R; ++p) {
        double2 x = in[j + p * stride];
        if (p > 0 && k > 0) x = cmul(x, tw(W, p * step, inv));
        v[p] = x;
    }
    double2 wr[R];
#pragma unroll
    for (int m = 0; m < R; ++m) wr[m] = tw(W, m * stride, inv);
#pragma unroll
    for (int q = 0; q < R; ++q) {
        double2 acc = v[0];
#pragma unroll
        for (int p = 1; p < R; ++p) {
            const double2 t = cmul(v[p], wr[(p * q) % R]);
            acc.x += t.x; acc.y += t.y;
        }
        out[o + q * Ns] = acc;
    }
}

// Any radix (prime factors > 7): no register arrays, inputs re-read from shared memory.
__device__ __forceinline__ void butterfly_generic(const double2* __restrict__ in, double2* __restrict__ out, int j, int o,
                                                  int step, int Ns, int N, int r, const double2* __restrict__ W, bool inv) {
    const int stride = N / r;
    for (int q = 0; q < r; ++q) {
        double2 acc = make_double2(0.0, 0.0);
        for (int p = 0; p < r; ++p) {
            const int e = (int)(((long long)p * step + (long long)((p * q) % r) * stride) % N);
            const double2 t = cmul(in[j + p * stride], tw(W, e, inv));
            acc.x += t.x; acc.y += t.y;
        }
        out[o + q * Ns] = acc;
    }
}

// One warp transforms one line of length N held in `a` (scratch `b`); returns the buffer holding the result.
// Unnormalised; inv = conjugate twiddles.  Only __syncwarp() between stages: a line never leaves its warp.
__device__ double2* fft_line(double2* a, double2* b, const FftPlan& pl, const double2* W, bool inv) {
    const int N = pl.N;
    const int lane = threadIdx.x & 31;
    int Ns = 1;
    double2* in = a; double2* out = b;
    for (int s = 0; s < pl.nfac; ++s) {
        const int r = pl.radix[s];
        const int nb = N / r;
        const int mul = N / (Ns * r);
        const float inv_ns = 1.0f / (float)Ns;
        for (int j = lane; j < nb; j += 32) {
            const int jq = (int)(((float)j + 0.5f) * inv_ns);          // j / Ns (exact for these sizes)
            const int k = j - jq * Ns;
            const int o = jq * Ns * r + k;
            const int step = k * mul;
            switch (r) {
                case 2: butterfly<2>(in, out, j, k, o, step, Ns, N, W, inv); break;
                case 3: butterfly<3>(in, out, j, k, o, step, Ns, N, W, inv); break;
                case 4: butterfly<4>(in, out, j, k, o, step, Ns, N, W, inv); break;
                case 5: butterfly<5>(in, out, j, k, o, step, Ns, N, W, inv); break;
                case 7: butterfly<7>(in, out, j, k, o, step, Ns, N, W, inv); break;
                default: butterfly_generic(in, out, j, o, step, Ns, N, r, W, inv); break;
            }
        }
        __syncwarp();
        Ns *= r;
        double2* t = in; in = out; out = t;
    }
    return in;
}

// Radial frequency in units of Nyquist at FFT-order index (ky, kx): Generate_Filters.py:25-32 after ifftshift.
__device__ __forceinline__ double grid_q(int ky, int kx, int N) {
    const int ax = kx <= N - kx ? kx : N - kx;
    const int ay = ky <= N - ky ? ky : N - ky;
    return __dmul_rn(1.0 / ((double)N / 2.0), sqrt((double)(ax * ax + ay * ay)));
}
// Generate_Filters.py:41-58 with B = inf.  par = (w1, w2, ampc, 2 px).
__device__ __forceinline__ double ctf_eval(double Q, const double* __restrict__ par) {
    const double k = Q / par[3];
    const double k2 = k * k;
    const double wr = (0.5 * par[0] * k2 - par[1]) * k2;
    double sn, cs;
    sincos(wr, &sn, &cs);
    return sn - par[2] * cs;
}

// C[i][k] for all images, Butterworth table Gf[k] and the Wiener denominator sum S[k] = sum_i C_i(k)^2
// (gen_wiener, Distances.py:155-161).  Thread per (k, image chunk); chunk partial sums are added in chunk order.
constexpr int GEN_CHUNKS = 8;
__global__ void __launch_bounds__(256) ctf_gen_kernel(int N, int Kh, int n, const double* __restrict__ par,
                                                      double* __restrict__ C, double* __restrict__ Spart, double* __restrict__ Gf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kh) return;
    const int kx = k / N, ky = k - kx * N;
    const double Q = grid_q(ky, kx, N);
    if (blockIdx.y == 0) Gf[k] = sqrt(1.0 / (1.0 + pow(Q / 0.5, 16.0)));   // create_filter('Butter', 8, 0.5, Q)
    const int i0 = (int)((long long)n * blockIdx.y / gridDim.y), i1 = (int)((long long)n * (blockIdx.y + 1) / gridDim.y);
    double s = 0.0;
    for (int i = i0; i < i1; ++i) {
        const double c = ctf_eval(Q, par + 4 * (size_t)i);
        C[(size_t)i * Kh + k] = c;
        s = s + c * c;
    }
    Spart[(size_t)blockIdx.y * Kh + k] = s;
}
__global__ void __launch_bounds__(256) ctf_sum_kernel(int Kh, int chunks, const double* __restrict__ Spart, double* __restrict__ S) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kh) return;
    double s = 0.0;
    for (int q = 0; q < chunks; ++q) s += Spart[(size_t)q * Kh + k];
    S[k] = s;
}

__device__ __forceinline__ void load_twiddles(double2* W, const double2* __restrict__ Wg, int N) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) W[i] = Wg[i];
}

// rows: 16 image rows -> 8 packed complex lines (one per warp) -> FFT along x -> half spectra, stored transposed
// T[img][kx][y].
__global__ void __launch_bounds__(THREADS, 3) ctf_rows_kernel(const float* __restrict__ raw, const RowStat* __restrict__ st,
                                                           int N, int Nh, FftPlan pl, const double2* __restrict__ Wg,
                                                           double2* __restrict__ T) {
    extern __shared__ double2 sm[];
    double2* W = sm; double2* A = W + N; double2* B = A + (size_t)LINES * N;
    const int img = blockIdx.y, y0 = blockIdx.x * 2 * LINES;
    const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_twiddles(W, Wg, N);
    const float mean = st[img].mean, stdv = st[img].stdv;
    const float* src = raw + (size_t)img * N * N;
    {
        const int ya = y0 + 2 * l, yb = ya + 1;
        const float* ra = src + (size_t)(ya < N ? ya : 0) * N;
        const float* rb = src + (size_t)(yb < N ? yb : 0) * N;
        for (int x0 = lane; x0 < N; x0 += 128) {                      // 8 independent loads in flight per lane
            float va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = x0 + 32 * u;
                va[u] = x < N ? __ldg(ra + x) : 0.f;
                vb[u] = x < N ? __ldg(rb + x) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = x0 + 32 * u;
                if (x < N)
                    A[(size_t)l * N + x] = make_double2(ya < N ? (double)zscore(va[u], mean, stdv) : 0.0,
                                                        yb < N ? (double)zscore(vb[u], mean, stdv) : 0.0);
            }
        }
    }
    __syncthreads();                                                   // twiddles
    const double2* Rl = fft_line(A + (size_t)l * N, B + (size_t)l * N, pl, W, false);
    const bool in_a = (Rl == A + (size_t)l * N);                       // same for every warp
    __syncthreads();
    const double2* R = in_a ? A : B;
    for (int idx = threadIdx.x; idx < Nh * 2 * LINES; idx += blockDim.x) {
        const int kx = idx >> 4, r = idx & 15;
        const int y = y0 + r;
        if (y >= N) continue;
        const int ll = r >> 1;
        const double2 zk = R[(size_t)ll * N + kx];
        double2 zm = R[(size_t)ll * N + (kx ? N - kx : 0)];
        zm.y = -zm.y;
        double2 o;
        if ((r & 1) == 0) o = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
        else              o = make_double2(0.5 * (zk.y - zm.y), -0.5 * (zk.x - zm.x));
        T[((size_t)img * Nh + kx) * N + y] = o;
    }
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_store(float* __restrict__ hi, float* __restrict__ lo, size_t at, double v) {
    const float c = (float)v;
    const float h = tf32_rna(c);
    hi[at] = h;
    lo[at] = tf32_rna(__fsub_rn(c, h));
}

// cols: 8 kx lines of one image: FFT along y, Butterworth, operand planes + fy; Wiener multiply and inverse FFT along y.
__global__ void __launch_bounds__(THREADS, 3)
ctf_cols_kernel(double2* __restrict__ T, int N, int Nh, FftPlan pl, const double2* __restrict__ Wg,
                const double* __restrict__ C, const double* __restrict__ S, const double* __restrict__ Gf, double2* __restrict__ fy_out,
                float* __restrict__ q_hi, float* __restrict__ q_lo, int ldq,
                float* __restrict__ a_hi, float* __restrict__ a_lo, float* __restrict__ b_hi, float* __restrict__ b_lo, int lda,
                int do_wiener) {
    extern __shared__ double2 sm[];
    double2* W = sm; double2* A = W + N; double2* B = A + (size_t)LINES * N;
    const int img = blockIdx.y, kx0 = blockIdx.x * LINES;
    const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nl = min(LINES, Nh - kx0);
    const int Kh = Nh * N;
    load_twiddles(W, Wg, N);
    __syncthreads();
    if (l >= nl) return;                                               // no CTA-wide barrier below
    const int kx = kx0 + l;
    double2* line = T + ((size_t)img * Nh + kx) * N;
    double2* a = A + (size_t)l * N; double2* b = B + (size_t)l * N;
    for (int y0 = lane; y0 < N; y0 += 128) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (y0 + 32 * u < N) ? line[y0 + 32 * u] : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 4; ++u) if (y0 + 32 * u < N) a[y0 + 32 * u] = v[u];
    }
    __syncwarp();
    double2* R = fft_line(a, b, pl, W, false);
    double2* other = (R == a) ? b : a;
    const bool self = (kx == 0) || (2 * kx == N);
    const double w = self ? 1.0 : 2.0, sw = self ? 1.0 : 1.4142135623730951;
#pragma unroll 2
    for (int ky = lane; ky < N; ky += 32) {
        const int k = kx * N + ky;
        const double G = Gf[k];
        const double c = C[(size_t)img * Kh + k];
        const double2 F = R[ky];
        const double2 f = make_double2(G * F.x, G * F.y);
        if (fy_out) fy_out[(size_t)img * Kh + k] = f;
        split_store(q_hi, q_lo, (size_t)img * ldq + k, sw * c * f.x);
        split_store(q_hi, q_lo, (size_t)img * ldq + Kh + k, sw * c * f.y);
        split_store(a_hi, a_lo, (size_t)img * lda + k, f.x * f.x + f.y * f.y);
        split_store(b_hi, b_lo, (size_t)img * lda + k, w * c * c);
        if (do_wiener) {
            const double t = c / (-(S[k] + 1.0 / 5.0));                // Distances.py:129,133 (gen_wiener SNR = 5)
            other[ky] = make_double2(f.x * t, f.y * t);
        }
    }
    if (!do_wiener) return;
    __syncwarp();
    const double2* R2 = fft_line(other, R, pl, W, true);
    for (int y = lane; y < N; y += 32) line[y] = R2[y];
}

// irows: inverse FFT along x of 16 rows (two real rows per complex line) -> float32 Wiener image rows.
__global__ void __launch_bounds__(THREADS, 3) ctf_irows_kernel(const double2* __restrict__ T, int N, int Nh, FftPlan pl,
                                                            const double2* __restrict__ Wg, float* __restrict__ wiener) {
    extern __shared__ double2 sm[];
    double2* W = sm; double2* A = W + N; double2* B = A + (size_t)LINES * N; double2* Hs = B + (size_t)LINES * N;
    const int img = blockIdx.y, y0 = blockIdx.x * 2 * LINES;
    load_twiddles(W, Wg, N);
    for (int idx0 = threadIdx.x; idx0 < Nh * 2 * LINES; idx0 += 4 * THREADS) {     // 4 independent loads in flight per thread
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * THREADS;
            const int kx = idx >> 4, y = y0 + (idx & 15);
            v[u] = (idx < Nh * 2 * LINES && y < N) ? T[((size_t)img * Nh + kx) * N + y] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * THREADS;
            if (idx < Nh * 2 * LINES) Hs[(size_t)(idx & 15) * Nh + (idx >> 4)] = v[u];
        }
    }
    __syncthreads();
    const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* a = A + (size_t)l * N; double2* b = B + (size_t)l * N;
    for (int k = lane; k < N; k += 32) {
        double2 ha, hb;
        if (k < Nh) { ha = Hs[(size_t)(2 * l) * Nh + k]; hb = Hs[(size_t)(2 * l + 1) * Nh + k]; }
        else { ha = Hs[(size_t)(2 * l) * Nh + (N - k)]; hb = Hs[(size_t)(2 * l + 1) * Nh + (N - k)]; ha.y = -ha.y; hb.y = -hb.y; }
        a[k] = make_double2(ha.x - hb.y, ha.y + hb.x);
    }
    __syncwarp();
    const double2* R = fft_line(a, b, pl, W, true);
    const double inv = 1.0 / ((double)N * (double)N);
    float* dst = wiener + (size_t)img * N * N;
    const int ya = y0 + 2 * l, yb = ya + 1;
    for (int x = lane; x < N; x += 32) {
        const double2 v = R[x];
        if (ya < N) dst[(size_t)ya * N + x] = (float)(v.x * inv);
        if (yb < N) dst[(size_t)yb * N + x] = (float)(v.y * inv);
    }
}

// D_ij = sqrt(max(T1_ij + T1_ji - 2 T2_ij, 0)), diagonal 0 (Distances.py:142-145; the reference's `D -= np.diag(D)` subtracts
// a diagonal that is zero up to its own rounding).  Pairs whose D^2 is a small fraction of the summed terms are queued
// for the pair-by-pair evaluation.
__global__ void __launch_bounds__(256) ctf_dist_kernel(const double* __restrict__ T1, const double* __restrict__ T2, int n,
                                                       double tau, double* __restrict__ D, int2* __restrict__ pairs,
                                                       unsigned int* __restrict__ count, unsigned int cap) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    if (i == j) { D[idx] = 0.0; return; }
    const double s = T1[idx] + T1[(size_t)j * n + i];
    const double d2 = s - 2.0 * T2[idx];
    D[idx] = sqrt(d2 > 0.0 ? d2 : 0.0);
    if (pairs && j > i && d2 < tau * s) {
        const unsigned int k = atomicAdd(count, 1u);
        if (k < cap) pairs[k] = make_int2(i, j);
    }
}

// Pair-by-pair form (no Gram expansion): D_ij^2 = sum_k w_k |C_j fy_i - C_i fy_j|^2 in FP64.
__global__ void __launch_bounds__(256) ctf_refine_kernel(const double2* __restrict__ fy, const double* __restrict__ C, int n,
                                                         int N, int Kh, const int2* __restrict__ pairs,
                                                         const unsigned int* __restrict__ count, unsigned int cap,
                                                         double* __restrict__ D) {
    __shared__ double sh[32];
    unsigned int total = *count;
    if (total > cap) total = cap;
    for (unsigned int q = blockIdx.x; q < total; q += gridDim.x) {
        const int2 pr = pairs[q];
        const double2* fi = fy + (size_t)pr.x * Kh; const double2* fj = fy + (size_t)pr.y * Kh;
        const double* ci = C + (size_t)pr.x * Kh; const double* cj = C + (size_t)pr.y * Kh;
        double acc = 0.0;
        for (int k = threadIdx.x; k < Kh; k += blockDim.x) {
            const int kx = k / N;
            const double w = (kx == 0 || 2 * kx == N) ? 1.0 : 2.0;
            const double2 a = fi[k], b = fj[k];
            // separately rounded products (no FMA contraction): identical images give exactly 0
            const double x = __dsub_rn(__dmul_rn(cj[k], a.x), __dmul_rn(ci[k], b.x));
            const double y = __dsub_rn(__dmul_rn(cj[k], a.y), __dmul_rn(ci[k], b.y));
            acc += w * (x * x + y * y);
        }
        acc = block_sum(acc, sh);
        if (threadIdx.x == 0) {
            const double d = sqrt(acc);
            D[(size_t)pr.x * n + pr.y] = d;
            D[(size_t)pr.y * n + pr.x] = d;
        }
        __syncthreads();
    }
}

}  // namespace ctf

static void factor_radices(int N, ctf::FftPlan& pl) {
    pl.N = N; pl.nfac = 0;
    int m = N;
    const int pref[5] = {4, 2, 3, 5, 7};
    for (int r : pref) while (m % r == 0 && pl.nfac < ctf::MAXF) { pl.radix[pl.nfac++] = r; m /= r; }
    for (int p = 11; m > 1 && pl.nfac < ctf::MAXF; p += 2) {
        while (m % p == 0 && pl.nfac < ctf::MAXF) { pl.radix[pl.nfac++] = p; m /= p; }
        if ((long long)p * p > m && m > 1) { pl.radix[pl.nfac++] = m; m = 1; }
    }
}

// params [n, 5] = (pixel size, Cs [mm], defocus [A], voltage [keV], amplitude contrast) per image (host).
int run_ctf_dist(esper_ctx* c, int n, int N, const double* params_h, unsigned flags, bool want_wiener) {
    using namespace ctf;
    const int P = N * N, Nh = N / 2 + 1, Kh = Nh * N;
    const int n_pad = (int)round_up((size_t)n, 128);
    const int lda = (int)round_up((size_t)Kh, 32), ldq = (int)round_up((size_t)2 * Kh, 32);
    const size_t smem_fft = sizeof(double2) * ((size_t)N + 2 * (size_t)LINES * N);
    const size_t smem_irows = smem_fft + sizeof(double2) * (size_t)2 * LINES * Nh;
    if (smem_irows > (size_t)227 * 1024)
        return fail(c, ESPER_E_ARG, "ctf_dist: box size %d needs %zu bytes of shared memory per CTA (limit 227 KB)", N, smem_irows);
    FftPlan pl;
    factor_radices(N, pl);
    {
        long long prod = 1;
        for (int i = 0; i < pl.nfac; ++i) prod *= pl.radix[i];
        if (prod != N) return fail(c, ESPER_E_ARG, "ctf_dist: cannot factor box size %d", N);
    }
    cudaStream_t st = c->stream;

    // per-image CTF constants in the reference's operation order (Generate_Filters.py:42-52), twiddle table
    std::vector<double> par((size_t)n * 4);
    for (int i = 0; i < n; ++i) {
        const double px = params_h[5 * (size_t)i], Cs = params_h[5 * (size_t)i + 1] * 1.0e7, df = params_h[5 * (size_t)i + 2],
                     kev = params_h[5 * (size_t)i + 3], ampc = params_h[5 * (size_t)i + 4];
        if (!(px > 0.0) || !(kev > 0.0)) return fail(c, ESPER_E_ARG, "ctf_dist: image %d has pixel size %g, voltage %g", i, px, kev);
        double wav = (2 * 511.0) + kev;
        wav = 12.3986 / sqrt(wav * kev);
        par[4 * (size_t)i] = M_PI * Cs * wav * wav * wav;
        par[4 * (size_t)i + 1] = M_PI * wav * df;
        par[4 * (size_t)i + 2] = ampc;
        par[4 * (size_t)i + 3] = 2 * px;
    }
    std::vector<double> twid((size_t)2 * N);
    for (int k = 0; k < N; ++k) {
        const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)N;
        twid[2 * (size_t)k] = (double)cosl(a);
        twid[2 * (size_t)k + 1] = (double)sinl(a);
    }
    ESPER_RESERVE(c, c->ctf_par, sizeof(double) * par.size() + sizeof(double) * twid.size());
    double* par_d = c->ctf_par.as<double>();
    double2* W_d = reinterpret_cast<double2*>(par_d + par.size());
    ESPER_CUDA(c, cudaMemcpyAsync(par_d, par.data(), sizeof(double) * par.size(), cudaMemcpyHostToDevice, st));
    ESPER_CUDA(c, cudaMemcpyAsync(W_d, twid.data(), sizeof(double) * twid.size(), cudaMemcpyHostToDevice, st));
    ESPER_CUDA(c, cudaStreamSynchronize(st));                      // the host vectors go out of scope

    ESPER_RESERVE(c, c->row_stats, sizeof(RowStat) * (size_t)n_pad);
    ESPER_RESERVE(c, c->ctf_T, sizeof(double2) * (size_t)n * Kh);
    const bool refine = (flags & ESPER_DIST_REFINE) && c->refine_tau > 0.0 && n > 1;
    if (refine) ESPER_RESERVE(c, c->ctf_fy, sizeof(double2) * (size_t)n * Kh);
    ESPER_RESERVE(c, c->ctf_C, sizeof(double) * ((size_t)n * Kh + (2 + (size_t)GEN_CHUNKS) * Kh));
    ESPER_RESERVE(c, c->plane_hi, sizeof(float) * (size_t)n_pad * ldq);
    ESPER_RESERVE(c, c->plane_lo, sizeof(float) * (size_t)n_pad * ldq);
    ESPER_RESERVE(c, c->ctf_ab, sizeof(float) * (size_t)4 * n_pad * lda);
    ESPER_RESERVE(c, c->ctf_T12, sizeof(double) * (size_t)2 * n * n);
    ESPER_RESERVE(c, c->dist, sizeof(double) * (size_t)n * n);
    if (want_wiener) ESPER_RESERVE(c, c->ctf_wiener, sizeof(float) * (size_t)n * P);
    RowStat* rs = c->row_stats.as<RowStat>();
    double2* T = c->ctf_T.as<double2>();
    double2* fy = refine ? c->ctf_fy.as<double2>() : nullptr;
    double* C = c->ctf_C.as<double>();
    double* S = C + (size_t)n * Kh;
    double* Gf = S + Kh;
    double* Spart = Gf + Kh;
    float* q_hi = c->plane_hi.as<float>(); float* q_lo = c->plane_lo.as<float>();
    float* a_hi = c->ctf_ab.as<float>(); float* a_lo = a_hi + (size_t)n_pad * lda;
    float* b_hi = a_lo + (size_t)n_pad * lda; float* b_lo = b_hi + (size_t)n_pad * lda;
    double* T1 = c->ctf_T12.as<double>(); double* T2 = T1 + (size_t)n * n;
    // the padding (rows n..n_pad, columns K..ld) of the operand planes must be zero; the rest is overwritten
    auto zero_pad = [&](float* plane, int K, int ld) -> cudaError_t {
        cudaError_t e = cudaSuccess;
        if (n_pad > n) e = cudaMemsetAsync(plane + (size_t)n * ld, 0, sizeof(float) * (size_t)(n_pad - n) * ld, st);
        if (e == cudaSuccess && ld > K)
            e = cudaMemset2DAsync(plane + K, sizeof(float) * (size_t)ld, 0, sizeof(float) * (size_t)(ld - K), (size_t)n, st);
        return e;
    };
    ESPER_CUDA(c, zero_pad(q_hi, 2 * Kh, ldq));
    ESPER_CUDA(c, zero_pad(q_lo, 2 * Kh, ldq));
    ESPER_CUDA(c, zero_pad(a_hi, Kh, lda));
    ESPER_CUDA(c, zero_pad(a_lo, Kh, lda));
    ESPER_CUDA(c, zero_pad(b_hi, Kh, lda));
    ESPER_CUDA(c, zero_pad(b_lo, Kh, lda));

    if (!c->attr_ctf) {
        ESPER_CUDA(c, cudaFuncSetAttribute(ctf_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ESPER_CUDA(c, cudaFuncSetAttribute(ctf_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ESPER_CUDA(c, cudaFuncSetAttribute(ctf_irows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        c->attr_ctf = true;
    }
    {
        LaunchScope ls(c, "ctf_stats");
        const int rc_st = row_stats_numpy(c, n, P, rs);
        if (rc_st) return rc_st;
    }
    { LaunchScope ls(c, "ctf_gen"); ctf_gen_kernel<<<dim3(ceil_div(Kh, 256), GEN_CHUNKS), 256, 0, st>>>(N, Kh, n, par_d, C, Spart, Gf); }
    { LaunchScope ls(c, "ctf_gen"); ctf_sum_kernel<<<ceil_div(Kh, 256), 256, 0, st>>>(Kh, GEN_CHUNKS, Spart, S); }
    {
        LaunchScope ls(c, "ctf_rows");
        ctf_rows_kernel<<<dim3(ceil_div(N, 2 * LINES), n), THREADS, smem_fft, st>>>(c->imgs.as<float>(), rs, N, Nh, pl, W_d, T);
    }
    {
        LaunchScope ls(c, "ctf_cols");
        ctf_cols_kernel<<<dim3(ceil_div(Nh, LINES), n), THREADS, smem_fft, st>>>(T, N, Nh, pl, W_d, C, S, Gf, fy, q_hi, q_lo, ldq,
                                                                                 a_hi, a_lo, b_hi, b_lo, lda, want_wiener ? 1 : 0);
    }
    if (want_wiener) {
        LaunchScope ls(c, "ctf_irows");
        ctf_irows_kernel<<<dim3(ceil_div(N, 2 * LINES), n), THREADS, smem_irows, st>>>(T, N, Nh, pl, W_d, c->ctf_wiener.as<float>());
    }
    ESPER_CUDA(c, cudaGetLastError());

    int rc = run_product(c, q_hi, q_lo, n, nullptr, nullptr, n, ldq, T2);
    if (rc) return rc;
    rc = run_product(c, b_hi, b_lo, n, a_hi, a_lo, n, lda, T1);
    if (rc) return rc;

    int2* pairs = nullptr; unsigned int* count = nullptr; unsigned int cap = 0;
    if (refine) {
        const size_t max_pairs = (size_t)n * (n - 1) / 2;
        cap = (unsigned int)(max_pairs < ((size_t)1 << 26) ? max_pairs : (size_t)1 << 26);
        ESPER_RESERVE(c, c->pair_list, sizeof(int2) * (size_t)cap + 256);
        count = c->pair_list.as<unsigned int>();
        pairs = reinterpret_cast<int2*>(c->pair_list.as<char>() + 256);
        ESPER_CUDA(c, cudaMemsetAsync(count, 0, sizeof(unsigned int), st));
    }
    {
        LaunchScope ls(c, "dist_finalize");
        ctf_dist_kernel<<<ceil_div((long long)n * n, 256), 256, 0, st>>>(T1, T2, n, c->refine_tau, c->dist.as<double>(), pairs, count, cap);
    }
    if (refine) {
        LaunchScope ls(c, "refine");
        ctf_refine_kernel<<<c->sm_count * 8, 256, 0, st>>>(fy, C, n, N, Kh, pairs, count, cap, c->dist.as<double>());
    }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

}  // namespace esper
