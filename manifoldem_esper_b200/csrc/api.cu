// api.cu -- the extern "C" surface declared in include/esper_b200.h.
#include "common.cuh"
#include "devutil.cuh"

using namespace esper;

static std::string g_create_error;

#define ESPER_ENTER(c)                                                                      \
    do {                                                                                    \
        if (!(c)) return ESPER_E_ARG;                                                       \
        cudaError_t e__ = cudaSetDevice((c)->device);                                       \
        if (e__ != cudaSuccess) return fail((c), ESPER_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

int esper_abi_version(void) { return ESPER_ABI_VERSION; }

esper_ctx* esper_create(int device, void* stream) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                         " (libesper_b200 has no CPU fallback)";
        return nullptr;
    }
    if (device < 0 || device >= count) { g_create_error = "device ordinal out of range"; return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return nullptr;
    }
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof buf, "device %d (%s) is sm_%d%d; libesper_b200 contains sm_100a code only (tcgen05/TMEM/TMA)",
                 device, prop.name, prop.major, prop.minor);
        g_create_error = buf;
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return nullptr;
    }
    esper_ctx* c = new esper_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (const char* ev = getenv("ESPER_GRAM_F16")) c->gram_mode = (atoi(ev) != 0) ? 1 : 0;           // see esper_gram_config
    if (const char* ev = getenv("ESPER_SWEEP_MOMENTS")) c->sweep_mode = (atoi(ev) != 0) ? 1 : 0;   // see esper_sweep_config
    if (stream) {
        c->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
            delete c;
            return nullptr;
        }
        c->own_stream = true;
    }
    return c;
}

void esper_destroy(esper_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->timers)
        for (auto& pr : kv.second.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto ev : c->event_pool) cudaEventDestroy(ev);
    DevBuf* bufs[] = {&c->imgs, &c->dist, &c->plane_hi, &c->plane_lo, &c->row_stats, &c->col_part, &c->gram_part,
                      &c->pair_list, &c->sweep_part, &c->sweep_aux, &c->markov, &c->degree, &c->lan_V, &c->lan_w,
                      &c->lan_h, &c->lan_small, &c->rot_U, &c->rot_R, &c->rot_idx, &c->rot_out, &c->rot_H,
                      &c->rot_theta, &c->rot_aux, &c->prep_src, &c->prep_out,
                      &c->ctf_par, &c->ctf_T, &c->ctf_fy, &c->ctf_C, &c->ctf_ab, &c->ctf_T12, &c->ctf_wiener, &c->pw_plan, &c->chain_ws, &c->chain_trace};
    run_comm_close(c);
    c->comm_buf.release();
    for (DevBuf* b : bufs) b->release();
    c->pin.release(); c->pin_chain.release();
    c->pin_stream[0].release(); c->pin_stream[1].release();
    if (c->copy_event) cudaEventDestroy(c->copy_event);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* esper_last_error(esper_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

void* esper_stream(esper_ctx* c) { return c ? reinterpret_cast<void*>(c->stream) : nullptr; }

int esper_sync(esper_ctx* c) {
    ESPER_ENTER(c);
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

int esper_set_profiling(esper_ctx* c, int on) {
    ESPER_ENTER(c);
    c->profiling = on != 0;
    c->profiling_gram_only = (on == 2);
    return ESPER_OK;
}

static void drain_timers(esper_ctx* c) {
    for (auto& kv : c->timers) {
        for (auto& pr : kv.second.pending) {
            float ms = 0.f;
            cudaEventSynchronize(pr.second);
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
                kv.second.total_ms += ms;
                kv.second.launches += 1;
            }
            c->event_pool.push_back(pr.first);
            c->event_pool.push_back(pr.second);
        }
        kv.second.pending.clear();
    }
}

int esper_get_kernel_ms(esper_ctx* c, const char* name, double* total_ms, int* launches) {
    ESPER_ENTER(c);
    if (!name) return fail(c, ESPER_E_ARG, "name is NULL");
    drain_timers(c);
    auto it = c->timers.find(name);
    if (total_ms) *total_ms = it == c->timers.end() ? 0.0 : it->second.total_ms;
    if (launches) *launches = it == c->timers.end() ? 0 : it->second.launches;
    return ESPER_OK;
}

int esper_reset_profiling(esper_ctx* c) {
    ESPER_ENTER(c);
    drain_timers(c);
    for (auto& kv : c->timers) { kv.second.total_ms = 0.0; kv.second.launches = 0; }
    return ESPER_OK;
}

long long esper_launch_count(esper_ctx* c) { return c ? c->launches : 0; }

long long esper_stat(esper_ctx* c, const char* name) {
    if (!c || !name) return -1;
    if (!strcmp(name, "lanczos_fallbacks")) return c->lanczos_fallbacks;
    if (!strcmp(name, "cluster_launches")) return c->cluster_launches;
    if (!strcmp(name, "merged_launches")) return c->merged_launches;
    if (!strcmp(name, "launches")) return c->launches;
    return -1;
}

// ---- stage 1 -----------------------------------------------------------------------------------
int esper_upload_images(esper_ctx* c, const float* imgs, int n, int P) {
    ESPER_ENTER(c);
    if (!imgs || n < 1 || P < 1) return fail(c, ESPER_E_ARG, "upload_images: need imgs != NULL, n >= 1, P >= 1");
    ESPER_RESERVE(c, c->imgs, sizeof(float) * (size_t)n * P);
    ESPER_CUDA(c, cudaMemcpyAsync(c->imgs.p, imgs, sizeof(float) * (size_t)n * P, cudaMemcpyHostToDevice, c->stream));
    c->img_n = n; c->img_P = P;
    return ESPER_OK;
}

int esper_dist_config(esper_ctx* c, int split_k, double refine_tau) {
    ESPER_ENTER(c);
    if (split_k < 0 || refine_tau < 0.0) return fail(c, ESPER_E_ARG, "dist_config: negative value");
    c->split_k = split_k;
    c->refine_tau = refine_tau;
    return ESPER_OK;
}

int esper_gram_config(esper_ctx* c, int mode) {
    ESPER_ENTER(c);
    if (mode != 0 && mode != 1) return fail(c, ESPER_E_ARG, "gram_config: mode must be 0 (3xTF32) or 1 (3xFP16)");
    c->gram_mode = mode;
    return ESPER_OK;
}

int esper_gram_info(esper_ctx* c, int* last_operands) {
    ESPER_ENTER(c);
    if (last_operands) *last_operands = c->gram_last_f16;
    return ESPER_OK;
}

int esper_dist_rms(esper_ctx* c, const float* imgs, int n, int P, unsigned flags, double* D) {
    ESPER_ENTER(c);
    NvtxScope nvtx_("esper_dist_rms");
    if (n < 1 || P < 1) return fail(c, ESPER_E_ARG, "dist_rms: need n >= 1 and P >= 1 (got n=%d, P=%d)", n, P);
    if (imgs) {
        int rc = esper_upload_images(c, imgs, n, P);
        if (rc) return rc;
    } else if (c->img_n != n || c->img_P != P || !c->imgs.p) {
        return fail(c, ESPER_E_STATE, "dist_rms: imgs == NULL but no resident [%d, %d] stack (resident: [%d, %d])",
                    n, P, c->img_n, c->img_P);
    }
    const int rows_pad = (int)round_up((size_t)n, 128);
    // operand planes: FP16 pairs (esper_gram_config 1, the default) for z-scored stacks of at least
    // two 128-row tiles (the bound on |c| that makes FP16's range sufficient needs the standardisation), TF32 pairs otherwise
    int hs = -1;
    if (c->gram_mode == 1 && (flags & ESPER_DIST_STANDARDIZE) && rows_pad >= 256) hs = gram_f16_scale_log2(P);
    const int ld = (int)round_up((size_t)P, hs >= 0 ? 64 : 32);
    ESPER_RESERVE(c, c->dist, sizeof(double) * (size_t)n * n);
    c->dist_m = 0;
    const bool exact_small = (flags & ESPER_DIST_REFINE) && (flags & ESPER_DIST_STANDARDIZE) && n <= ESPER_EXACT_PAIRS_MAX_N;
    int rc = run_prep(c, n, P, flags, rows_pad, ld, false, hs, exact_small);
    if (rc) return rc;
    rc = run_gram(c, n, P, rows_pad, ld, flags, -1, 0, nullptr, hs);
    if (rc) return rc;
    c->gram_last_f16 = hs >= 0 ? 1 : 0;
    c->dist_m = n; c->dist_rows = 0; c->dist_row0 = -1;
    c->dist_symmetric = 1;                  // mirrored writes: symmetric by construction
    if (D) {
        ESPER_CUDA(c, cudaMemcpyAsync(D, c->dist.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, c->stream));
        ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return ESPER_OK;
}

int esper_upload_dist(esper_ctx* c, const double* D, int m) {
    ESPER_ENTER(c);
    if (!D || m < 1) return fail(c, ESPER_E_ARG, "upload_dist: need D != NULL and m >= 1");
    ESPER_RESERVE(c, c->dist, sizeof(double) * (size_t)m * m);
    ESPER_CUDA(c, cudaMemcpyAsync(c->dist.p, D, sizeof(double) * (size_t)m * m, cudaMemcpyHostToDevice, c->stream));
    c->dist_m = m; c->dist_rows = 0; c->dist_row0 = -1;
    c->dist_symmetric = -1;
    return ESPER_OK;
}

static int need_dist(esper_ctx* c, const double* D, int m, const char* who) {
    if (m < 1) return fail(c, ESPER_E_ARG, "%s: need m >= 1", who);
    if (D) return esper_upload_dist(c, D, m);
    if (c->dist_m != m || !c->dist.p || c->dist_row0 >= 0)
        return fail(c, ESPER_E_STATE, "%s: D == NULL but no resident %dx%d distance matrix (resident m=%d)", who, m, m, c->dist_m);
    return ESPER_OK;
}

// ---- stage 2 -----------------------------------------------------------------------------------
int esper_eps_sweep(esper_ctx* c, const double* D, int m, const double* coef, int n_eps, double thr, double* logSum) {
    ESPER_ENTER(c);
    NvtxScope nvtx_("esper_eps_sweep");
    if (!coef || !logSum || n_eps < 1) return fail(c, ESPER_E_ARG, "eps_sweep: need coef, logSum != NULL and n_eps >= 1");
    int rc = need_dist(c, D, m, "eps_sweep");
    if (rc) return rc;
    return run_sweep(c, m, coef, n_eps, thr, logSum);
}

int esper_tanh_fit(const double* x, const double* y, int n, const double* a0, double* popt, double* resnorm,
                   double* r_squared, int* info, int* nfev) {
    return tanh_fit_host(x, y, n, a0, popt, resnorm, r_squared, info, nfev);
}

int esper_sweep_config(esper_ctx* c, int mode) {
    ESPER_ENTER(c);
    if (mode != 0 && mode != 1) return fail(c, ESPER_E_ARG, "sweep_config: mode must be 0 (general) or 1 (moments)");
    c->sweep_mode = mode;
    return ESPER_OK;
}

int esper_sweep_info(esper_ctx* c, int* last_path, int* bins) {
    ESPER_ENTER(c);
    if (last_path) *last_path = c->sweep_last_path;
    if (bins) *bins = c->sweep_last_bins;
    return ESPER_OK;
}

// ---- stage 3 -----------------------------------------------------------------------------------
int esper_markov(esper_ctx* c, const double* D, int m, double eps, double* degree, double* Ms) {
    ESPER_ENTER(c);
    if (!(eps > 0.0)) return fail(c, ESPER_E_ARG, "markov: eps must be > 0");
    int rc = need_dist(c, D, m, "markov");
    if (rc) return rc;
    rc = run_markov(c, m, eps);
    if (rc) return rc;
    if (degree) ESPER_CUDA(c, cudaMemcpyAsync(degree, c->degree.p, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    if (Ms) ESPER_CUDA(c, cudaMemcpyAsync(Ms, c->markov.p, sizeof(double) * (size_t)m * m, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

int esper_dmap_config(esper_ctx* c, int mode) {
    ESPER_ENTER(c);
    if (mode < 0 || mode > 3) return fail(c, ESPER_E_ARG, "dmap_config: mode must be 0 (barrier-free chain kernel), 1 (streamed barrier kernel, two grids per SM), 2 (the same without the speculative batch) or 3 (resident barrier kernel)");
    c->lanczos_mode = mode;
    return ESPER_OK;
}

int esper_dmap_embed(esper_ctx* c, const double* D, int m, double eps, int k, double tol, int max_iter,
                     double* val, double* vec, int* iters) {
    ESPER_ENTER(c);
    NvtxScope nvtx_("esper_dmap_embed");
    if (!(eps > 0.0)) return fail(c, ESPER_E_ARG, "dmap_embed: eps must be > 0");
    if (k < 1 || k > m) return fail(c, ESPER_E_ARG, "dmap_embed: need 1 <= k <= m (k=%d, m=%d)", k, m);
    int rc = need_dist(c, D, m, "dmap_embed");
    if (rc) return rc;
    if (c->dist_symmetric < 0 && (rc = check_symmetry(c, m))) return rc;
    if (c->dist_symmetric == 0)
        return fail(c, ESPER_E_ARG, "dmap_embed: the distance matrix is not symmetric (the Lanczos solver needs Ms = Ms^T)");
    rc = run_markov(c, m, eps);
    if (rc) return rc;
    return run_lanczos(c, m, k, tol, max_iter, val, vec, iters);
}

// ---- stages 1-3 in one call -------------------------------------------------------------------------
int esper_pd_embed_fused(esper_ctx* c, const float* imgs, int n, int P, unsigned flags, const double* logEps,
                         const double* coef, int n_eps, const double* a0, int k, double tol, int max_iter,
                         double* D, double* logSum, double* fit, double* val, double* vec, int* iters) {
    ESPER_ENTER(c);
    NvtxScope nvtx_("esper_pd_embed_fused");
    if (!logEps || !coef || !a0 || !logSum || !fit || !val || !vec || n_eps < 4)
        return fail(c, ESPER_E_ARG, "pd_embed_fused: logEps, coef, a0, logSum, fit, val, vec must not be NULL and n_eps >= 4");
    if (k < 1 || k > n) return fail(c, ESPER_E_ARG, "pd_embed_fused: need 1 <= k <= n (k=%d, n=%d)", k, n);
    int rc = esper_dist_rms(c, imgs, n, P, flags, nullptr);
    if (rc) return rc;
    if (D) {                                               // 8 n^2 bytes leave on a second stream while stages 2-3 run
        if (!c->copy_stream) {
            ESPER_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            ESPER_CUDA(c, cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming));
        }
        ESPER_CUDA(c, cudaEventRecord(c->copy_event, c->stream));
        ESPER_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->copy_event, 0));
        ESPER_CUDA(c, cudaMemcpyAsync(D, c->dist.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    auto finish_copy = [&]() -> int {
        if (D) ESPER_CUDA(c, cudaStreamSynchronize(c->copy_stream));
        return ESPER_OK;
    };
    rc = run_sweep(c, n, coef, n_eps, 10.0, logSum);
    if (rc) { finish_copy(); return rc; }
    double resnorm = INFINITY, r2 = 0.0;
    int info = 0, nfev = 0;
    rc = tanh_fit_host(logEps, logSum, n_eps, a0, fit, &resnorm, &r2, &info, &nfev);
    if (rc) { finish_copy(); return fail(c, ESPER_E_ARG, "pd_embed_fused: the sweep produced non-finite values (tanh fit impossible)"); }
    fit[4] = resnorm; fit[5] = r2; fit[6] = 0.0; fit[7] = fit[0] * fit[2];
    if (!(resnorm <= 100.0)) {
        rc = finish_copy();
        if (rc) return rc;
        return fail(c, ESPER_E_NOFIT, "pd_embed_fused: tanh fit from this start has resnorm %g > 100 (MINPACK code %d); retry with a new start", resnorm, info);
    }
    const double eps = exp(-(fit[1] / fit[0]));
    fit[6] = eps;
    if (!(eps > 0.0) || !(eps < INFINITY)) { finish_copy(); return fail(c, ESPER_E_NOFIT, "pd_embed_fused: epsilon %g from the fit is unusable", eps); }
    rc = run_markov(c, n, eps);
    if (!rc) rc = run_lanczos(c, n, k, tol, max_iter, val, vec, iters);
    const int rc2 = finish_copy();
    return rc ? rc : rc2;
}

// float32 mean and population std per image with numpy's summation order (the statistics of the CTF branch)
int esper_image_stats(esper_ctx* c, const float* imgs, int n, int P, float* mean, float* stdv) {
    ESPER_ENTER(c);
    if (n < 1 || P < 1 || !mean || !stdv) return fail(c, ESPER_E_ARG, "image_stats: need n >= 1, P >= 1 and output arrays");
    if (imgs) {
        int rc = esper_upload_images(c, imgs, n, P);
        if (rc) return rc;
    } else if (c->img_n != n || c->img_P != P || !c->imgs.p) {
        return fail(c, ESPER_E_STATE, "image_stats: imgs == NULL but no resident [%d, %d] stack", n, P);
    }
    ESPER_RESERVE(c, c->row_stats, sizeof(esper::RowStat) * esper::round_up((size_t)n, 128));
    c->launches++;
    int rc = esper::row_stats_numpy(c, n, P, c->row_stats.as<esper::RowStat>());
    if (rc) return rc;
    std::vector<esper::RowStat> h((size_t)n);
    ESPER_CUDA(c, cudaMemcpyAsync(h.data(), c->row_stats.p, sizeof(esper::RowStat) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) { mean[i] = h[(size_t)i].mean; stdv[i] = h[(size_t)i].stdv; }
    return ESPER_OK;
}

// ---- CTF branch of stage 1 ----------------------------------------------------------
int esper_ctf_dist(esper_ctx* c, const float* imgs, int n, int box, const double* params, unsigned flags,
                   double* D, float* wiener) {
    ESPER_ENTER(c);
    if (n < 1 || box < 2) return fail(c, ESPER_E_ARG, "ctf_dist: need n >= 1 and box >= 2 (got n=%d, box=%d)", n, box);
    if (!params) return fail(c, ESPER_E_ARG, "ctf_dist: params == NULL (need [n, 5] = pixel size, Cs, defocus, voltage, amplitude contrast)");
    const int P = box * box;
    if (imgs) {
        int rc = esper_upload_images(c, imgs, n, P);
        if (rc) return rc;
    } else if (c->img_n != n || c->img_P != P || !c->imgs.p) {
        return fail(c, ESPER_E_STATE, "ctf_dist: imgs == NULL but no resident [%d, %d] stack", n, P);
    }
    c->dist_m = 0;
    int rc = run_ctf_dist(c, n, box, params, flags, wiener != nullptr);
    if (rc) return rc;
    c->dist_m = n; c->dist_rows = 0; c->dist_row0 = -1;
    c->dist_symmetric = 1;
    if (D) ESPER_CUDA(c, cudaMemcpyAsync(D, c->dist.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, c->stream));
    if (wiener) ESPER_CUDA(c, cudaMemcpyAsync(wiener, c->ctf_wiener.p, sizeof(float) * (size_t)n * P, cudaMemcpyDeviceToHost, c->stream));
    if (D || wiener) ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

// ---- PCA embedding -----------------------------------------------------------------
__global__ void fill_ones_kernel(double* x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = 1.0;
}

int esper_pca_embed(esper_ctx* c, const float* imgs, int n, int P, int k, double tol, int max_iter,
                    double* val, double* U, int* iters) {
    ESPER_ENTER(c);
    if (n < 2 || P < 1) return fail(c, ESPER_E_ARG, "pca_embed: need n >= 2 and P >= 1 (got n=%d, P=%d)", n, P);
    if (k < 1 || k > n - 1) return fail(c, ESPER_E_ARG, "pca_embed: need 1 <= k <= n-1 (the centred stack has rank <= n-1; k=%d, n=%d)", k, n);
    if (imgs) {
        int rc = esper_upload_images(c, imgs, n, P);
        if (rc) return rc;
    } else if (c->img_n != n || c->img_P != P || !c->imgs.p) {
        return fail(c, ESPER_E_STATE, "pca_embed: imgs == NULL but no resident [%d, %d] stack", n, P);
    }
    const int rows_pad = (int)round_up((size_t)n, 128);
    const int ld = (int)round_up((size_t)P, 32);
    ESPER_RESERVE(c, c->markov, sizeof(double) * (size_t)n * n);
    ESPER_RESERVE(c, c->degree, sizeof(double) * (size_t)n);
    c->dist_m = 0;                                         // the plane buffers are reused: no distance matrix is resident
    int rc = run_prep(c, n, P, ESPER_DIST_CENTER, rows_pad, ld, /*exact_center=*/true);
    if (rc) return rc;
    rc = run_gram(c, n, P, rows_pad, ld, 0, -1, 0, c->markov.as<double>());
    if (rc) return rc;
    // the mean image was subtracted, so G 1 = 0: lock the constant vector (sqrt of an all-ones "degree")
    fill_ones_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->degree.as<double>(), n);
    c->launches++;
    ESPER_CUDA(c, cudaGetLastError());
    rc = run_lanczos(c, n, k, tol, max_iter, val, U, iters, /*plain=*/true);
    if (rc) return rc;
    if (U && val)                                          // scores Y^T W = V S (PCA.py:68): scale unit eigenvectors by the singular value
        for (int j = 0; j < k; ++j) {
            const double sv = val[j] > 0.0 ? sqrt(val[j]) : 0.0;
            for (int i = 0; i < n; ++i) U[(size_t)i * k + j] *= sv;
        }
    return ESPER_OK;
}

// ---- oversized-PD row-block path ----------------------------------------------------------------
int esper_dist_rms_rows(esper_ctx* c, const float* imgs, int n, int P, unsigned flags, int row0, int nrows, double* D_rows) {
    ESPER_ENTER(c);
    if (n < 1 || P < 1) return fail(c, ESPER_E_ARG, "dist_rms_rows: need n >= 1 and P >= 1");
    if (row0 < 0 || nrows < 1 || row0 + nrows > n || (row0 % 128) != 0)
        return fail(c, ESPER_E_ARG, "dist_rms_rows: need 0 <= row0 (multiple of 128), nrows >= 1, row0 + nrows <= n");
    // the epilogue stores whole 128-row tiles guarded by the matrix order only: a short block must be the last one
    if ((nrows % 128) != 0 && row0 + nrows != n)
        return fail(c, ESPER_E_ARG, "dist_rms_rows: nrows must be a multiple of 128 unless the block ends at row n (row0=%d, nrows=%d, n=%d)",
                    row0, nrows, n);
    if (imgs) {
        int rc = esper_upload_images(c, imgs, n, P);
        if (rc) return rc;
    } else if (c->img_n != n || c->img_P != P || !c->imgs.p) {
        return fail(c, ESPER_E_STATE, "dist_rms_rows: imgs == NULL but no resident [%d, %d] stack", n, P);
    }
    const int rows_pad = (int)round_up((size_t)n, 128);
    int hs = -1;                                           // FP16 operand planes: z-scored stacks only (see esper_dist_rms)
    if (c->gram_mode == 1 && (flags & ESPER_DIST_STANDARDIZE)) hs = gram_f16_scale_log2(P);
    const int ld = (int)round_up((size_t)P, hs >= 0 ? 64 : 32);
    ESPER_RESERVE(c, c->dist, sizeof(double) * (size_t)nrows * n);
    c->dist_m = 0;
    int rc = run_prep(c, n, P, flags, rows_pad, ld, false, hs);
    if (rc) return rc;
    rc = run_gram(c, n, P, rows_pad, ld, flags, row0, nrows, nullptr, hs);
    if (rc) return rc;
    c->gram_last_f16 = hs >= 0 ? 1 : 0;
    c->dist_m = n; c->dist_rows = nrows; c->dist_row0 = row0; c->dist_symmetric = 0;
    if (D_rows) {
        ESPER_CUDA(c, cudaMemcpyAsync(D_rows, c->dist.p, sizeof(double) * (size_t)nrows * n, cudaMemcpyDeviceToHost, c->stream));
        ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return ESPER_OK;
}

static int need_rows(esper_ctx* c, int nrows, int n, const char* who) {
    if (c->dist_row0 < 0 || c->dist_rows != nrows || c->dist_m != n || !c->dist.p)
        return fail(c, ESPER_E_STATE, "%s: no resident [%d, %d] distance row block (call esper_dist_rms_rows first)", who, nrows, n);
    return ESPER_OK;
}

int esper_moment_layout(int* out5) {
    if (!out5) return ESPER_E_ARG;
    moment_layout_host(out5);
    return ESPER_OK;
}

int esper_moment_stats_rows(esper_ctx* c, int nrows, int n, double* stats4) {
    ESPER_ENTER(c);
    if (!stats4) return fail(c, ESPER_E_ARG, "moment_stats_rows: stats4 is NULL");
    int rc = need_rows(c, nrows, n, "moment_stats_rows");
    if (rc) return rc;
    return run_moment_stats_rows(c, nrows, n, stats4);
}

int esper_moment_accum_rows(esper_ctx* c, int nrows, int n, const double* coef, int n_eps, double thr, const double* plan, double* table) {
    ESPER_ENTER(c);
    if (!coef || !plan || !table || n_eps < 1) return fail(c, ESPER_E_ARG, "moment_accum_rows: need coef, plan, table != NULL and n_eps >= 1");
    int rc = need_rows(c, nrows, n, "moment_accum_rows");
    if (rc) return rc;
    return run_moment_accum_rows(c, nrows, n, coef, n_eps, thr, plan, table);
}

int esper_eps_sweep_rows(esper_ctx* c, int nrows, int n, const double* coef, int n_eps, double thr, double* sums) {
    ESPER_ENTER(c);
    if (!coef || !sums || n_eps < 1) return fail(c, ESPER_E_ARG, "eps_sweep_rows: need coef, sums != NULL and n_eps >= 1");
    int rc = need_rows(c, nrows, n, "eps_sweep_rows");
    if (rc) return rc;
    return run_sweep_rows(c, nrows, n, coef, n_eps, thr, sums, 0);
}

int esper_degree_rows(esper_ctx* c, int nrows, int n, double eps, double* degree_rows) {
    ESPER_ENTER(c);
    if (!(eps > 0.0) || !degree_rows) return fail(c, ESPER_E_ARG, "degree_rows: need eps > 0 and degree_rows != NULL");
    int rc = need_rows(c, nrows, n, "degree_rows");
    if (rc) return rc;
    return run_markov_rows(c, nrows, n, c->dist_row0, eps, nullptr, degree_rows, 0);
}

int esper_markov_rows(esper_ctx* c, int nrows, int n, double eps, const double* degree_all) {
    ESPER_ENTER(c);
    if (!(eps > 0.0) || !degree_all) return fail(c, ESPER_E_ARG, "markov_rows: need eps > 0 and degree_all != NULL");
    int rc = need_rows(c, nrows, n, "markov_rows");
    if (rc) return rc;
    return run_markov_rows(c, nrows, n, c->dist_row0, eps, degree_all, nullptr, 1);
}

int esper_symv_rows_d(esper_ctx* c, int nrows, int n, const double* q_d, double* w_rows_d) {
    ESPER_ENTER(c);
    if (!q_d || !w_rows_d || nrows < 1 || n < 1 || !c->markov.p) return fail(c, ESPER_E_ARG, "symv_rows_d: bad argument or no resident Ms block");
    return run_symv_rows(c, nrows, n, q_d, w_rows_d);
}

// ---- fused SYMV + all-gather over peer memory ---------------------------------------------------------------------
int esper_comm_alloc(esper_ctx* c, int n_pad, int world, void* ipc_handle) {
    ESPER_ENTER(c);
    if (!ipc_handle) return fail(c, ESPER_E_ARG, "comm_alloc: ipc_handle is NULL");
    run_comm_close(c);
    return run_comm_alloc(c, n_pad, world, ipc_handle);
}

int esper_comm_open(esper_ctx* c, const void* ipc_handles, int rank, int world) {
    ESPER_ENTER(c);
    if (!ipc_handles) return fail(c, ESPER_E_ARG, "comm_open: ipc_handles is NULL");
    return run_comm_open(c, ipc_handles, rank, world);
}

int esper_comm_close(esper_ctx* c) {
    ESPER_ENTER(c);
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return run_comm_close(c);
}

int esper_symv_allgather_d(esper_ctx* c, int nrows, int n, int row0, const double* q_d, int step) {
    ESPER_ENTER(c);
    if (!q_d || nrows < 1 || n < 1 || !c->markov.p) return fail(c, ESPER_E_ARG, "symv_allgather_d: bad argument or no resident Ms block");
    return run_symv_allgather(c, nrows, n, row0, q_d, step);
}

int esper_comm_wait_d(esper_ctx* c, int step, int n, double** w_full_d) {
    ESPER_ENTER(c);
    return run_comm_wait(c, step, n, w_full_d);
}

int esper_lanczos_fused_steps_d(esper_ctx* c, int nrows, int n, int row0, double* V_d, double* ab_d, int j0, int n_steps, int step0) {
    ESPER_ENTER(c);
    if (!V_d || !ab_d || nrows < 1 || n < 2 || j0 < 1 || n_steps < 1 || step0 < 1 || !c->markov.p)
        return fail(c, ESPER_E_ARG, "lanczos_fused_steps_d: bad argument or no resident Ms block");
    for (int i = 0; i < n_steps; ++i) {
        const int j = j0 + i;
        int rc = run_symv_allgather(c, nrows, n, row0, V_d + (size_t)j * n, step0 + i);
        if (rc) return rc;
        double* w = nullptr;
        rc = run_comm_wait(c, step0 + i, n, &w);
        if (rc) return rc;
        rc = run_lanczos_ortho(c, n, j, V_d, w, ab_d + 2 * (size_t)(j - 1));
        if (rc) return rc;
    }
    return ESPER_OK;
}

int esper_comm_status(esper_ctx* c, unsigned int* status) {
    ESPER_ENTER(c);
    if (!status) return fail(c, ESPER_E_ARG, "comm_status: status is NULL");
    return run_comm_status(c, status);
}

int esper_lanczos_start_d(esper_ctx* c, int n, const double* degree_all, double* V_d) {
    ESPER_ENTER(c);
    if (!degree_all || !V_d || n < 2) return fail(c, ESPER_E_ARG, "lanczos_start_d: bad argument");
    return run_lanczos_start(c, n, degree_all, V_d);
}

int esper_lanczos_ortho_d(esper_ctx* c, int n, int j, double* V_d, double* w_d, double* alpha_beta_d) {
    ESPER_ENTER(c);
    if (!V_d || !w_d || !alpha_beta_d || n < 2 || j < 1) return fail(c, ESPER_E_ARG, "lanczos_ortho_d: bad argument");
    return run_lanczos_ortho(c, n, j, V_d, w_d, alpha_beta_d);
}

int esper_tridiag_eigen_all(const double* alpha, const double* beta, int n, double* theta, double* Z) {
    if (!alpha || !beta || !theta || !Z || n < 1) return ESPER_E_ARG;
    return tridiag_eigen_all_host(alpha, beta, n, theta, Z);
}

int esper_tridiag_ritz(const double* alpha, const double* beta, int steps, int k, double tol, int* converged,
                       double* theta, double* S) {
    if (!alpha || !beta || steps < 1 || k < 2 || !converged) return ESPER_E_ARG;
    if (tol <= 0.0) tol = ESPER_DEFAULT_EIG_TOL;
    const int kw = k - 1;
    std::vector<double> a(alpha, alpha + steps), b(beta, beta + steps), th, Sv;
    const int guard = 2;
    const bool ok = ritz_converged(a, b, steps, kw, guard, tol, th, Sv);
    const int want = (int)th.size();
    *converged = (ok && want >= kw) ? 1 : 0;
    if (theta) for (int i = 0; i < kw; ++i) theta[i] = i < want ? th[i] : 0.0;
    if (S) for (int l = 0; l < steps; ++l) for (int i = 0; i < kw; ++i) S[(size_t)l * kw + i] = i < want ? Sv[(size_t)l * want + i] : 0.0;
    return ESPER_OK;
}

int esper_ritz_vectors_d(esper_ctx* c, int n, int steps, int k, const double* V_d, const double* S, double* vec) {
    ESPER_ENTER(c);
    if (!V_d || !vec || n < 1 || k < 1 || (k > 1 && (!S || steps < 1))) return fail(c, ESPER_E_ARG, "ritz_vectors_d: bad argument");
    return run_ritz_vectors(c, n, steps, k, V_d, S, vec);
}

// ---- stage 4 -----------------------------------------------------------------------------------
int esper_nd_rotation(int n, const double* theta, int n_theta, double* R) {
    if (n < 1 || !R || (n > 1 && !theta)) return ESPER_E_ARG;
    if (n_theta != n * (n - 1) / 2) return ESPER_E_ARG;
    nd_rotation_host(n, theta, R);
    return ESPER_OK;
}

int esper_hist_edges(int bins, double lo, double hi, double* edges) {
    if (bins < 1 || !edges) return ESPER_E_ARG;
    hist_edges_host(bins, lo, hi, edges);
    return ESPER_OK;
}

int esper_rot_hist_eval(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, const double* R,
                        const int* pd_of_eval, const int* v1, const int* v2, int n_eval, int bins,
                        double lo, double hi, int* nonzero, int* H) {
    ESPER_ENTER(c);
    if (!Uinit || !R || !pd_of_eval || !v1 || !v2 || !nonzero || n_pd < 1 || m < 1 || n_eval < 1)
        return fail(c, ESPER_E_ARG, "rot_hist_eval: NULL argument or empty batch");
    return run_rot_eval(c, Uinit, n_pd, m, dim, R, pd_of_eval, v1, v2, n_eval, bins, lo, hi, nonzero, H);
}

int esper_rot_sweep(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, const int* combo,
                    const int* n_sub, int max_sub, const int* rij, int n_rij, const double* theta_grid,
                    int n_theta, int bins, double lo, double hi, double* thetas_out, int* counts_out) {
    ESPER_ENTER(c);
    NvtxScope nvtx_("esper_rot_sweep");
    if (!combo || !n_sub || (!rij && n_rij > 0) || !theta_grid || !thetas_out || n_pd < 1 || m < 1)
        return fail(c, ESPER_E_ARG, "rot_sweep: NULL argument or empty batch");
    if (!Uinit && (c->rot_n_pd != n_pd || c->rot_m != m || c->rot_dim != dim))
        return fail(c, ESPER_E_STATE, "rot_sweep: Uinit == NULL but no resident [%d, %d, %d] manifolds (resident: [%d, %d, %d])",
                    n_pd, m, dim, c->rot_n_pd, c->rot_m, c->rot_dim);
    return run_rot_sweep(c, Uinit, n_pd, m, dim, combo, n_sub, max_sub, rij, n_rij, theta_grid, n_theta, bins, lo,
                         hi, thetas_out, counts_out);
}

int esper_manifold_prepare(esper_ctx* c, const double* U0, int n_pd, int m, int ncol, int first, int dim, int kind,
                           double* Uinit, double* R2, double* coef) {
    ESPER_ENTER(c);
    if (!U0 || (kind != 0 && !R2)) return fail(c, ESPER_E_ARG, "manifold_prepare: U0 (and R2 when kind != 0) must not be NULL");
    c->rot_n_pd = 0;
    int rc = run_manifold_prepare(c, U0, n_pd, m, ncol, first, dim, kind, Uinit, R2, coef);
    if (rc == ESPER_OK) { c->rot_n_pd = n_pd; c->rot_m = m; c->rot_dim = dim; }
    return rc;
}

int esper_conic_fit_batch(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, int kind, double* R2, double* coef) {
    ESPER_ENTER(c);
    if (!Uinit && (c->rot_n_pd != n_pd || c->rot_m != m || c->rot_dim != dim))
        return fail(c, ESPER_E_STATE, "conic_fit_batch: Uinit == NULL but no resident [%d, %d, %d] manifolds", n_pd, m, dim);
    if (Uinit) c->rot_n_pd = 0;
    int rc = run_conic_fits(c, Uinit, n_pd, m, dim, kind, R2, coef);
    if (rc == ESPER_OK && Uinit) { c->rot_n_pd = n_pd; c->rot_m = m; c->rot_dim = dim; }
    return rc;
}

// ---- data preparation and formats ---------------------------------------------------
int esper_tau_snr_stack(esper_ctx* c, const float* pristine, int ss, int box, int tau, double snr, unsigned long long seed,
                        double* params, float* stack) {
    ESPER_ENTER(c);
    if (!pristine) return fail(c, ESPER_E_ARG, "tau_snr_stack: pristine is NULL");
    return run_tau_snr_stack(c, pristine, ss, box, tau, snr, seed, params, stack);
}

int esper_download_images(esper_ctx* c, float* imgs, int n, int P) {
    ESPER_ENTER(c);
    if (!imgs) return fail(c, ESPER_E_ARG, "download_images: imgs is NULL");
    if (c->img_n != n || c->img_P != P || !c->imgs.p)
        return fail(c, ESPER_E_STATE, "download_images: no resident [%d, %d] stack (resident: [%d, %d])", n, P, c->img_n, c->img_P);
    ESPER_CUDA(c, cudaMemcpyAsync(imgs, c->imgs.p, sizeof(float) * (size_t)n * P, cudaMemcpyDeviceToHost, c->stream));
    ESPER_CUDA(c, cudaStreamSynchronize(c->stream));
    return ESPER_OK;
}

int esper_mrc_info(const char* path, int* dims, int* mode, long long* data_offset) {
    if (!path || !dims || !mode || !data_offset) { g_create_error = "mrc_info: NULL argument"; return ESPER_E_ARG; }
    char err[400];
    int rc = mrc_info_host(path, dims, mode, data_offset, err, sizeof err);
    if (rc) g_create_error = err;
    return rc;
}

int esper_upload_mrc(esper_ctx* c, const char* path, int first, int count, int* dims) {
    ESPER_ENTER(c);
    if (!path) return fail(c, ESPER_E_ARG, "upload_mrc: path is NULL");
    return run_upload_mrc(c, path, first, count, dims);
}

int esper_bin_accumulate(esper_ctx* c, const float* imgs, int n, int P, const int* order, const int* offsets, int n_bins, double* sums) {
    ESPER_ENTER(c);
    if (!order || !offsets || !sums || n < 1 || P < 1) return fail(c, ESPER_E_ARG, "bin_accumulate: NULL argument or empty stack");
    if (imgs) {
        int rc = esper_upload_images(c, imgs, n, P);
        if (rc) return rc;
    } else if (c->img_n != n || c->img_P != P || !c->imgs.p) {
        return fail(c, ESPER_E_STATE, "bin_accumulate: imgs == NULL but no resident [%d, %d] stack (resident: [%d, %d])", n, P, c->img_n, c->img_P);
    }
    return run_bin_accumulate(c, n, P, order, offsets, n_bins, sums);
}

}  // extern "C"
