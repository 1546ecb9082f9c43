// common.cuh -- context, error handling, device-buffer and profiling plumbing shared by the kernels.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/esper_b200.h"

namespace esper {

// A growable device buffer owned by the ctx.
struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct KernelTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double total_ms = 0.0;
    int launches = 0;
};

}  // namespace esper

struct esper_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;                        // device-to-host copies that overlap later stages (esper_pd_embed_fused)
    cudaEvent_t copy_event = nullptr;
    std::string err;

    // profiling
    bool profiling = false;
    bool profiling_gram_only = false;                          // esper_set_profiling(ctx, 2): time the stage-1 Gram launches only
    std::map<std::string, esper::KernelTimer> timers;
    std::vector<cudaEvent_t> event_pool;
    long long launches = 0;

    bool attr_gram_wide = false, attr_gram = false, attr_sweep = false, attr_lanczos_pro = false, attr_lanczos = false;
    long long lanczos_smem_set = 0;                            // dynamic shared memory opted in for the resident kernel
    int lanczos_mode = 0;                                      // 0 latency (Ms resident on chip), 1 throughput (streamed, two grids per SM)
    int lanczos_fallbacks = 0;                                 // PRO runs repeated with full reorthogonalisation
    // stage-1 configuration
    int split_k = 0;
    int drain_kb = 2;                                          // k-blocks (of 32 pixels) per tensor-core accumulation chain
    double refine_tau = 0.05;
    int gram_mode = 1, gram_last_f16 = 0;                      // esper_gram_config: 0 = 3xTF32 operand planes, 1 = 3xFP16 (default)

    // resident data
    esper::DevBuf imgs;     int img_n = 0, img_P = 0;          // raw float32 stack [n, P]
    esper::DevBuf dist;     int dist_m = 0;                    // float64 [m, m] (or the row block [dist_rows, m])
    int dist_rows = 0, dist_row0 = -1;                         // row-block path: rows held, first global row
    int dist_symmetric = -1;                                   // 1 symmetric, 0 not, -1 unknown (checked lazily)
    // stage-1 workspaces
    esper::DevBuf plane_hi, plane_lo;                          // [rows_pad, ld] float32 (tf32 values)
    esper::DevBuf row_stats;                                   // per image: mean32, std32 (float2) + norm (double)
    esper::DevBuf col_part;                                    // column-mean partials
    esper::DevBuf gram_part;                                   // split-K partial tiles
    esper::DevBuf pair_list;                                   // refine: flagged pairs + counter
    // stage-2/3 workspaces
    esper::DevBuf sweep_part, sweep_aux;
    int sweep_mode = 1, sweep_last_path = 0, sweep_last_bins = 0;   // esper_sweep_config / esper_sweep_info
    esper::DevBuf markov, degree;                              // Ms [m,m], d [m]
    esper::DevBuf ctf_par, ctf_T, ctf_fy, ctf_C, ctf_ab, ctf_T12, ctf_wiener;   // CTF branch (k6_ctf.cu)
    bool attr_ctf = false;
    esper::DevBuf pw_plan; int pw_P = -1, pw_L = 0, pw_nodes = 0, pw_levels = 0, pw_root = 0;   // numpy pairwise-sum tree (k0_prep.cu)
    esper::DevBuf lan_V, lan_w, lan_h, lan_small;              // Lanczos basis etc.
    esper::DevBuf chain_ws; esper::PinBuf pin_chain; long long chain_smem_set = 0;   // barrier-free schedule (k4_chain.cu)
    esper::DevBuf chain_trace; int chain_trace_ctas = 0;      // diagnostic timeline of one step (ESPER_DEBUG_LANCZOS)
    int cluster_cap_size = -1, cluster_cap = 0, cluster_launches = 0; long long cluster_smem = 0;   // its cluster variant: resident clusters of that size
    int merged_cap_size = -1, merged_cap = 0, merged_launches = 0; long long merged_smem = 0; bool chain_trace_merged = false;   // merged (one-exchange) schedule
    // stage-4 workspaces
    esper::DevBuf rot_U, rot_R, rot_idx, rot_out, rot_H, rot_theta, rot_aux;
    int rot_n_pd = 0, rot_m = 0, rot_dim = 0;                  // shape of the normalised manifolds resident in rot_U
    // row-block path: peer-memory communicator (k4_eigen.cu)
    esper::DevBuf comm_buf;
    void* comm_peer[16] = {nullptr};
    int comm_n_pad = 0, comm_world = 0, comm_rank = 0;
    bool comm_open = false;
    // data preparation / formats (k8_dataprep.cu)
    esper::DevBuf prep_src, prep_out;                          // pristine states + SNR table; bin sums
    esper::PinBuf pin_stream[2];                               // page-locked double buffer of the MRC reader
    esper::PinBuf pin;                                         // small pinned staging area
};

namespace esper {

inline int fail(esper_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define ESPER_CUDA(c, call)                                                                    \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            cudaGetLastError();                                                                \
            return esper::fail((c), ESPER_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                               cudaGetErrorString(e__));                                       \
        }                                                                                      \
    } while (0)

#define ESPER_RESERVE(c, buf, bytes)                                                           \
    do {                                                                                       \
        if ((buf).reserve(bytes) != 0)                                                         \
            return esper::fail((c), ESPER_E_NOMEM, "device allocation of %zu bytes failed (%s)", \
                               (size_t)(bytes), #buf);                                         \
    } while (0)

// NVTX range around one stage of the path (tracing); visible in Nsight Systems / Compute timelines.
struct NvtxScope {
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
};

// Scoped CUDA-event timer around one kernel launch (only when profiling is on).
struct LaunchScope {
    esper_ctx* c;
    cudaEvent_t a = nullptr, b = nullptr;
    const char* name;
    LaunchScope(esper_ctx* ctx, const char* nm) : c(ctx), name(nm) {
        c->launches++;
        if (!c->profiling) return;
        if (c->profiling_gram_only && strcmp(nm, "gram") != 0) return;
        a = get(); b = get();
        cudaEventRecord(a, c->stream);
    }
    ~LaunchScope() {
        if (!a) return;
        cudaEventRecord(b, c->stream);
        c->timers[name].pending.emplace_back(a, b);
    }
    cudaEvent_t get() {
        if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

struct RowStat;
// ---- per-stage entry points implemented in the .cu files (host side) ----
int prep_and_gram(esper_ctx* c, int n, int P, unsigned flags);                 // k0_prep.cu + k1_gram.cu
int run_prep(esper_ctx* c, int n, int P, unsigned flags, int rows_pad, int ld, bool exact_center = false, int half_scale_log2 = -1, bool numpy_stats = false);
int run_gram(esper_ctx* c, int n, int P, int rows_pad, int ld, unsigned flags, int row0 = -1, int nrows = 0, double* gram_out = nullptr,
             int half_scale_log2 = -1);
int gram_f16_scale_log2(int P);
int run_sweep(esper_ctx* c, int m, const double* coef_h, int n_eps, double thr, double* logsum_h);
int run_sweep_rows(esper_ctx* c, int rows, int m, const double* coef_h, int n_eps, double thr, double* out_h, int take_log);
int run_markov_rows(esper_ctx* c, int rows, int m, int row0, double eps, const double* degree_all_h, double* degree_rows_h, int phase);
bool ritz_converged(const std::vector<double>& a, const std::vector<double>& b, int n, int kw, int guard, double tol,
                    std::vector<double>& theta, std::vector<double>& S);
int check_symmetry(esper_ctx* c, int m);
int run_moment_stats_rows(esper_ctx* c, int rows, int m, double* out4_h);
int run_moment_accum_rows(esper_ctx* c, int rows, int m, const double* coef_h, int n_eps, double thr, const double* plan_h, double* table_h);
void moment_layout_host(int* out5);
int run_product(esper_ctx* c, float* a_hi, float* a_lo, int na, float* b_hi, float* b_lo, int nb, int ld, double* M);
int run_ctf_dist(esper_ctx* c, int n, int box, const double* params_h, unsigned flags, bool want_wiener);
int row_stats_numpy(esper_ctx* c, int n, int P, RowStat* st, int centered_std = 0);
int run_markov(esper_ctx* c, int m, double eps);
int run_lanczos(esper_ctx* c, int m, int k, double tol, int max_iter, double* val_h, double* vec_h, int* iters, bool plain = false);
int run_rot_eval(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, const double* R,
                 const int* pd_of_eval, const int* v1, const int* v2, int n_eval, int bins,
                 double lo, double hi, int* nonzero, int* H);
int run_rot_sweep(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, const int* combo,
                  const int* n_sub, int max_sub, const int* rij, int n_rij, const double* theta_grid,
                  int n_theta, int bins, double lo, double hi, double* thetas_out, int* counts_out);
int finish_full_spectrum(esper_ctx* c, int m, int k, int steps, const std::vector<double>& a, const std::vector<double>& b, double* V,
                         double** U_out, std::vector<double>& theta_out);
int tridiag_eigen_all_host(const double* alpha, const double* beta, int n, double* theta, double* Z);
int run_lanczos_chain(esper_ctx* c, int m, int k, int guard, double tol, int max_iter, double* V, double* U_d, bool unit_locked);
int chain_result(esper_ctx* c, int k, double* theta_out, int* iters_out);
int run_symv_rows(esper_ctx* c, int rows, int m, const double* q_d, double* w_rows_d);
int run_comm_alloc(esper_ctx* c, int n_pad, int world, void* handle_out);
int run_comm_open(esper_ctx* c, const void* handles, int rank, int world);
int run_comm_close(esper_ctx* c);
int run_symv_allgather(esper_ctx* c, int rows, int m, int row0, const double* q_d, int step);
int run_comm_wait(esper_ctx* c, int step, int n, double** w_full_d);
int run_comm_status(esper_ctx* c, unsigned int* status);
int run_lanczos_start(esper_ctx* c, int m, const double* degree_all_h, double* V_d);
int run_lanczos_ortho(esper_ctx* c, int m, int j, double* V_d, double* w_d, double* ab_d);
int run_ritz_vectors(esper_ctx* c, int m, int steps, int k, const double* V_d, const double* S_h, double* vec_h);
int run_manifold_prepare(esper_ctx* c, const double* U0, int n_pd, int m, int ncol, int first, int dim, int kind,
                         double* Uinit_h, double* R2_h, double* coef_h);
int run_conic_fits(esper_ctx* c, const double* Uinit, int n_pd, int m, int dim, int kind, double* R2_h, double* coef_h);
int run_tau_snr_stack(esper_ctx* c, const float* pristine_h, int ss, int box, int tau, double snr, unsigned long long seed,
                      double* params_h, float* stack_h);
int mrc_info_host(const char* path, int* dims, int* mode, long long* data_offset, char* err, size_t err_len);
int run_upload_mrc(esper_ctx* c, const char* path, int first, int count, int* dims_out);
int run_bin_accumulate(esper_ctx* c, int n, int P, const int* order_h, const int* offsets_h, int n_bins, double* sums_h);
int tanh_fit_host(const double* x, const double* y, int m, const double* a0, double* popt, double* resnorm, double* r_squared,
                  int* info_out, int* nfev_out);                        // k9_fit.cu
void nd_rotation_host(int n, const double* theta, double* R);
void hist_edges_host(int bins, double lo, double hi, double* edges);

}  // namespace esper
