/*
 * esper_b200.h -- C ABI of libesper_b200.so: the B200 (sm_100a) implementation of ESPER's
 * per-projection-direction (PD) diffusion-map embedding and Nd eigen-rotation histogram sweep.
 *
 * The reference (evanseitz/ManifoldEM_ESPER) is pure Python and has no FFI; the surface this
 * library sits behind is the reference's module-function surface.  Each entry point below
 * names the reference interface it replaces (path:line in the upstream tree).  The Python host
 * layer (the modules under manifoldem_esper_b200/) mirrors those module functions one-to-one and binds this
 * ABI with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every array is dense row-major; `double` = IEEE binary64.
 *   - pointers are HOST pointers unless the parameter name ends in `_d` (device pointer).
 *   - the caller owns every array it passes; outputs are caller-allocated; the library never
 *     frees caller memory.  An opaque `esper_ctx` owns device buffers, one stream, workspaces.
 *   - a NULL input pointer where documented means "use the copy the previous stage left on the
 *     device"; a NULL output pointer means "leave the result on the device only".  This lets the
 *     stages chain without PCIe round trips.
 *   - return value: 0 = ok; <0 = error (ESPER_E_*); message via esper_last_error().
 *   - one ctx per (process, GPU); calls on one ctx are serialised on its stream; ctxs are
 *     independent of each other.  There is NO CPU fallback: without a B200 the calls fail.
 */
#ifndef ESPER_B200_H
#define ESPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESPER_ABI_VERSION 1

#define ESPER_OK              0
#define ESPER_E_ARG          -1   /* bad argument (Python wrapper raises ValueError)        */
#define ESPER_E_CUDA         -2   /* CUDA runtime / driver error (RuntimeError)            */
#define ESPER_E_NOMEM        -3   /* device or host allocation failed                      */
#define ESPER_E_NOCONV       -4   /* eigensolver did not converge within max_iter          */
#define ESPER_E_STATE        -5   /* a NULL "use resident copy" argument with nothing resident */
#define ESPER_E_ARCH         -6   /* device is not sm_100 (no fallback path exists)        */
#define ESPER_E_NOFIT        -7   /* esper_pd_embed_fused: the tanh fit from this start is unusable (resnorm > 100); stages 1-2 are done */

typedef struct esper_ctx esper_ctx;

/* ---- context -------------------------------------------------------------------------- */
int         esper_abi_version(void);
/* device: CUDA ordinal.  stream: a cudaStream_t to launch on (e.g. torch's), or NULL to let the
 * ctx create its own non-blocking stream. */
esper_ctx*  esper_create(int device, void* stream);
void        esper_destroy(esper_ctx* ctx);
const char* esper_last_error(esper_ctx* ctx);          /* ctx may be NULL: last create error */
int         esper_sync(esper_ctx* ctx);                 /* cudaStreamSynchronize              */
void*       esper_stream(esper_ctx* ctx);               /* the cudaStream_t every call is enqueued on (callers order their own work against it) */
/* Per-kernel CUDA-event timing of the calls that follow (off by default).  on: 0 off, 1 every launch, 2 the stage-1 Gram
 * launches only (two events per PD: cheap enough to leave on inside a timed region). */
int         esper_set_profiling(esper_ctx* ctx, int on);
/* Sum / count of the kernel launches recorded under `name` since the last reset.
 * names: "prep", "gram", "dist_finalize", "refine", "sweep", "markov", "lanczos", "rot", "conic", "dataprep".    */
int         esper_get_kernel_ms(esper_ctx* ctx, const char* name, double* total_ms, int* launches);
int         esper_reset_profiling(esper_ctx* ctx);
/* Total kernels launched by this ctx since creation (the bench's `gpu_launches` claim). */
long long   esper_launch_count(esper_ctx* ctx);
/* Counters: "launches", "lanczos_fallbacks" (partial-reorthogonalisation runs that failed their independent
 * residual / orthogonality check and were repeated with full reorthogonalisation).  -1 = unknown name. */
long long   esper_stat(esper_ctx* ctx, const char* name);

/* ---- stage 1: pairwise RMS image distances -------------------------------------------
 * Replaces the non-CTF branch of 1_Embedding/DM/Distances.py:87-104:
 *     x_i <- (x_i - mean)/std per image (float32, ddof 0)              :88-92
 *     D_ij = sqrt( sum_p (x_ip - x_jp)^2 / P ), D_ii = 0, symmetric      :96-104
 * computed as a 3xTF32 split-precision Gram contraction on tcgen05 tensor cores with short FP32
 * accumulation chains combined in FP64, ||x||^2+||y||^2-2x.y epilogue, clamp at 0, sqrt.
 * flags: */
#define ESPER_DIST_STANDARDIZE  1u  /* apply the per-image z-score of Distances.py:88-92        */
#define ESPER_DIST_CENTER       2u  /* subtract the PD mean image before the Gram (distance-
                                       invariant; removes the common-mode term)   */
#define ESPER_DIST_REFINE       4u  /* recompute pairs with D^2*P < refine_tau*(|x|^2+|y|^2) by
                                       direct differences in FP64 (near-duplicate images)     */
#define ESPER_DIST_DEFAULT      (ESPER_DIST_STANDARDIZE | ESPER_DIST_CENTER | ESPER_DIST_REFINE)

/* imgs: [n, P] float32 (an MRC stack [n, N, N] flattened, P = N*N), or NULL = resident stack
 * uploaded by esper_upload_images.  D: [n, n] float64 out, or NULL = keep on device.         */
int esper_dist_rms(esper_ctx* ctx, const float* imgs, int n, int P, unsigned flags, double* D);
/* Asynchronous H2D of a stack into the ctx (pinned host memory recommended); becomes resident. */
int esper_upload_images(esper_ctx* ctx, const float* imgs, int n, int P);
/* Tuning knobs of stage 1 (defaults in parentheses): split_k = number of K chunks whose FP32
 * partial Grams are combined in FP64 (0 = auto: chunks of ~2048 pixels, rounded so the work
 * units fill the SMs); refine_tau (0.05: pairs with D^2*P < 0.05(|x|^2+|y|^2), i.e. near-duplicate images, where
 * Gram cancellation would exceed 1e-5; calibrated with tools/calib_gram_error.py). */
int esper_dist_config(esper_ctx* ctx, int split_k, double refine_tau);
/* With ESPER_DIST_REFINE, stacks of at most this many images are computed pair by pair in the reference's own arithmetic
 * (float64 sum of squared differences of the float32 z-scores, Distances.py:96-104): the reference's CPU-runnable inputs (400
 * noise-free states) sit entirely in the Gram-cancellation regime, and the exp(-D^2/2eps) of stage 3 amplifies a 1e-5
 * distance error ~20x.  Larger stacks use the tensor-core contraction with refinement of the near-duplicate pairs only. */
#define ESPER_EXACT_PAIRS_MAX_N 512
#define ESPER_DEFAULT_EIG_TOL 1e-8
/* Operand format of the stage-1 tensor-core contraction.  mode 0: 3xTF32 -- every centred z-score c is split into
 * two TF32 values, hi*hi + hi*lo + lo*hi accumulate in FP32.  mode 1 (default): 3xFP16 -- the same split into two FP16 values of
 * c * 2^s (an FP16 significand has the same 11 bits as a TF32 one; s is chosen from P so that |c| <= 2 sqrt(P) stays in
 * range), tcgen05 kind::f16 at twice the TF32 rate with half the operand bytes; used for esper_dist_rms (stacks of more
 * than 128 images) and esper_dist_rms_rows with ESPER_DIST_STANDARDIZE, the 3xTF32 path otherwise (unstandardised input,
 * PCA, CTF branch).  ESPER_GRAM_F16=0/1 sets the initial mode of new contexts.
 * esper_gram_info: operand format the last esper_dist_rms used (0 TF32, 1 FP16). */
int esper_gram_config(esper_ctx* ctx, int mode);
int esper_gram_info(esper_ctx* ctx, int* last_operands);
/* Upload a distance matrix so that stages 2/3 can use it as the resident D. */
int esper_upload_dist(esper_ctx* ctx, const double* D, int m);

/* ---- stage 2: epsilon sweep -------------------------------------------------------------
 * Replaces the loop of 1_Embedding/DM/GaussianBandwidth.py:36-44:
 *     logSum[k] = log sum_{ij : t < thr} exp(-t),  t = coef[k] * D_ij^2
 * coef[k] = 1/(2*exp(logEps[k])) is computed by the caller exactly as the reference does
 * (`1./(2.*eps)`), so the inclusion test `t < thr` is bit-identical.  D: [m, m] or NULL=resident.
 * thr: GaussianBandwidth.find_threshold (:26-32); pass +inf for "no truncation".          */
int esper_eps_sweep(esper_ctx* ctx, const double* D, int m, const double* coef, int n_eps,
                    double thr, double* logSum);
/* ---- stage 2b: the tanh fit (host code, no ctx, no device) ---------------------------------
 * Replaces the body of the retry loop of 1_Embedding/DM/GaussianBandwidth.py:47-57 for one start vector:
 *     popt, pcov = curve_fit(fun, logEps, logSumWij, p0=a0);  resnorm = sum(sqrt(abs(diag(pcov))));  R_squared
 * with fun(x) = a3 + a2 tanh(a0 x + a1) (:18-24).  curve_fit is MINPACK's lmdif behind scipy.optimize.leastsq; this is a
 * restatement of that algorithm with scipy's default tolerances (k9_fit.cu), so that a whole PD can run without the
 * Python interpreter; popt agrees with scipy's to ~1e-12 relative on the golden sweeps (tests/test_host.py).
 * x, y: [n] finite values (ESPER_E_ARG otherwise -- curve_fit raises ValueError there); a0, popt: [4];
 * info: MINPACK termination code (1..4 = converged); resnorm = +inf when the covariance cannot be estimated (curve_fit
 * returns an infinite pcov there), which makes the caller's `while resnorm > 100` loop draw a new start as in the reference. */
int esper_tanh_fit(const double* x, const double* y, int n, const double* a0, double* popt, double* resnorm,
                   double* r_squared, int* info, int* nfev);

/* Sweep algorithm: mode 0 evaluates exp() for every (element, k) between the plateaus; mode 1 (default) is the
 * moment sweep -- D^2 grouped into geometric bins, 17 moments per bin give every fully included k at once, only the
 * k whose truncation boundary falls inside a bin are evaluated per element (the inclusion test stays bit-identical) --
 * with an automatic return to mode 0's kernel when D^2 spans more than 32 bins or the coefficients are not a
 * non-increasing grid.  Results agree to ~1e-14 in logSum.  The environment variable
 * ESPER_SWEEP_MOMENTS=0/1 sets the initial mode of new contexts.
 * esper_sweep_info: which kernel produced the last esper_eps_sweep result (0 general, 1 moments) and its bin count. */
int esper_sweep_config(esper_ctx* ctx, int mode);
int esper_sweep_info(esper_ctx* ctx, int* last_path, int* bins);

/* ---- stage 3: kernel, degree normalisation, leading eigenpairs ---------------------------
 * Replaces 1_Embedding/DM/DiffusionMaps.py:109 (A = exp(-D^2/(2 eps))), :131-152
 * (Ms = D^-1/2 A D^-1/2, i.e. alpha = 0) and :166-169 (svd; saves s^2 and U).
 * esper_markov exposes the intermediate for tests: degree [m] and Ms [m, m] (either may be NULL). */
int esper_markov(esper_ctx* ctx, const double* D, int m, double eps, double* degree, double* Ms);
/* Leading k eigenpairs of Ms by deflated Lanczos with full reorthogonalisation, 1 <= k <= m.  k <= 64: the iteration stops when
 * the k pairs have converged.  k > 64, up to k = m (the reference's full U): Lanczos is run to exhaustion (m - 1 steps), the host
 * diagonalises the tridiagonal matrix by implicit QL recording its plane rotations, the device applies them to the rows of the
 * basis (csrc/k4_full.cu; 0.4 s at m = 2000); ESPER_E_NOCONV when fewer than k pairs exist (duplicated images: rank-deficient Ms).
 *   val: [k]   = lambda^2, descending (what the reference saves as PD_%s_val.npy)
 *   vec: [m,k] row-major; column 0 = sqrt(degree)/|sqrt(degree)| (steady state), unit columns,
 *              sign fixed so that the entry of largest magnitude is positive
 *   tol: relative residual |Ms v - lambda v| <= tol*lambda_1  (0 = ESPER_DEFAULT_EIG_TOL = 1e-8).  The default is tied to the
 *        parity contract (eigenvalues 1e-6 relative, eigenvectors 1e-4): a Ritz value is off by residual^2 / gap (~1e-14 here)
 *        and a Ritz vector by residual / gap, i.e. <= 5e-6 for every pair whose gap to its neighbours is at least 2e-3 lambda_1
 *        (the pairs the contract compares one by one); members of tighter clusters -- the bulk edge of a noisy PD, where the
 *        tenth and eleventh eigenvalue differ by 1e-4 relative -- are determined as a subspace, as the SVD of the reference
 *        itself determines them.  On the benchmark matrix 1e-8 needs 164 Lanczos steps, 1e-10 needs 184 (tools/eig_probe.py).
 *   max_iter: Lanczos steps (0 = min(m-1, 1024));  iters (optional) receives the steps used.  */
/* Eigensolver scheduling for m <= 2048.  mode 0 (default; lowest latency AND, since the merged schedule, highest throughput):
 * ONE persistent kernel for the whole solve -- 15 thread-block clusters of 8 CTAs keep the rows of Ms and of the Lanczos basis
 * in shared memory / registers, one exchange per step (distributed shared memory inside a cluster, flag-carrying packets
 * between clusters), convergence tested on the device by checker CTAs on a fixed schedule of sizes (the step count of the
 * result is reproducible); the host is not involved between the launch and the final copy (csrc/k4_chain.cu).  It covers
 * m <= 2048 and up to ~300 steps; anything else, and a solve that does not converge within that, runs on the barrier kernels.
 * mode 1: rows streamed from L2, batches of 16 steps, convergence tests on the host while the next batch runs; the kernel fits
 * twice per SM.  mode 2: mode 1 without the speculatively launched batch.  mode 3: rows in shared memory, three grid barriers
 * per step (round-1 default).  Same arithmetic, same results to rounding (the modes stop at different steps).
 * Environment (experiments only): ESPER_CHAIN_MERGED=0 two exchanges per step; ESPER_CLUSTER_SIZE=0/2/4/8/16; ESPER_CHAIN_NOCOOP=1
 * launches the cluster kernel without the cooperative attribute (Nsight Compute cannot launch it otherwise). */
int esper_dmap_config(esper_ctx* ctx, int mode);
int esper_dmap_embed(esper_ctx* ctx, const double* D, int m, double eps, int k, double tol,
                     int max_iter, double* val, double* vec, int* iters);

/* ---- stages 1-3 in one call (SURVEY 8b `esper_pd_embed_fused`; the throughput harness and launcher.py use it) --------
 * One PD from the image stack to its leading eigenpairs without returning to the caller in between:
 *   esper_dist_rms (Distances.py:87-104)  ->  threshold 10 (GaussianBandwidth.py:26-32: the diagonal of this D is exactly 0,
 *   so find_threshold's first term is <= 4.6)  ->  esper_eps_sweep (:36-44)  ->  esper_tanh_fit from the start a0 (:47-57, one
 *   curve_fit)  ->  eps = exp(-a1/a0) (DiffusionMaps.py:91-93)  ->  esper_dmap_embed (:109-169).
 * imgs [n, P] float32 or NULL = the resident stack.  logEps / coef [n_eps]: the grid and the caller's 1./(2.*exp(logEps))
 * (same convention as esper_eps_sweep).  a0 [4].  Outputs are caller-allocated; D [n, n] may be NULL (when given, its
 * device-to-host copy runs on a second stream behind stages 2-3); logSum [n_eps]; fit [8] = popt[0..3], resnorm, R_squared,
 * eps, slope = popt[0]*popt[2]; val [k]; vec [n, k]; iters optional.
 * Returns ESPER_E_NOFIT when the fit from a0 fails the reference's `resnorm > 100` retry test: D stays resident, logSum and
 * fit[0..5] are filled, and the caller continues with a new start (esper_tanh_fit) and esper_dmap_embed(D = NULL). */
int esper_pd_embed_fused(esper_ctx* ctx, const float* imgs, int n, int P, unsigned flags, const double* logEps,
                         const double* coef, int n_eps, const double* a0, int k, double tol, int max_iter,
                         double* D, double* logSum, double* fit, double* val, double* vec, int* iters);

/* float32 mean and population standard deviation of each image, bit-identical to numpy's `img.mean()` / `img.std()`
 * on a float32 array (pairwise summation tree of numpy/_core/src/umath/loops_utils.h.src, two-pass variance of
 * numpy/_core/_methods.py): the statistics the CTF branch standardises with (Distances.py:114-115).
 * imgs [n, P] float32 (NULL: resident stack); mean, stdv: [n] float32. */
int esper_image_stats(esper_ctx* ctx, const float* imgs, int n, int P, float* mean, float* stdv);

/* ---- stage 1, CTF branch (1_Embedding/DM/Distances.py:107-145, the reference's shipped default) ----
 * imgs [n, box, box] float32; params [n, 5] float64 = pixel size [A], spherical aberration [mm], defocus [A],
 * voltage [keV], amplitude contrast per image -- the star-file columns read at Distances.py:68-72
 * (rlnPixelSize, rlnSphericalAberration, rlnDefocusU, rlnVoltage, rlnAmplitudeContrast).  Per image: float32
 * z-score (Distances.py:114-115), order-8 Butterworth low-pass at 0.5 Nyquist in Fourier space (:118-122,
 * Generate_Filters.py:34-39,78-80), CTF on the FFT grid with B = inf (Generate_Filters.py:41-58,70-75).
 *   D [n, n]            defocus-tolerant distances sqrt(sum_k |CTF_j fy_i - CTF_i fy_j|^2) with the reference's
 *                       unnormalised fft2 (Distances.py:139-145); diagonal exactly 0, symmetric; D == NULL keeps the
 *                       matrix resident only (chain with esper_eps_sweep / esper_dmap_embed(D = NULL))
 *   wiener [n, box, box] float32 CTF-corrected stack ifft2(fft2(img) CTF_i / -(sum_i CTF_i^2 + 1/5)).real
 *                       (Distances.py:128-135,155-161; the file `*_filtered.mrcs` read by Manifold_Binning.py:98);
 *                       NULL skips it
 * flags: ESPER_DIST_REFINE recomputes pairs whose D^2 is below refine_tau x the summed Gram terms pair by pair in
 * FP64.  The reference's `D -= np.diag(D)` (:144, a row-broadcast of a diagonal that is zero up to rounding) is not
 * reproduced; negative round-off under the square root is clamped to 0 where the reference yields NaN. */
int esper_ctf_dist(esper_ctx* ctx, const float* imgs, int n, int box, const double* params, unsigned flags,
                   double* D, float* wiener);

/* ---- PCA embedding (1_Embedding/PCA/PCA.py:49-71) ---------------------------------
 * Y [P, n] = the raw images as columns (PCA.py:49-51, no per-image standardisation), minus the mean image
 * over the stack (PCA.py:53-55); the reference takes the SVD Y = u s v^T (PCA.py:58) and saves s^2 and the
 * scores Y^T u[:, :dim] = v[:, :dim] s (PCA.py:66-68).  Here the leading k eigenpairs (s^2, v) of the n x n
 * Gram matrix Y^T Y are computed with the stage-1 tensor-core Gram kernel and the stage-3 Lanczos solver
 * (the constant vector, eigenvalue 0 after centring, is locked out), so 1 <= k <= n-1.
 *   val [k]    = s^2 descending (the reference saves all n)
 *   U   [n, k] = scores, row-major; column sign: the entry of largest |.| is positive (the reference's sign is
 *                whatever LAPACK returns)
 * imgs == NULL: use the resident stack. */
int esper_pca_embed(esper_ctx* ctx, const float* imgs, int n, int P, int k, double tol, int max_iter,
                    double* val, double* U, int* iters);

/* ---- oversized PD: distance-matrix row blocks (BASELINE config 4) ------------------
 * One PD too large for one GPU is split by row blocks of D / A / Ms: rank g owns rows [row0, row0+nrows)
 * (row0 a multiple of 128; nrows a multiple of 128 unless the block ends at row n).  The same reference arithmetic as above (Distances.py:87-104,
 * GaussianBandwidth.py:36-44, DiffusionMaps.py:109-169) is evaluated on the block; the collectives
 * (sum of the sweep partials, all-gather of the degrees and of the Lanczos vector) are done by the host
 * layer between these calls (manifoldem_esper_b200/rowblock.py, torch.distributed over NCCL).
 * Tiles below the diagonal are computed as the mirror image of the upper-triangle tile, so the block holds
 * bit-identical values to the corresponding rows of esper_dist_rms for the same split_k. */
int esper_dist_rms_rows(esper_ctx* ctx, const float* imgs /*[n,P] or NULL=resident*/, int n, int P, unsigned flags,
                        int row0, int nrows, double* D_rows /*[nrows,n] or NULL*/);
/* sums[k] = sum over the resident block of exp(-t) with t < thr (NOT the logarithm: ranks add these). */
int esper_eps_sweep_rows(esper_ctx* ctx, int nrows, int n, const double* coef, int n_eps, double thr, double* sums);
/* Moment sweep over row blocks (opt-in, not yet run on hardware; DESIGN.md 6f, manifoldem_esper_b200/moment_sweep.py).
 * stats4 = {min non-zero D^2, max D^2, number of exact zeros, number of valid elements} of the resident block; the ranks
 * all-reduce them (min, max, sum, sum) and build ONE plan (moment_sweep.plan: geometric bins, k ranges per bin).
 * moment_accum_rows: table [32][21] = per bin 17 moments + 4 boundary sums of the block against that plan; the ranks
 * all-reduce(sum) the tables and moment_sweep.finalize evaluates every epsilon.  esper_moment_layout: {MP, MS, MB, plan
 * header, plan entries per bin} the library was built with (host only). */
int esper_moment_layout(int* out5);
int esper_moment_stats_rows(esper_ctx* ctx, int nrows, int n, double* stats4);
int esper_moment_accum_rows(esper_ctx* ctx, int nrows, int n, const double* coef, int n_eps, double thr,
                            const double* plan, double* table);
/* degree of the owned rows (DiffusionMaps.py:131-137): degree_rows [nrows]. */
int esper_degree_rows(esper_ctx* ctx, int nrows, int n, double eps, double* degree_rows);
/* Ms rows from the gathered degrees of all n rows (DiffusionMaps.py:139-152); block stays resident. */
int esper_markov_rows(esper_ctx* ctx, int nrows, int n, double eps, const double* degree_all);
/* w_rows_d[nrows] = Ms_rows . q_d[n]   (device pointers; asynchronous on the ctx stream). */
int esper_symv_rows_d(esper_ctx* ctx, int nrows, int n, const double* q_d, double* w_rows_d);
/* Fused SYMV + all-gather over NVLink peer memory (one process per GPU, CUDA IPC; replaces symv_rows_d followed by an
 * NCCL all-gather).  comm_alloc: this rank's receive buffer for gathered vectors of n_pad = rows_per_rank * world
 * elements; ipc_handle: 64 bytes to exchange with the peers (e.g. torch.distributed all_gather).
 * comm_open: ipc_handles = [world][64] in rank order.  symv_allgather_d (step = 1, 2, ...): w = Ms_rows . q_d, every
 * element stored as a 16-byte {value, step} packet at [row0, row0 + nrows) of slot step & 1 of EVERY rank's buffer as
 * soon as its CTA has it (flag-in-data, no fences).  comm_wait_d: enqueues on the ctx stream the unpack of elements
 * [0, n) of `step` (spins per element until its flags show the step); *w_full_d = the dense gathered vector on this
 * rank (valid for kernels enqueued afterwards; pass it to lanczos_ortho_d).  Steps must be consecutive on all ranks,
 * each rank on a GPU of its own.  comm_status: 0, or non-zero if an unpack timed out (10 s watchdog, no hang). */
int esper_comm_alloc(esper_ctx* ctx, int n_pad, int world, void* ipc_handle);
int esper_comm_open(esper_ctx* ctx, const void* ipc_handles, int rank, int world);
int esper_comm_close(esper_ctx* ctx);
int esper_symv_allgather_d(esper_ctx* ctx, int nrows, int n, int row0, const double* q_d, int step);
int esper_comm_wait_d(esper_ctx* ctx, int step, int n, double** w_full_d);
int esper_comm_status(esper_ctx* ctx, unsigned int* status);
/* n_steps Lanczos steps j0 .. j0+n_steps-1 of the fused exchange enqueued by ONE host call: per step esper_symv_allgather_d (tag
 * step0 + i) -> esper_comm_wait_d -> esper_lanczos_ortho_d on the gathered vector; alpha_beta_d is the [max_steps, 2] array, row j-1
 * receives step j.  The host paces the iteration once per convergence test instead of once per step. */
int esper_lanczos_fused_steps_d(esper_ctx* ctx, int nrows, int n, int row0, double* V_d, double* alpha_beta_d, int j0, int n_steps, int step0);
/* Replicated Lanczos state, vectors owned by the caller on the device: V_d is [max_steps+2, n].
 * start: V_d[0] = sqrt(degree)/|.| (locked), V_d[1] = start vector.  ortho (step j >= 1, w_d = gathered
 * Ms V_d[j]): Gram-Schmidt twice against V_d[0..j], alpha_beta_d = {alpha_j, beta_j}, V_d[j+1] = w/beta_j. */
int esper_lanczos_start_d(esper_ctx* ctx, int n, const double* degree_all, double* V_d);
int esper_lanczos_ortho_d(esper_ctx* ctx, int n, int j, double* V_d, double* w_d, double* alpha_beta_d);
/* Host only: Ritz values/vectors of T = tridiag(beta, alpha, beta) (steps x steps), convergence of the leading
 * k-1 pairs to relative residual tol; theta [k-1] (may be NULL), S [steps, k-1] (may be NULL). */
/* Host only: ALL eigenpairs of T = tridiag(beta, alpha, beta) of order n (implicit QL with recorded plane rotations, the host
 * half of the full-spectrum path of esper_dmap_embed for k > 64): theta [n] descending, Z [n, n] row-major, eigenvectors in
 * the columns.  beta[n-1] is ignored. */
int esper_tridiag_eigen_all(const double* alpha, const double* beta, int n, double* theta, double* Z);
int esper_tridiag_ritz(const double* alpha, const double* beta, int steps, int k, double tol, int* converged,
                       double* theta, double* S);
/* vec [n, k] = [V_d[0], V_d[1..steps] S], unit columns, sign fixed (same convention as esper_dmap_embed). */
int esper_ritz_vectors_d(esper_ctx* ctx, int n, int steps, int k, const double* V_d, const double* S, double* vec);

/* ---- stage 4: Nd rotation + 2-D histogram sweep -----------------------------------------
 * esper_nd_rotation: host-side restatement of 2_Manifold_Rotations/Generate_Nd_Rot.py:4-45,
 * R = G_0 G_1 ... G_{T-1} (T = n(n-1)/2, lexicographic planes), left to right, each two-term sum
 * accumulated as a dgemm micro-kernel does (first product rounded, second fused).  R: [n, n].  Returns ESPER_E_ARG unless n_theta == T (ValueError :7-12). */
int esper_nd_rotation(int n, const double* theta, int n_theta, double* R);
/* numpy.linspace(lo, hi, bins+1) as used by numpy.histogramdd; edges: [bins+1]. */
int esper_hist_edges(int bins, double lo, double hi, double* edges);
/* Batch of independent evaluations of the sweep body, Rotation_Histograms.py:304-312:
 *   U_rot = (R_e @ Uinit[pd_e].T).T ; H = histogramdd(U_rot[:, (v1_e, v2_e)], bins, range lo..hi)
 * Uinit: [n_pd, m, dim]; R: [n_eval, dim, dim]; nonzero: [n_eval] = count_nonzero(H);
 * H (optional): [n_eval, bins, bins] int32 counts.                                           */
int esper_rot_hist_eval(esper_ctx* ctx, const double* Uinit, int n_pd, int m, int dim,
                        const double* R, const int* pd_of_eval, const int* v1, const int* v2,
                        int n_eval, int bins, double lo, double hi, int* nonzero, int* H);
/* The whole sequential search of Rotation_Histograms.py:283-365 for a batch of PDs on device:
 * for every PD and subspace (v1,v2) of its comboList, for each Rij in order, evaluate all
 * n_theta angles, keep the FIRST argmax of the empty-bin score, move on to the next Rij.
 *   combo:  [n_pd, max_sub, 2] int (v1, v2), n_sub: [n_pd];  rij: [n_pd, n_rij] int
 *   theta_grid: [n_theta] radians (np.arange(start,stop,step)*pi/180 computed by the caller)
 *   Uinit == NULL: the [n_pd, m, dim] manifolds left on the device by esper_manifold_prepare
 *   thetas_out: [n_pd, max_sub, T] radians (rows >= n_sub[p] are zero)
 *   counts_out (optional): [n_pd, max_sub, n_rij, n_theta] non-zero-bin counts               */
int esper_rot_sweep(esper_ctx* ctx, const double* Uinit, int n_pd, int m, int dim,
                    const int* combo, const int* n_sub, int max_sub, const int* rij, int n_rij,
                    const double* theta_grid, int n_theta, int bins, double lo, double hi,
                    double* thetas_out, int* counts_out);

/* ---- stage-4 preparation on the device ---------------------------------------
 * esper_manifold_prepare: Rotation_Histograms.py:100-112 -- `normalize` (every column of U0 to [0,1], then the whole
 * matrix to [-1,1]; bit-identical to numpy) and the slice U[:, first:first+dim] (first = 1 for diffusion maps, 0 for
 * PCA); then, when kind != 0, the conic fits of findParabolas (:153-212) for every column pair v1 < v2:
 * kind 2 = 3_Manifold_Binning/ConicFit.py:205-264 (`fit2`), kind 3 = ConicFit.py:267-283 (`fit3`).
 *   U0: [n_pd, m, ncol]; Uinit (optional out): [n_pd, m, dim] -- it also stays resident on the device, and a following
 *   esper_rot_sweep / esper_conic_fit_batch with Uinit == NULL uses that copy;
 *   R2: [n_pd, dim(dim-1)/2] coefficient of determination per pair in lexicographic order (0,1),(0,2),...;
 *   coef (optional): [n_pd, pairs, 5] fit coefficients (fit2: B, C, D, E, F of x^2+Bxy+Cy^2+Dx+Ey+F; fit3: a, b, c, 0, 0). */
int esper_manifold_prepare(esper_ctx* ctx, const double* U0, int n_pd, int m, int ncol, int first, int dim, int kind,
                           double* Uinit, double* R2, double* coef);
/* The conic fits alone on caller-supplied (already normalised / pre-rotated) manifolds Uinit [n_pd, m, dim]. */
int esper_conic_fit_batch(esper_ctx* ctx, const double* Uinit, int n_pd, int m, int dim, int kind, double* R2, double* coef);

/* ---- data preparation and formats either side of the path ---------------------
 * esper_tau_snr_stack: 0_Data_Inputs/Pristine_AddTauSNR/Generate_TauSNR.py:56-64 with AddNoise.py:17-86 per image:
 * every pristine state -> tau copies with additive Gaussian noise N(signal mean, sqrt(var(signal)/snr)) (signal =
 * pixels outside +-0.25 std), each standardised by the mean/std of its pixels outside the inscribed circle.
 *   pristine: [ss, box, box] float32 (host); image s*tau + t of the result is copy t of state s;
 *   snr == 0: tau exact copies without noise or standardisation (apply_noise = False, :62-63);
 *   seed: Philox4x32-10 key (the reference uses numpy's unseeded global generator);
 *   params (optional out): [ss, 2] = (signal mean, noise std) per state -- AddNoise.find_SNR;
 *   stack (optional out): [ss*tau, box, box] float32.  The stack also stays resident: a following
 *   esper_dist_rms(imgs = NULL, n = ss*tau, P = box*box) embeds it without any host round trip.                */
int esper_tau_snr_stack(esper_ctx* ctx, const float* pristine, int ss, int box, int tau, double snr,
                        unsigned long long seed, double* params, float* stack);
/* D2H copy of the resident image stack [n, P] (e.g. to write a generated stack as .mrcs). */
int esper_download_images(esper_ctx* ctx, float* imgs, int n, int P);
/* MRC2014 header of `path` (host only): dims = (nx, ny, nz), mode, byte offset of the first image
 * (1024 + extended header).  Replaces what mrcfile.mmap reports (Distances.py:75-76).  Error text: esper_last_error(NULL). */
int esper_mrc_info(const char* path, int* dims, int* mode, long long* data_offset);
/* Stream images [first, first+count) (count <= 0: to the end) of a mode-2 MRC stack from disk into the resident image
 * buffer through two page-locked staging buffers (disk read of chunk i+1 overlaps the H2D copy of chunk i).
 * dims (optional out) = (images, ny, nx).  Follow with esper_dist_rms(imgs = NULL, ...).                            */
int esper_upload_mrc(esper_ctx* ctx, const char* path, int first, int count, int* dims);
/* 3_Manifold_Binning/Manifold_Binning.py:940-974: sums[b] = sum of the float32 images order[offsets[b] .. offsets[b+1])
 * in float64, added in that order (bit-identical to the reference's `finalStackBin[b] += image` loop).
 *   imgs: [n, P] float32 or NULL = resident stack; order: image indices; offsets: [n_bins+1]; sums: [n_bins, P].   */
int esper_bin_accumulate(esper_ctx* ctx, const float* imgs, int n, int P, const int* order, const int* offsets,
                         int n_bins, double* sums);

#ifdef __cplusplus
}
#endif
#endif /* ESPER_B200_H */
