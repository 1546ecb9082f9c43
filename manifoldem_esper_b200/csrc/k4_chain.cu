// k4_chain.cu -- K4, default schedule for m <= 2048: Lanczos with full reorthogonalisation as ONE persistent kernel
// WITHOUT grid barriers.  Replaces the full SVD of 1_Embedding/DM/DiffusionMaps.py:166-169 (leading pairs only).
//
// Three generations of worker kernels share the checker CTAs, the packet protocol and the host side of this file:
//   lanczos_merged_kernel   (default for the Markov matrix)  clusters of 8 CTAs, ONE exchange per step: the unorthogonalised
//                            candidate is gathered together with the reduction of its coefficients, Ms q by linearity
//                            (comment in front of the kernel)
//   lanczos_cluster_kernel  clusters of 8 CTAs, two exchanges per step (coefficients, then the new vector); distributed shared
//                            memory inside a cluster, packets between clusters; used when the locked pair's eigenvalue is unknown
//   lanczos_chain_kernel    no clusters: every hand-over is a packet through L2 (described below); used when clusters of 8
//                            cannot be made resident
//
// The three-barrier kernel of k4_eigen.cu spent 40 % of every step inside grid barriers and the rest waiting for one L2
// round trip per phase (profiles/r1_bench_lines.md).  Here every hand-over between CTAs is a flag-carrying 16-byte packet
// {lo, tag, hi, tag} (the protocol NCCL calls LL; already used by the row-block all-gather): the consumer polls the data
// itself, so a hand-over costs one store and one load instead of arrive + poll + load.
//
//   worker CTA b (owns rows R_b of Ms on chip: shared memory + registers, and the same rows of every basis vector in shared
//   memory, Vs[r][i]); the whole current Lanczos vector q_j is in its shared memory:
//     S0   w_r = Ms[r,:] q_j - sigma q_j[r] - beta_{j-1} q_{j-1}[r]          r in R_b          (on-chip operands only)
//     P1   partial coefficients  p_b[s] = sum_{r in R_b} Vs[r][s-1] w_r  (s >= 1),  p_b[0] = sum_r w_r^2   -> packets pbuf[s][b]
//     P2   the CTA that owns slot s (s mod workers) adds the partials of all workers in a fixed order     -> packet  hbuf[s]
//     P3   every CTA reads all h; w_r -= sum_i Vs[r][i] h_i;  beta^2 = |w|^2 - sum h^2 (exact for an orthonormal basis; the
//          DGKS test beta^2 < |w|^2 / 2 asks for a second P1-P3 pass, in which |w'|^2 is summed again);
//          q_{j+1}[r] = w_r / beta  -> Vs, global V (for the Ritz vectors) and packets qbuf[r]
//     P4   every CTA reads q_{j+1} (m packets) into shared memory
//   Three dependent hand-overs per step, no barrier, every reduction in a fixed order (deterministic, identical in all CTAs).
//
//   checker CTAs (the CTAs that own no rows): follow alpha_j / beta_j (packets abuf[j] from CTA 0) and test convergence of
//   the leading Ritz pairs on the device, at a FIXED schedule of sizes (the oldest passing test wins: the step count of the
//   result does not depend on timing): multisection Sturm counts with all threads (division-free scaled recurrence); a gate
//   ("have the values grown over the last 8 steps": one Sturm count per value on the leading block T_{n-8}); the last components
//   of the Ritz vectors by the three-term recurrence run backwards (one thread per vector) for the residuals |beta_n s_{n,i}|.
//   The winner raises the stop word; CTA 0 forwards it to all workers inside the coefficient exchange of the next step, so
//   that all of them leave at the same step; it then forms the Ritz vectors (the same recurrence, stored, its residual measured;
//   inverse iteration with a pivoted LU as the fallback) and publishes theta, S and n.  The host is not involved until the final copy.
//
// A failure of any spin wait (~2 s) raises the abort word and ends the kernel; the host then falls back to the
// barrier kernels of k4_eigen.cu, as it does for sizes this schedule does not cover.
#include "common.cuh"
#include "devutil.cuh"
#include <algorithm>
#include <vector>

namespace esper {
namespace chain {

constexpr int CH_THREADS = 512;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr int CH_MAXROWS = 14;          // rows of Ms per worker CTA
constexpr int CH_REGROWS = 4;           // of which in registers (the rest in shared memory)
constexpr int CH_MAXM = 2048;
constexpr int CH_MAXWANT = 24;          // Ritz pairs followed by a checker (kw + guard)
constexpr int CH_MAXCHK = 4;
constexpr int CH_POINTS = 16;           // multisection points per eigenvalue and round
constexpr int CH_HWARP = 8;             // warps 0..7 may own coefficient slots (P2); warps 8..11 fetch the coefficients (P3)
constexpr int CH_PK = 8;                // partial packets per lane and slot: workers <= 32 * CH_PK
constexpr long long CH_TIMEOUT = 4000000000LL;     // ~2 s
constexpr int CH_MAXTESTS = 192;        // convergence tests of one solve (sizes first_check + t check_stride)
// cluster schedule (lanczos_cluster_kernel): thread-block clusters with distributed shared memory as the first exchange level
constexpr int CL_MAXROWS = 17;          // rows of Ms per worker CTA (a B200 keeps 15 clusters of 8 resident: 118 workers at m = 2000)
constexpr int CL_REGROWS = 7;           // of which in registers
constexpr int CL_GPK = 4;               // packets per thread in flight in the gather
constexpr int CL_SMROWS = CL_MAXROWS - CL_REGROWS;   // rows in shared memory
constexpr int CL_LD = 2048;             // row stride in shared memory: fixed, so that the products need no address arithmetic
constexpr int CL_MAXCLUSTERS = 32;      // clusters that take part in an exchange (one lane each)

struct Ctl {
    unsigned int stop;                  // 0, or 0x80000000 | checker << 16 | n  (first passing check)
    unsigned int abort;                 // a spin wait timed out
    unsigned int final_n;               // workers ended without a stop: 0x80000000 | breakdown << 30 | steps
    unsigned int progress;              // steps completed (alpha/beta published)
    int status;                         // 1 converged, 2 not converged at max_steps
    int n_final, winner, steps_run;
    long long prof[32];
    unsigned int verdict[CH_MAXTESTS];  // per scheduled test t: 0 not decided, 1 failed, 2 passed (the OLDEST passing test stops the iteration)
};

struct Params {
    const double* Ms; int m;
    double* V;                          // [max_steps + 2][m]: V[0] locked vector, V[1] start vector (both set by the host side)
    double* alpha; double* beta;        // plain copies for the host (alpha[j-1], beta[j-1] of step j)
    int chunk, n_workers, n_checkers;
    int max_steps;                      // <= vs_cap - 2
    int vs_cap;                         // basis vectors per owned row held in shared memory
    int kw, guard, cap_steps;           // pairs wanted (without the locked one), guard pairs, Krylov dimension m - 1
    double tol;
    int check_back;                     // the gate of a test compares T_n with T_{n - check_back}
    int first_check, check_stride;      // convergence tests at the sizes first_check + t check_stride, t = 0, 1, ...; checker c takes t = c (mod n_checkers)
    uint4* pbuf;                        // [slots][n_workers]
    uint4* hbuf;                        // [slots + 1]; the last packet is the control word
    uint4* qbuf;                        // [m]
    uint4* abuf;                        // [max_steps + 1][2]
    Ctl* ctl;
    double* theta_out;                  // [n_checkers][CH_MAXWANT]
    double* S_out;                      // [n_checkers][max_steps * kw]
    int s_stride;                       // doubles per checker in S_out
    int prof_on;
    // cluster schedule only
    int cl_size;                        // CTAs per cluster
    int cl_count;                       // clusters that hold workers
    int cl_group;                       // lanes per slot in the gather: smallest power of two >= cl_count
    uint4* cbuf;                        // [slots + 1][cl_count] per-cluster partial coefficients
    int merged;                         // 1: lanczos_merged_kernel (one exchange per step): cbuf has 4 buffers of cb_stride packets, qbuf 2 of m
    size_t cb_stride;
    int mg_order;                       // merged schedule: 1 forward, product, gather (default); 2 forward and gather side by side, then the product
    double lambda0;                     // merged schedule: eigenvalue of the locked vector V[0] (1 for the Markov matrix)
    long long* trace;                   // diagnostic (ESPER_DEBUG_LANCZOS): [CTAs][16] globaltimer stamps of thread 0 in step trace_step
    int trace_step;
};

// ---- packet and flag primitives (the CPU stand-in of tests/emu replaces exactly these) ----------------------------------
#ifndef ESPER_CHAIN_EMU
__constant__ unsigned int g_relax_ns = 0;      // experiment knob (ESPER_CLUSTER_RELAX_NS)
__device__ __forceinline__ void st_ll(uint4* p, double v, unsigned int tag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)__double2loint(v)), "r"(tag),
                 "r"((unsigned int)__double2hiint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ld_ll(const uint4* p, unsigned int tag, double& v) {
    uint4 x;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p) : "memory");
    v = __hiloint2double((int)x.z, (int)x.x);
    return x.y == tag && x.w == tag;
}
// the same in two halves: the load may be issued long before the packet is looked at
__device__ __forceinline__ uint4 ld_ll_raw(const uint4* p) {
    uint4 x;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p) : "memory");
    return x;
}
__device__ __forceinline__ bool ll_decode(const uint4& x, unsigned int tag, double& v) {
    v = __hiloint2double((int)x.z, (int)x.x);
    return x.y == tag && x.w == tag;
}
__device__ __forceinline__ unsigned int ld_word(const unsigned int* p) { return *((const volatile unsigned int*)p); }
__device__ __forceinline__ void st_word(unsigned int* p, unsigned int v) { *((volatile unsigned int*)p) = v; }
__device__ __forceinline__ unsigned int cas_word(unsigned int* p, unsigned int cmp, unsigned int v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ void publish_fence() { __threadfence(); }
__device__ __forceinline__ long long now() { return clock64(); }
__device__ __forceinline__ void relax() { __nanosleep(40); }
__device__ __forceinline__ void relax_short() { if (g_relax_ns) __nanosleep(g_relax_ns); }   // cluster kernel: the reload itself paces the poll
// Distributed shared memory.  dsm_send stores v at the copy of `dst_local` in the shared memory of CTA `rank` of this cluster
// and completes 8 transaction bytes on that CTA's copy of the mbarrier `bar_local` (st.async): the consumer waits on its own
// mbarrier, the producer neither fences nor arrives.  An mbarrier is armed once per phase by ONE local thread (mbar_expect =
// arrive + expected bytes); bytes that land before the arming only drive the transaction count negative.
typedef unsigned long long Mbar;
__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int cluster_rank() { unsigned int r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ void mbar_init(Mbar* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect(Mbar* b, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(Mbar* b, unsigned int parity) {
    unsigned int ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_addr(b)), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void dsm_send(double* dst_local, int rank, double v, Mbar* bar_local) {
    unsigned int ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_addr(dst_local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_addr(bar_local)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(ra), "l"(__double_as_longlong(v)), "r"(rb) : "memory");
}
// count-based mbarrier (no transaction bytes): `count` arrivals complete a phase; dsm_arrive arrives at the copy of `bar_local`
// in CTA `rank` of this cluster (release at cluster scope: what the arriving thread read before is done)
__device__ __forceinline__ void mbar_init_count(Mbar* b, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void dsm_arrive(Mbar* bar_local, int rank) {
    unsigned int rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_addr(bar_local)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ long long wall_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define CH_DEBUG_TEST(c, n, mode, track, stagnant)
#define CH_RR_BIG 1e100                 // the backward recurrences of the checkers rescale themselves beyond this magnitude
#endif

// ---- begin chain kernel text (tests/test_emu.py cuts from here) ----
// Spin until the packet carries `tag`; false after the watchdog or when another CTA raised the abort word.
__device__ __forceinline__ bool wait_ll(const uint4* p, unsigned int tag, double& v, Ctl* ctl) {
    if (ld_ll(p, tag, v)) return true;
    const long long t0 = now();
    for (unsigned int polls = 1;; ++polls) {
        relax();
        if (ld_ll(p, tag, v)) return true;
        if ((polls & 63u) == 0u) {
            if (ld_word(&ctl->abort) != 0u) return false;
            if (now() - t0 > CH_TIMEOUT) { st_word(&ctl->abort, 1u); return false; }
        }
    }
}

__device__ __forceinline__ bool wait_ll_short(const uint4* p, unsigned int tag, double& v, Ctl* ctl) {
    if (ld_ll(p, tag, v)) return true;
    const long long t0 = now();
    for (unsigned int polls = 1;; ++polls) {
        relax_short();
        if (ld_ll(p, tag, v)) return true;
        if ((polls & 63u) == 0u) {
            if (ld_word(&ctl->abort) != 0u) return false;
            if (now() - t0 > CH_TIMEOUT) { st_word(&ctl->abort, 1u); return false; }
        }
    }
}

// 16 per-lane values -> 16 warp totals with 16 shuffles; lane l ends with the total of value index
// ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1); fixed order.
__device__ __forceinline__ double multi_sum16(double (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 16;
        const double send = up ? v[i] : v[i + 8];
        const double keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 8;
        const double send = up ? v[i] : v[i + 4];
        const double keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 4;
        const double send = up ? v[i] : v[i + 2];
        const double keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool up = lane & 2;
        const double send = up ? v[0] : v[1];
        const double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Checker: convergence test of the leading Ritz pairs of T_n = tridiag(b, a, b) by one CTA.
// ---------------------------------------------------------------------------------------------------------------------
// Number of eigenvalues of T_n below x: sign changes of the Sturm sequence p_i = (a_i - x) p_{i-1} - b_{i-1}^2 p_{i-2},
// evaluated without divisions and rescaled by exact powers of two (a zero takes the sign opposite to its predecessor).
__device__ __forceinline__ int sturm_below(const double* a, const double* b2, int n, double x) {
    double pm = 1.0, pc = a[0] - x;
    int cnt = 0;
    bool neg = (pc < 0.0) || (pc == 0.0);
    if (neg) ++cnt;                                             // sign(p_1) != sign(p_0 = 1)
    for (int i = 1; i < n; ++i) {
        const double pn = (a[i] - x) * pc - b2[i - 1] * pm;
        const bool nneg = (pn < 0.0) || (pn == 0.0 && !neg);
        if (nneg != neg) ++cnt;
        neg = nneg;
        pm = pc; pc = pn;
        if ((i & 7) == 7) {                                     // keep |p| near 1: both by the same power of two
            const double mag = fmax(fabs(pc), fabs(pm));
            if (mag > 0.0 && (mag < 1e-100 || mag > 1e100)) {
                int e;
                frexp(mag, &e);
                pc = ldexp(pc, -e); pm = ldexp(pm, -e);
            }
        }
    }
    return cnt;
}

struct CheckSmem {
    double* a; double* b; double* b2;       // [cap] each
    double* theta;                          // [CH_MAXWANT]
    double* lo; double* hi;                 // [CH_MAXWANT]
    double* y; double* di; double* du; double* ff;   // [want][cap]: vector, 1/pivot, super-diagonal, multiplier (sign bit of swp = row swap)
    unsigned char* swp;                     // [want][cap]
    int* cnt;                               // [CH_THREADS]
    double* xs;                             // [CH_THREADS] section points of the current round
    double* red;                            // [3][CH_WARPS]
    double* prev;                           // [CH_MAXWANT] eigenvalues of this checker's previous test
    double* dots;                           // [CH_MAXWANT] orthogonalisation coefficients
};

// Shared-memory doubles a checker needs for `want` Ritz pairs and a Krylov dimension of at most cap (host and device).
__host__ __device__ inline size_t checker_doubles(int want, int cap) {
    return 3 * (size_t)cap + 5 * CH_MAXWANT + 3 * CH_WARPS + 4 * (size_t)want * cap + CH_THREADS + CH_THREADS / 2 + ((size_t)want * cap + 7) / 8 + 8;
}

// Factor T - x I = P L U (partial pivoting, as LAPACK dgttrf) once, keep it, and run `iters` inverse-iteration solves.
// One thread per Ritz vector.  du2[i] (second super-diagonal of U) is b[i+1] where rows were swapped, else 0.
__device__ void inverse_iteration_thread(const CheckSmem& cs, int n, int cap, int col, double x, double tiny) {
    double* di = cs.di + (size_t)col * cap;
    double* du = cs.du + (size_t)col * cap; double* ff = cs.ff + (size_t)col * cap;
    unsigned char* swp = cs.swp + (size_t)col * cap;
    const double* a = cs.a; const double* b = cs.b;
    // factorisation
    double d = a[0] - x;                    // current pivot candidate d[i]
    double dui = n > 1 ? b[0] : 0.0;        // du[i] as modified by the previous step
    for (int i = 0; i < n - 1; ++i) {
        const double dl = b[i];
        const double dn = a[i + 1] - x;     // unmodified d[i+1]
        const double dun = (i < n - 2) ? b[i + 1] : 0.0;   // unmodified du[i+1]
        if (fabs(d) >= fabs(dl)) {
            if (d == 0.0) d = tiny;
            const double r = 1.0 / d;       // the only division of the step (the multiplier is a product with it)
            const double f = dl * r;
            ff[i] = f; swp[i] = 0;
            di[i] = r; du[i] = dui;
            d = dn - f * dui;
            dui = dun;
        } else {                            // swap rows i and i+1
            const double r = 1.0 / dl;
            const double f = d * r;
            ff[i] = f; swp[i] = 1;
            di[i] = r; du[i] = dn;          // U[i][i] = dl, U[i][i+1] = old d[i+1], U[i][i+2] = old du[i+1]
            d = dui - f * dn;
            dui = -f * dun;
        }
    }
    if (d == 0.0) d = tiny;
    di[n - 1] = 1.0 / d; du[n - 1] = 0.0;
}

__device__ void inverse_solve_thread(const CheckSmem& cs, int n, int cap, int col) {
    double* y = cs.y + (size_t)col * cap; const double* di = cs.di + (size_t)col * cap;
    const double* du = cs.du + (size_t)col * cap; const double* ff = cs.ff + (size_t)col * cap;
    const unsigned char* swp = cs.swp + (size_t)col * cap;
    const double* b = cs.b;
    // forward: apply P L^-1
    double yi = y[0];
    for (int i = 0; i < n - 1; ++i) {
        const double yn = y[i + 1];
        if (!swp[i]) { y[i] = yi; yi = yn - ff[i] * yi; }
        else { y[i] = yn; yi = yi - ff[i] * yn; }
    }
    y[n - 1] = yi;
    // backward: U z = y
    double z1 = y[n - 1] * di[n - 1], z2 = 0.0;
    y[n - 1] = z1;
    for (int i = n - 2; i >= 0; --i) {
        const double du2 = (swp[i] && i < n - 2) ? b[i + 1] : 0.0;
        const double z = (y[i] - du[i] * z1 - du2 * z2) * di[i];
        y[i] = z;
        z2 = z1; z1 = z;
    }
}

// Leading `want` eigenvalues of T_n (descending) by multisection, all threads of the CTA.  eigenvalue i = smallest x with
// below(x) >= n - i.  With the values of an earlier, smaller T (have_prev) the search starts from them: by interlacing a
// leading eigenvalue can only grow with n, so round 0 places its points geometrically above the old value (a converged
// value is bracketed to a few ulps at once); otherwise round 0 spreads CH_THREADS points over the Gershgorin interval.
// Later rounds: P uniformly spaced points per eigenvalue, until no bracket moves (max_rounds = 0) or for max_rounds rounds
// (tracking: the brackets stay in cs.lo / cs.hi); resume = continue from the brackets a tracking call left.  Returns the
// norm bound tn.
// resume with max_rounds > 0: that many more rounds on the brackets a tracking call left.
__device__ double leading_eigenvalues(const CheckSmem& cs, int n, int want, bool have_prev, bool resume, int max_rounds) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double glo = 1e300, ghi = -1e300, tn = 0.0;
    for (int i = tid; i < n; i += CH_THREADS) {
        const double r = (i > 0 ? fabs(cs.b[i - 1]) : 0.0) + (i < n - 1 ? fabs(cs.b[i]) : 0.0);
        glo = fmin(glo, cs.a[i] - r); ghi = fmax(ghi, cs.a[i] + r); tn = fmax(tn, fabs(cs.a[i]) + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
        ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
        tn = fmax(tn, __shfl_xor_sync(0xffffffffu, tn, o));
    }
    __syncthreads();
    if (lane == 0) { cs.red[warp] = glo; cs.red[CH_WARPS + warp] = ghi; cs.red[2 * CH_WARPS + warp] = tn; }
    __syncthreads();
    glo = cs.red[0]; ghi = cs.red[CH_WARPS]; tn = cs.red[2 * CH_WARPS];
    for (int w = 1; w < CH_WARPS; ++w) { glo = fmin(glo, cs.red[w]); ghi = fmax(ghi, cs.red[CH_WARPS + w]); tn = fmax(tn, cs.red[2 * CH_WARPS + w]); }
    __syncthreads();
    const double span = ghi - glo;
    glo -= 1e-12 * span + 1e-300; ghi += 1e-12 * span + 1e-300;
    const double d0 = 4.0 * 2.220446049250313e-16 * tn;          // first geometric offset above an old eigenvalue
    if (!resume && tid < want) { cs.lo[tid] = have_prev ? fmax(glo, cs.prev[tid] - d0) : glo; cs.hi[tid] = ghi; }
    __syncthreads();
    // points per eigenvalue and round: the Sturm recurrences are FP64-throughput bound (measured: 512 of them cost 30k cycles
    // at n = 185), so a round uses 16 points per eigenvalue (6 warps for 11 values) rather than all threads
    const int P = min(CH_THREADS / want, CH_POINTS);
    const int round_end = max_rounds > 0 ? (resume ? 1 + max_rounds : max_rounds) : 80;
    for (int round = resume ? 1 : 0; round < round_end; ++round) {
        const bool global_scan = (round == 0 && !have_prev);
        double x = 0.0; int e = -1;
        if (global_scan) {
            x = glo + (ghi - glo) * ((double)(tid + 1) / (double)(CH_THREADS + 1));
        } else {
            e = tid / P;
            const int q = tid - e * P;
            if (e < want) {
                const double l = cs.lo[e], h = cs.hi[e];
                if (round == 0) x = fmin(l + ldexp(d0, 4 * q + 1), h);                       // lo + 2 d0 16^q
                else x = l + (h - l) * ((double)(q + 1) / (double)(P + 1));
            }
        }
        int c = 0;
        if (global_scan || e < want) c = sturm_below(cs.a, cs.b2, n, x);
        cs.cnt[tid] = c;
        cs.xs[tid] = x;
        __syncthreads();
        bool moved = false;
        if (tid < want) {                                            // thread e narrows bracket e (ascending points, monotone counts)
            const int target = n - tid;
            double l = cs.lo[tid], h = cs.hi[tid];
            const int p0 = global_scan ? 0 : tid * P, p1 = global_scan ? CH_THREADS : p0 + P;
            for (int q = p0; q < p1; ++q) {
                const double xq = cs.xs[q];
                if (!(xq > l && xq < h)) continue;
                if (cs.cnt[q] >= target) { h = xq; break; }
                l = xq;
            }
            moved = (l != cs.lo[tid]) || (h != cs.hi[tid]);
            cs.lo[tid] = l; cs.hi[tid] = h;
        }
        const int any = __syncthreads_or(moved ? 1 : 0);
        if (!any && round > 0) break;
    }
    if (tid < want) cs.theta[tid] = 0.5 * (cs.lo[tid] + cs.hi[tid]);
    __syncthreads();
    return tn;
}

// Unit eigenvectors of T_n for cs.theta[0..want) in cs.y by inverse iteration, one thread per vector (LAPACK dstein's scheme):
// ritz_prepare factors T - theta_i I and draws the start vectors; ritz_iterate does one solve for every vector, then (all
// threads) orthogonalises the vectors of a cluster |theta_i - theta_{i-1}| <= 1e-3 |T| against each other in index order
// (coefficients by one warp per earlier vector, then the update; skipped when ortho = false) and normalises them;
// ritz_residuals_ok tests |beta_n s_{n,i}| <= factor tol |theta_0| for the kw leading pairs (guard pairs: 1e-5).
__device__ void ritz_prepare(const CheckSmem& cs, int n, int cap, int want, double tn) {
    const int tid = threadIdx.x;
    const double tiny = fmax(tn * 2.2e-16, 1e-300);
    const double cluster_tol = 1e-3 * tn;
    if (tid < want) {
        int cstart = tid;
        while (cstart > 0 && fabs(cs.theta[cstart] - cs.theta[cstart - 1]) <= cluster_tol) --cstart;
        double x = cs.theta[tid];
        if (tid > cstart && fabs(x - cs.theta[tid - 1]) < 10.0 * tiny) x = cs.theta[tid - 1] - 10.0 * tiny * (tid - cstart);
        inverse_iteration_thread(cs, n, cap, tid, x, tiny);
        double* y = cs.y + (size_t)tid * cap;
        unsigned int seed = 12345u + 7919u * (unsigned int)tid;
        for (int r = 0; r < n; ++r) { seed = seed * 1664525u + 1013904223u; y[r] = ((double)(seed >> 8) / 16777216.0) - 0.5; }
    }
    __syncthreads();
}

// Clusters |theta_i - theta_{i-1}| <= 1e-3 |T| orthogonalised in index order (two passes, coefficients by one warp per earlier
// vector), every vector normalised.  All threads.
// unit_in: the vectors arrive normalised -- only the members of a cluster that were modified are normalised again.
__device__ void ritz_orthonormalise(const CheckSmem& cs, int n, int cap, int want, double tn, bool unit_in = false) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double cluster_tol = 1e-3 * tn;
    int cstart = 0;
    for (int i = 0; i < want; ++i) {
        if (i > 0 && fabs(cs.theta[i] - cs.theta[i - 1]) > cluster_tol) cstart = i;
        if (unit_in && i == cstart) continue;                           // uniform: theta is shared
        double* yi = cs.y + (size_t)i * cap;
        if (i > cstart) {
            for (int pass = 0; pass < 2; ++pass) {
                for (int pp = cstart + warp; pp < i; pp += CH_WARPS) {
                    const double* yp = cs.y + (size_t)pp * cap;
                    double dot = 0.0;
                    for (int r = lane; r < n; r += 32) dot = fma(yp[r], yi[r], dot);
                    dot = wsum(dot);
                    if (lane == 0) cs.dots[pp] = dot;
                }
                __syncthreads();
                for (int r = tid; r < n; r += CH_THREADS) {
                    double acc = yi[r];
                    for (int pp = cstart; pp < i; ++pp) acc = fma(-cs.dots[pp], cs.y[(size_t)pp * cap + r], acc);
                    yi[r] = acc;
                }
                __syncthreads();
            }
        }
        double nrm = 0.0;
        for (int r = tid; r < n; r += CH_THREADS) nrm = fma(yi[r], yi[r], nrm);
        nrm = wsum(nrm);
        if (lane == 0) cs.red[warp] = nrm;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < CH_WARPS; ++w) tot += cs.red[w];
        tot = sqrt(tot);
        if (!(tot > 0.0) || !(tot < 1e300)) {
            for (int r = tid; r < n; r += CH_THREADS) yi[r] = (r == (i % n)) ? 1.0 : 0.0;
        } else {
            const double inv = 1.0 / tot;
            for (int r = tid; r < n; r += CH_THREADS) yi[r] *= inv;
        }
        __syncthreads();
    }
}

__device__ void ritz_iterate(const CheckSmem& cs, int n, int cap, int want, double tn) {
    if (threadIdx.x < want) inverse_solve_thread(cs, n, cap, threadIdx.x);
    __syncthreads();
    ritz_orthonormalise(cs, n, cap, want, tn);
}

// Unit eigenvectors of T_n for cs.theta[0..want) in cs.y by the three-term recurrence run BACKWARDS, one thread per vector.  Row k of
// (T - theta) s = 0 gives s_{k-1} = -((a_k - theta) s_k + b_k s_{k+1}) / b_{k-1}; from s_n = 1 upwards the wanted solution grows (a
// converged Ritz vector is tiny at the bottom), so rounding errors stay relatively small; ~100 cycles per row and no division
// (reciprocals of the off-diagonals once, all threads), a fifth of the pivoted factorisation + solve of inverse iteration.  Row 1
// is not used: the result solves (T - theta) s = rho e_1 exactly (one inverse-iteration step from e_1), and |rho| / |s| IS its
// residual -- returned in cs.dots[i] relative to tn, so the caller knows how good each vector is (theta within a few ulps of |T| of
// a separated eigenvalue: ~1e-16).  cs.di[0..n), cs.ff[0..n), cs.du and cs.xs[0..want) are scratch here (ritz_prepare rewrites them).
constexpr int RR_MAXEV = 4;                                                 // rescaling events per vector (32 + want * 9 <= CH_THREADS doubles of cs.xs)
__device__ void ritz_recurrence(const CheckSmem& cs, int n, int cap, int want, double tn, long long* prof = nullptr) {
    const int tid = threadIdx.x;
    long long t0 = now();
#define RR_PROF(i) do { if (prof && tid == 0) { const long long t_ = now(); prof[i] += t_ - t0; t0 = t_; } } while (0)
    double* rb = cs.di;                                                      // [n]  1 / b_k
    double* Q = cs.ff;                                                       // [n]  q_k = -b_k / b_{k-1}   (b_{n-1} := 0: there is no s_n)
    double* P = cs.du;                                                       // [want][cap]  p_k = -(a_k - theta_i) / b_{k-1}
    for (int i = tid; i + 1 < n; i += CH_THREADS) rb[i] = cs.b[i] != 0.0 ? 1.0 / cs.b[i] : 0.0;
    __syncthreads();
    RR_PROF(0);
    for (int k = 1 + tid; k < n; k += CH_THREADS) Q[k] = (k < n - 1) ? -cs.b[k] * rb[k - 1] : 0.0;
    for (int idx = tid; idx < want * n; idx += CH_THREADS) {
        const int i = idx / n, k = idx - i * n;
        if (k >= 1) P[(size_t)i * cap + k] = -(cs.a[k] - cs.theta[i]) * rb[k - 1];
    }
    __syncthreads();
    RR_PROF(1);
    if (tid < want) {
        // s_{k-1} = p_k s_k + q_k s_{k+1}: ONE dependent FMA per row; eight rows at a time with their factors loaded first and their
        // results stored last (loads cannot be moved across the stores by the compiler, and a warp issues in order)
        const double* Pk = P + (size_t)tid * cap;
        double* y = cs.y + (size_t)tid * cap;
        double s_cur = 1.0, s_next = 0.0, nrm2 = 1.0;
        y[n - 1] = 1.0;
        double* ev = cs.xs + 32 + (2 * RR_MAXEV + 1) * tid;                  // rescaling events of this vector: (row, exponent) pairs, then the count
        int n_ev = 0;
        int k = n - 1;
        while (k >= 1) {
            const int cnt = min(8, k);
            double pr[8], qr[8], out[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { pr[u] = u < cnt ? Pk[k - u] : 0.0; qr[u] = u < cnt ? Q[k - u] : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double s_prev = fma(pr[u], s_cur, qr[u] * s_next);
                if (u < cnt) { s_next = s_cur; s_cur = s_prev; nrm2 = fma(s_prev, s_prev, nrm2); }
                out[u] = s_prev;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) if (u < cnt) y[k - 1 - u] = out[u];
            k -= cnt;
            if (k >= 1 && (!(fabs(s_cur) < CH_RR_BIG) || !(fabs(s_next) < CH_RR_BIG))) {
                // keep the magnitudes in range (a pair that converged long ago is tiny at the bottom: 1e-100 and less): continue
                // scaled by 2^-e and remember from which row on the stored entries owe that factor -- they get it in the parallel
                // scaling pass below (rescaling them here, one ldexp per stored row, cost more than the whole recurrence)
                int e;
                frexp(fmax(fabs(s_cur), fabs(s_next)), &e);
                s_cur = ldexp(s_cur, -e); s_next = ldexp(s_next, -e); nrm2 = ldexp(nrm2, -2 * e);
                y[k] = s_cur;                                                // row k was stored unscaled a moment ago; rows > k owe the factor
                if (n_ev < RR_MAXEV) { ev[2 * n_ev] = (double)k; ev[2 * n_ev + 1] = (double)e; }
                ++n_ev;
            }
        }
        const double x = cs.theta[tid];
        const double nr = sqrt(nrm2);
        const bool fine = nr > 0.0 && nr < 1e300 && n_ev <= RR_MAXEV;
        const double rho = n > 1 ? fma(cs.a[0] - x, y[0], cs.b[0] * y[1]) : (cs.a[0] - x);
        ev[2 * RR_MAXEV] = (double)min(n_ev, RR_MAXEV);
        cs.xs[tid] = fine ? 1.0 / nr : 0.0;
        cs.dots[tid] = fine ? fabs(rho) / (nr * fmax(tn, 1e-300)) : 1.0;
    }
    __syncthreads();
    RR_PROF(2);
    for (int idx = tid; idx < want * n; idx += CH_THREADS) {
        const int i = idx / n, r = idx - i * n;
        const double* ev = cs.xs + 32 + (2 * RR_MAXEV + 1) * i;
        const int n_ev = (int)ev[2 * RR_MAXEV];
        int e = 0;
        for (int j = 0; j < n_ev; ++j) if (r > (int)ev[2 * j]) e += (int)ev[2 * j + 1];     // rows below an event row were stored before it
        double v = cs.y[(size_t)i * cap + r] * cs.xs[i];
        if (e != 0) v = ldexp(v, -e);
        cs.y[(size_t)i * cap + r] = v;
    }
    __syncthreads();
    RR_PROF(3);
#undef RR_PROF
}

// The decision needs only the LAST component of every Ritz vector: the same recurrence without storing the vector (measured:
// 18 000 cycles at n = 165 against 56 000 with the stores and 84 000 for the pivoted factorisation + first solve of inverse
// iteration).  Result in cs.y[i * cap + n - 1], the only entry ritz_residuals_ok reads; cs.di[0..n) is scratch.
__device__ void ritz_last_components(const CheckSmem& cs, int n, int cap, int want) {
    const int tid = threadIdx.x;
    double* rb = cs.di;
    for (int i = tid; i + 1 < n; i += CH_THREADS) rb[i] = cs.b[i] != 0.0 ? 1.0 / cs.b[i] : 0.0;
    __syncthreads();
    if (tid < want) {
        const double x = cs.theta[tid];
        double s_cur = 1.0, carry = 0.0, last = 1.0, nrm2 = 1.0;             // carry = b_k s_{k+1}
        for (int k = n - 1; k >= 1; --k) {
            const double s_prev = -fma(cs.a[k] - x, s_cur, carry) * rb[k - 1];
            carry = cs.b[k - 1] * s_cur;
            s_cur = s_prev;
            nrm2 = fma(s_prev, s_prev, nrm2);
            if ((k & 7) == 0 && !(fabs(s_cur) < CH_RR_BIG)) {                // keep the magnitudes in range: everything by 2^-e
                int e;
                frexp(fmax(fabs(s_cur), fabs(carry)), &e);
                s_cur = ldexp(s_cur, -e); carry = ldexp(carry, -e); last = ldexp(last, -e); nrm2 = ldexp(nrm2, -2 * e);
            }
        }
        const double nr = sqrt(nrm2);
        cs.y[(size_t)tid * cap + (n - 1)] = (nr > 0.0 && nr < 1e300) ? last / nr : 1.0;
    }
    __syncthreads();
}

__device__ bool ritz_residuals_ok(const CheckSmem& cs, int n, int cap, int kw, int want, double tol, double factor) {
    const double scale = fmax(fabs(cs.theta[0]), 1e-300);
    const double beta_last = cs.b[n - 1];
    bool ok = (n >= kw);
    for (int i = 0; i < want; ++i) {
        const double res = fabs(beta_last * cs.y[(size_t)i * cap + (n - 1)]);
        if (!(res <= (i < kw ? tol : 1e-5) * factor * scale)) ok = false;
    }
    return ok;
}

__device__ void checker_main(const Params& p, double* smem, int c) {
    const int tid = threadIdx.x;
    const int cap = p.max_steps + 2;
    CheckSmem cs;
    double* q = smem;
    cs.a = q; q += cap; cs.b = q; q += cap; cs.b2 = q; q += cap;
    cs.theta = q; q += CH_MAXWANT; cs.lo = q; q += CH_MAXWANT; cs.hi = q; q += CH_MAXWANT;
    cs.prev = q; q += CH_MAXWANT; cs.dots = q; q += CH_MAXWANT;
    cs.red = q; q += 3 * CH_WARPS;
    const int wmax = min(p.kw + p.guard, CH_MAXWANT);
    cs.y = q; q += (size_t)wmax * cap; cs.di = q; q += (size_t)wmax * cap; cs.du = q; q += (size_t)wmax * cap; cs.ff = q; q += (size_t)wmax * cap;
    cs.xs = q; q += CH_THREADS;
    cs.cnt = reinterpret_cast<int*>(q); q += CH_THREADS / 2;
    cs.swp = reinterpret_cast<unsigned char*>(q);
    __shared__ int s_n;
    __shared__ int s_mode;
    int have = 0;                                              // alpha/beta of steps 1..have are in shared memory
    // The sizes tested are a fixed schedule (first_check + t check_stride; checker c takes t = c, c + n_checkers, ...) and the
    // OLDEST passing test stops the iteration, whichever checker finishes first: the step count of the result -- and with it
    // every bit of the output -- does not depend on how fast the checkers run next to the workers.
    int t_idx = c;
    int next_n = p.first_check + t_idx * p.check_stride;
    // A failed test of the eigenVECTORS (they cost 2 - 3 times an eigenvalue test) also gives up this checker's next slot, so that the
    // checkers keep up with the workers; the slot is marked failed for whoever waits for its verdict.
#define CH_TEST_FAILED(skip_one) do { \
        if (tid == 0 && mode == 1) { \
            if (t_idx < CH_MAXTESTS) st_word(&p.ctl->verdict[t_idx], 1u); \
            if ((skip_one) && t_idx + p.n_checkers < CH_MAXTESTS) st_word(&p.ctl->verdict[t_idx + p.n_checkers], 1u); \
        } \
        t_idx += ((skip_one) ? 2 : 1) * p.n_checkers; next_n = p.first_check + t_idx * p.check_stride; } while (0)
    int prev_want = 0;                                         // eigenvalues of this checker's previous test in cs.prev
    for (;;) {
        // ---- wait for the next size to test (thread 0 decides) ----
        if (tid == 0) {
            int mode = 0, n = 0;                               // 1 = test T_n, 2 = last test at max_steps, 4 = accept T_n, 3 = leave
            const long long t0 = now();
            for (;;) {
                if (ld_word(&p.ctl->stop) != 0u || ld_word(&p.ctl->abort) != 0u) { mode = 3; break; }
                const unsigned int fin = ld_word(&p.ctl->final_n);
                if (fin != 0u) {
                    if (c != 0) { mode = 3; break; }
                    n = (int)(fin & 0xffffu); mode = ((fin >> 30) & 1u) ? 4 : 2;       // 4 = breakdown: accept
                    if (n >= p.cap_steps) mode = 4;                                     // Krylov space exhausted: T_n is exact
                    break;
                }
                const int prog = (int)ld_word(&p.ctl->progress);
                if (prog >= next_n && t_idx < CH_MAXTESTS) { n = next_n; mode = 1; break; }
                if (now() - t0 > CH_TIMEOUT) { st_word(&p.ctl->abort, 1u); mode = 3; break; }
                relax();
            }
            s_n = n; s_mode = mode;
        }
        __syncthreads();
        const int n = s_n, mode = s_mode;
        if (mode == 3 || n < 1) return;
        // ---- fetch alpha/beta of the steps not yet seen ----
        int bad = 0;
        for (int jj = have + 1 + tid; jj <= n; jj += CH_THREADS) {
            double av, bv;
            if (!wait_ll(p.abuf + 2 * (size_t)jj, 1u, av, p.ctl) || !wait_ll(p.abuf + 2 * (size_t)jj + 1, 1u, bv, p.ctl)) { bad = 1; break; }
            cs.a[jj - 1] = av; cs.b[jj - 1] = bv; cs.b2[jj - 1] = bv * bv;
        }
        if (__syncthreads_or(bad)) return;
        have = n;
        const int want = min(p.kw + p.guard, n);
        const bool have_prev = prev_want >= want;
        const bool track = have_prev && mode == 1;
        // Tracking: cs.prev holds lower bounds of the leading eigenvalues from this checker's previous test (by interlacing
        // they only grow with n).  Two multisection rounds from them bracket every value to ~2 % of its growth; a converged
        // Ritz value no longer grows (its error is residual^2 / gap), so the eigenvectors are only worth computing once the
        // kw leading values have stayed within a few ulps of |T| above their old bounds.
        const long long tc0 = now();
        double tn = leading_eigenvalues(cs, n, want, have_prev, false, track ? 2 : 0);
        if (p.prof_on && tid == 0) p.ctl->prof[8 + 2 * c] += (1LL << 32) + (now() - tc0);      // tests, eigenvalue-tracking cycles
        // The gate in front of the eigenvector test: have the kw leading values grown by more than thr over the last check_back
        // steps?  T_{n - back} is a leading block of T_n, so one Sturm count per value on the shorter recurrence answers
        // "theta_i(n - back) >= lo_i(n) - thr" -- no: grown, whatever the bracket (lo is a lower bound); yes: the growth is at most
        // hi - lo + thr, and the bracket at n is refined (one round at a time) until that is decisive.  Self-contained: the answer
        // does not depend on how far back this checker's previous test lies.
        bool stagnant = track && n - p.check_back >= want;
        if (stagnant) {
            const double thr = fmax(64.0 * 2.220446049250313e-16, 1e3 * p.tol * p.tol) * tn;
            const int gk = min(p.kw, want), nb = n - p.check_back;
            for (int it = 0; it < 12; ++it) {
                int open = 0, grown = 0;
                if (tid < gk) {
                    const int below = sturm_below(cs.a, cs.b2, nb, cs.lo[tid] - thr);
                    if (nb - below < tid + 1) grown = 1;
                    else if (cs.hi[tid] - cs.lo[tid] > thr) open = 1;
                }
                const int any_grown = __syncthreads_or(grown);
                const int any_open = __syncthreads_or(open);
                if (any_grown) { stagnant = false; break; }
                if (!any_open) break;
                if (it == 11) { stagnant = false; break; }
                tn = leading_eigenvalues(cs, n, want, true, true, 1);
            }
        }
        CH_DEBUG_TEST(c, n, mode, track, stagnant);
        __syncthreads();
        if (tid < want) cs.prev[tid] = track ? cs.lo[tid] : cs.theta[tid];
        prev_want = want;
        __syncthreads();
        if (mode == 1 && !stagnant) { CH_TEST_FAILED(false); continue; }
        const long long tc1 = now();
        if (!track) { /* theta is final */ }
        else if (tid < want) cs.theta[tid] = 0.5 * (cs.lo[tid] + cs.hi[tid]);      // stagnant: the bracket is a few ulps of |T| wide
        __syncthreads();
        bool won = false;
        if (mode == 1) {
            // The decision needs the last component of every Ritz vector: the three-term recurrence run backwards (ritz_last_components).
            ritz_last_components(cs, n, cap, want);
            if (p.prof_on && tid == 0) p.ctl->prof[9 + 2 * c] += (1LL << 32) + (now() - tc1);  // vector tests, cycles up to the decision
            if (!ritz_residuals_ok(cs, n, cap, p.kw, want, p.tol, 0.25)) { CH_TEST_FAILED(true); __syncthreads(); continue; }
            if (tid == 0) {
                // passed: wait for the verdicts of the tests just before this one (one per other checker); an older pass wins
                bool lose = false;
                const long long t0 = now();
                for (int u = t_idx - 1; u >= 0 && u > t_idx - p.n_checkers && !lose; --u) {
                    for (;;) {
                        const unsigned int v = ld_word(&p.ctl->verdict[u]);
                        if (v == 2u || ld_word(&p.ctl->stop) != 0u || ld_word(&p.ctl->abort) != 0u) { lose = true; break; }
                        if (v == 1u) break;
                        if (now() - t0 > CH_TIMEOUT) { st_word(&p.ctl->abort, 1u); lose = true; break; }
                        relax();
                    }
                }
                if (!lose) {
                    st_word(&p.ctl->verdict[t_idx], 2u);
                    lose = cas_word(&p.ctl->stop, 0u, 0x80000000u | ((unsigned int)c << 16) | (unsigned int)n) != 0u;
                }
                s_mode = lose ? 0 : 1;
            }
            __syncthreads();
            won = s_mode != 0;
            __syncthreads();
            if (!won) return;                                               // another checker stopped the iteration first
            // The workers are leaving: two more multisection rounds on the eigenvalues (17^2 times tighter: the last ulp), then the
            // vectors for the final values by the same recurrence, stored this time.
            tn = leading_eigenvalues(cs, n, want, true, true, 2);
            ritz_recurrence(cs, n, cap, want, tn, p.prof_on ? &p.ctl->prof[26] : nullptr);
        }
        // The recurrence vectors are kept when each solves (T - theta) s = rho e_1 with |rho| <= 1e-12 |T| |s| (clusters orthogonalised as
        // dstein does); otherwise, and for the final tests at max_steps / breakdown, inverse iteration with the pivoted factorisation:
        // three solves from a random start.
        bool by_recurrence = (mode == 1);
        if (by_recurrence) {
            int bad_q = 0;
            if (p.prof_on && tid == 0) {
                double mw = 0.0, mg = 0.0;
                for (int i = 0; i < want; ++i) { if (i < p.kw) mw = fmax(mw, cs.dots[i]); else mg = fmax(mg, cs.dots[i]); }
                p.ctl->prof[24] = __double_as_longlong(mw); p.ctl->prof[25] = __double_as_longlong(mg);
            }
            if (tid < want && !(cs.dots[tid] <= (tid < p.kw ? 1e-12 : 1e-8))) bad_q = 1;   // guard pairs are not output: their test is the residual one
            by_recurrence = __syncthreads_or(bad_q) == 0;
        }
        if (by_recurrence) ritz_orthonormalise(cs, n, cap, want, tn, true);
        else {
            ritz_prepare(cs, n, cap, want, tn);
            for (int it = 0; it < 3; ++it) ritz_iterate(cs, n, cap, want, tn);
        }
        if (p.prof_on && tid == 0 && mode == 1) p.ctl->prof[7] += by_recurrence ? 1 : 1000;
        const bool ok = ritz_residuals_ok(cs, n, cap, p.kw, want, p.tol, 1.0);
        {
            double* th = p.theta_out + (size_t)c * CH_MAXWANT;
            double* S = p.S_out + (size_t)c * p.s_stride;
            if (tid < CH_MAXWANT) th[tid] = tid < want ? cs.theta[tid] : 0.0;
            for (int idx = tid; idx < n * p.kw; idx += CH_THREADS) {
                const int l = idx / p.kw, i = idx - l * p.kw;
                S[idx] = i < want ? cs.y[(size_t)i * cap + l] : 0.0;
            }
            publish_fence();
            __syncthreads();
            if (tid == 0) {
                const bool accept = (ok || mode == 4) && want >= p.kw;
                if (!won) won = cas_word(&p.ctl->stop, 0u, 0x80000000u | ((unsigned int)c << 16) | (unsigned int)n) == 0u;
                if (won) { p.ctl->n_final = n; p.ctl->winner = accept ? c : -1; p.ctl->status = accept ? 1 : 2; }
            }
        }
        return;
    }
#undef CH_TEST_FAILED
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CH_THREADS, 1) lanczos_chain_kernel(Params p) {
    extern __shared__ double ch_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    if (b >= p.n_workers) {
        if (b - p.n_workers < p.n_checkers) checker_main(p, ch_smem, b - p.n_workers);
        return;
    }
    const int m = p.m, nW = p.n_workers, cap = p.vs_cap;
    const int row0 = b * p.chunk, n_own = min(p.chunk, m - row0);           // 1 .. CH_MAXROWS
    const int n_sm = max(0, n_own - CH_REGROWS);                             // rows in shared memory; the rest in registers
    double* s_vec = ch_smem;                                                 // [m]   q_j
    double* s_rows = s_vec + m;                                              // [n_sm][m]
    double* Vs = s_rows + (size_t)n_sm * m;                                  // [n_own][cap]
    __shared__ double s_red[CH_WARPS * 16];
    __shared__ double s_w[16];
    __shared__ double s_upd[16];
    __shared__ double s_h[CH_THREADS + 8];
    __shared__ double s_sc[4];
    __shared__ unsigned int s_stop;
    Ctl* ctl = p.ctl;
    long long t_prof = now();
#define CH_PROF(slot) do { if (p.prof_on && b == 0 && tid == 0) { const long long t_ = now(); ctl->prof[slot] += t_ - t_prof; t_prof = t_; } } while (0)

    // ---- once: rows of Ms on chip, v0 / q1 on the owned rows, q1 everywhere ----
    for (int rr = 0; rr < n_sm; ++rr) {
        const double* a = p.Ms + (size_t)(row0 + rr) * m;
        for (int c = tid; c < m; c += CH_THREADS) s_rows[(size_t)rr * m + c] = __ldg(a + c);
    }
    double rrow[CH_REGROWS][4];
#pragma unroll
    for (int e = 0; e < CH_REGROWS; ++e)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = tid + CH_THREADS * u;
            rrow[e][u] = (n_sm + e < n_own && c < m) ? __ldg(p.Ms + (size_t)(row0 + n_sm + e) * m + c) : 0.0;
        }
    for (int c = tid; c < m; c += CH_THREADS) s_vec[c] = p.V[(size_t)m + c];
    if (tid < 2 * n_own) { const int r = tid >> 1, i = tid & 1; Vs[(size_t)r * cap + i] = p.V[(size_t)i * m + row0 + r]; }
    double sigma = 0.0, beta_prev = 0.0, alpha1 = 0.0;
    __syncthreads();
    CH_PROF(0);

    int j = 1;
    int end_reason = 0;                                                      // 1 stop word, 2 breakdown, 3 max_steps, 4 abort
    for (; j <= p.max_steps; ++j) {
        // ---- S0: w = Ms q_j - sigma q_j - beta_{j-1} q_{j-1} on the owned rows ----
        {
            double acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = tid + CH_THREADS * u;
                if (c < m) {
                    const double qv = s_vec[c];
#pragma unroll
                    for (int q = 0; q < CH_MAXROWS - CH_REGROWS; ++q)
                        if (q < n_sm) acc[q] = fma(s_rows[(size_t)q * m + c], qv, acc[q]);
#pragma unroll
                    for (int e = 0; e < CH_REGROWS; ++e) acc[CH_MAXROWS - CH_REGROWS + e] = fma(rrow[e][u], qv, acc[CH_MAXROWS - CH_REGROWS + e]);
                }
            }
            const double t = multi_sum16(acc, lane);
            if ((lane & 1) == 0) s_red[warp * 16 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = t;
            __syncthreads();
            if (tid < n_own) {
                const int slot = tid < n_sm ? tid : (CH_MAXROWS - CH_REGROWS) + (tid - n_sm);
                double tot = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < CH_WARPS; ++w2) tot += s_red[w2 * 16 + slot];
                const double qr = Vs[(size_t)tid * cap + j];
                double x = fma(-sigma, qr, tot);
                if (j >= 2) x = fma(-beta_prev, Vs[(size_t)tid * cap + j - 1], x);
                s_w[tid] = x;
            }
            __syncthreads();
        }
        CH_PROF(1);

        double alpha_j = sigma, beta_j = 0.0;
        bool leave = false;
        for (int pass = 0; pass < 2; ++pass) {
            const unsigned int tag = 2u * (unsigned int)j + (unsigned int)pass;
            const int n_slots = j + 2;                                       // slot 0: |w|^2, slot 1 + i: V_i . w, i = 0..j
            // ---- P1: partial coefficients of the owned rows ----
            for (int s = tid; s < n_slots; s += CH_THREADS) {
                double v = 0.0;
                if (s == 0) { for (int r = 0; r < n_own; ++r) v = fma(s_w[r], s_w[r], v); }
                else { for (int r = 0; r < n_own; ++r) v = fma(Vs[(size_t)r * cap + (s - 1)], s_w[r], v); }
                st_ll(p.pbuf + (size_t)s * nW + b, v, tag);
            }
            // Polling discipline (measured: with every thread of every CTA spinning on the same few lines the hand-overs cost
            // ~5k cycles each -- an L2 slice hot spot that also delays the stores being waited for): only the warps that need
            // a packet poll, with all their loads of a round in flight together and a pause between rounds.
            // ---- P2: slots owned by this CTA (warps 0..): fixed-order sum of all workers' partials ----
            int bad = 0;
            for (int s = b + warp * nW; s < n_slots && warp < CH_HWARP; s += CH_HWARP * nW) {
                double x[CH_PK];
                unsigned int pend = 0u;
#pragma unroll
                for (int q = 0; q < CH_PK; ++q) { x[q] = 0.0; if (lane + 32 * q < nW) pend |= 1u << q; }
                const uint4* src = p.pbuf + (size_t)s * nW + lane;
                const long long t0 = now();
                for (unsigned int rounds = 1; pend != 0u; ++rounds) {
                    const unsigned int cur = pend;
#pragma unroll
                    for (int q = 0; q < CH_PK; ++q)
                        if ((cur >> q) & 1u) { double v; if (ld_ll(src + 32 * q, tag, v)) { x[q] = v; pend &= ~(1u << q); } }
                    if (pend != 0u) {
                        relax();
                        if ((rounds & 63u) == 0u && (ld_word(&ctl->abort) != 0u || now() - t0 > CH_TIMEOUT)) { st_word(&ctl->abort, 1u); bad = 1; break; }
                    }
                }
                double v = x[0];
#pragma unroll
                for (int q = 1; q < CH_PK; ++q) v += x[q];
                v = wsum(v);
                if (lane == 0) {
                    st_ll(p.hbuf + s, v, tag);
                    if (s == 0) {                                            // CTA 0 forwards the stop word with slot 0
                        const unsigned int stop = ld_word(&ctl->stop) != 0u ? 1u : 0u;
                        st_ll(p.hbuf + (p.max_steps + 2), (double)stop, tag);
                    }
                }
            }
            // ---- P3: all coefficients, fetched by four warps (up to four packets per lane in flight) ----
            if (warp >= CH_HWARP && warp < CH_HWARP + 4) {
                const int t = tid - 32 * CH_HWARP;                           // 0..127
                unsigned int pend = 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) if (t + 128 * q < n_slots) pend |= 1u << q;
                if (t == 127) pend |= 1u << 4;                               // the control packet
                const long long t0 = now();
                for (unsigned int rounds = 1; pend != 0u; ++rounds) {
                    const unsigned int cur = pend;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if ((cur >> q) & 1u) { double v; if (ld_ll(p.hbuf + t + 128 * q, tag, v)) { s_h[t + 128 * q] = v; pend &= ~(1u << q); } }
                    if ((cur >> 4) & 1u) { double v; if (ld_ll(p.hbuf + (p.max_steps + 2), tag, v)) { s_stop = v != 0.0 ? 1u : 0u; pend &= ~(1u << 4); } }
                    if (pend != 0u) {
                        relax();
                        if ((rounds & 63u) == 0u && (ld_word(&ctl->abort) != 0u || now() - t0 > CH_TIMEOUT)) { st_word(&ctl->abort, 1u); bad = 1; break; }
                    }
                }
            }
            if (__syncthreads_or(bad)) { end_reason = 4; leave = true; break; }
            if (s_stop) { end_reason = 1; leave = true; break; }
            CH_PROF(2 + pass * 4);
            // update of the owned rows (warp r) and sum of squares of the coefficients (last warp)
            if (warp < n_own) {
                const double* vr = Vs + (size_t)warp * cap;
                double a0 = 0.0, a1 = 0.0;
                int i = lane;
                for (; i + 32 <= j; i += 64) { a0 = fma(vr[i], s_h[1 + i], a0); a1 = fma(vr[i + 32], s_h[1 + i + 32], a1); }
                if (i <= j) a0 = fma(vr[i], s_h[1 + i], a0);
                const double tot = wsum(a0 + a1);
                if (lane == 0) s_upd[warp] = s_w[warp] - tot;
            } else if (warp == CH_WARPS - 1) {
                double a0 = 0.0;
                for (int i = lane; i <= j; i += 32) a0 = fma(s_h[1 + i], s_h[1 + i], a0);
                a0 = wsum(a0);
                if (lane == 0) s_sc[0] = a0;
            }
            __syncthreads();
            const double ww = s_h[0], after = ww - s_sc[0];
            alpha_j += s_h[1 + j];
            if (tid < n_own) s_w[tid] = s_upd[tid];
            if (pass == 0 && after < 0.5 * ww) { __syncthreads(); continue; }   // DGKS: orthogonalise once more
            beta_j = after > 0.0 ? sqrt(after) : 0.0;
            break;
        }
        if (leave) break;
        if (j == 1) alpha1 = alpha_j;
        // ---- q_{j+1} on the owned rows -> Vs, global V, packets; alpha/beta for the checkers and the host ----
        const bool breakdown = !(beta_j > 1e-14 * fmax(fabs(alpha1), 1e-300));
        const double inv = beta_j > 0.0 ? 1.0 / beta_j : 0.0;
        __syncthreads();
        if (tid < n_own) {
            const double x = s_w[tid] * inv;
            Vs[(size_t)tid * cap + j + 1] = x;
            p.V[(size_t)(j + 1) * m + row0 + tid] = x;
            st_ll(p.qbuf + row0 + tid, x, 2u * (unsigned int)j);                 // lanes 0..n_own-1 of warp 0: one store instruction
        }
        if (b == 0 && tid == 32) {
            p.alpha[j - 1] = alpha_j; p.beta[j - 1] = beta_j;
            st_ll(p.abuf + 2 * (size_t)j, alpha_j, 1u);
            st_ll(p.abuf + 2 * (size_t)j + 1, beta_j, 1u);
            publish_fence();
            st_word(&ctl->progress, (unsigned int)j);
        }
        CH_PROF(3);
        if (breakdown) { end_reason = 2; ++j; break; }
        if (j == p.max_steps) { end_reason = 3; ++j; break; }
        // ---- P4: the whole next vector.  First one sentinel packet per producer (its last row; a producer stores all its
        // rows with one instruction), then every packet once; a packet that is still missing is waited for individually. ----
        int bad = 0;
        const unsigned int qtag = 2u * (unsigned int)j;
        for (int c = tid; c < nW; c += CH_THREADS) {
            double x;
            if (!wait_ll(p.qbuf + min(c * p.chunk + p.chunk - 1, m - 1), qtag, x, ctl)) bad = 1;
        }
        if (__syncthreads_or(bad)) { end_reason = 4; ++j; break; }
        {
            double x[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = tid + CH_THREADS * u;
                ok[u] = true; x[u] = 0.0;
                if (c < m) ok[u] = ld_ll(p.qbuf + c, qtag, x[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = tid + CH_THREADS * u;
                if (c < m) {
                    if (!ok[u] && !wait_ll(p.qbuf + c, qtag, x[u], ctl)) bad = 1;
                    s_vec[c] = x[u];
                }
            }
        }
        if (__syncthreads_or(bad)) { end_reason = 4; ++j; break; }
        sigma = alpha_j; beta_prev = beta_j;
        CH_PROF(4);
    }
    // j - 1 = steps completed
    if (b == 0 && tid == 0) {
        ctl->steps_run = j - 1;
        if (end_reason == 2 || end_reason == 3) {
            publish_fence();
            st_word(&ctl->final_n, 0x80000000u | (end_reason == 2 ? 0x40000000u : 0u) | (unsigned int)(j - 1));
        }
    }
#undef CH_PROF
}

// ---------------------------------------------------------------------------------------------------------------------
// Cluster schedule.  The flat kernel above pays 5 000 - 10 000 cycles for every exchange 143 CTAs take part in
// (profiles/r2_handover_microbench.md).  Here the CTAs form thread-block clusters of S (8) and every exchange has two
// levels: distributed shared memory inside the cluster (st.async: the store itself completes bytes on the consumer's
// mbarrier, ~250 cycles per hop), flag-carrying packets through L2 between the cl_count (16) clusters only.
//   A1  thread s: partial coefficient of the owned rows  -> DSMEM of the CTA (rank s mod W) that owns slot s in this cluster
//   A2  owner: adds the W partials in rank order          -> packet cbuf[s][cluster]
//   A3  owner: reads the packets of ALL clusters for its slots (one lane per cluster), adds them in a fixed shuffle tree
//       (every cluster computes the same bits)            -> DSMEM s_h[s] of the W CTAs of the cluster
//   Q   new vector: own rows -> DSMEM of the W CTAs and packets qbuf; rank r reads the r-th share of the rows of the OTHER
//       clusters from the packets and forwards them       -> DSMEM s_vec of the W CTAs
// Slot j + 2 of the coefficient exchange carries the stop word (CTA 0 contributes it, everybody else 0).
// A buffer that peers write (s_part, s_h, s_vec) cannot be overwritten early: whatever a peer sends for the next phase
// depends on a total that needs this CTA's contribution to the next phase, which follows its last read in program order.
// ---------------------------------------------------------------------------------------------------------------------
// Dynamic shared memory of a worker of the cluster kernel, in doubles: coefficients | partials | vector | rows of Ms | basis rows.
__host__ __device__ inline int cl_round16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline size_t cluster_worker_doubles(int chunk, int cap) {
    (void)chunk;                                                             // the basis always has CL_MAXROWS (zero-padded) rows
    return (size_t)cl_round16(cap + 8) + (size_t)cl_round16(cap + 24) + (size_t)CL_LD * (1 + CL_SMROWS) + (size_t)CL_MAXROWS * cap;
}

__device__ __forceinline__ bool wait_mbar(Mbar* bar, unsigned int parity, Ctl* ctl) {
    if (mbar_test(bar, parity)) return true;
    const long long t0 = now();
    for (unsigned int polls = 1;; ++polls) {
        if (mbar_test(bar, parity)) return true;
        if ((polls & 255u) == 0u) {
            if (ld_word(&ctl->abort) != 0u) return false;
            if (now() - t0 > CH_TIMEOUT) { st_word(&ctl->abort, 1u); return false; }
        }
    }
}

__global__ void __launch_bounds__(CH_THREADS, 1) lanczos_cluster_kernel(Params p) {
    extern __shared__ __align__(128) double ch_smem[];
    __shared__ Mbar s_bar[4];                                                // 0 partial coefficients | 1 coefficients | 2 vector
    __shared__ double s_red[CH_WARPS * CL_MAXROWS];
    __shared__ double s_w[CL_MAXROWS + 1];
    __shared__ double s_upd[CL_MAXROWS + 1];
    __shared__ double s_sc[4];
    __shared__ long long s_prof[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int S = p.cl_size, C = p.cl_count, G = p.cl_group;
    const int rank = cluster_rank(), cid = b / S;
    Ctl* ctl = p.ctl;
    if (p.trace && b == 0 && tid == 0) p.trace[150 * 16] = wall_ns();
    if (tid == 0) { mbar_init(&s_bar[0]); mbar_init(&s_bar[1]); mbar_init(&s_bar[2]); mbar_init_fence(); }
    __syncthreads();
    cluster_sync_all();
    if (b >= p.n_workers) {
        if (b - p.n_workers < p.n_checkers) checker_main(p, ch_smem, b - p.n_workers);
        if (p.trace && tid == 0 && b - p.n_workers < p.n_checkers) p.trace[150 * 16 + 4 + (b - p.n_workers)] = wall_ns();
        cluster_sync_all();
        return;
    }
    const int W = min(S, p.n_workers - cid * S);                             // workers in this cluster
    const int m = p.m, cap = p.vs_cap;
    const int row0 = b * p.chunk, n_own = min(p.chunk, m - row0);            // 1 .. CL_MAXROWS
    const int n_sm = max(0, n_own - CL_REGROWS);
    const int crow0 = cid * S * p.chunk, crow1 = min(m, crow0 + S * p.chunk);   // rows owned inside this cluster
    const int per = (m + W - 1) / W, f_lo = rank * per, f_hi = min(m, f_lo + per);   // share of the vector this CTA forwards
    // Rows and vector are padded with zeros to CL_SMROWS x CL_LD, the basis to CL_MAXROWS rows: no bounds inside the step.
    double* s_h = ch_smem;                                                   // [cap + 8]   coefficients of the step   } written by the peers:
    double* s_part = s_h + cl_round16(cap + 8);                              // [cap + 24]  [owned slot][source rank]  } same offsets in
    double* s_vec = s_part + cl_round16(cap + 24);                           // [CL_LD]     q_j                        } every CTA
    double* s_rows = s_vec + CL_LD;                                          // [CL_SMROWS][CL_LD]
    double* Vs = s_rows + (size_t)CL_SMROWS * CL_LD;                         // [CL_MAXROWS][cap]
    long long t_prof = now();
    // profile of thread 0 of CTA 0, kept in shared memory during the solve (a global read-modify-write per probe would cost more than the phases)
#define CL_PROF(slot) do { if (p.prof_on && b == 0 && tid == 0) { const long long t_ = now(); s_prof[slot] += t_ - t_prof; t_prof = t_; } } while (0)
    if (tid < 32) s_prof[tid] = 0;
#define CL_TRACE(k) do { if (p.trace && tid == 0 && j == p.trace_step) p.trace[b * 16 + (k)] = wall_ns(); } while (0)
    if (tid <= CL_MAXROWS) { s_w[tid] = 0.0; s_upd[tid] = 0.0; }

    for (int rr = 0; rr < CL_SMROWS; ++rr) {
        const double* a = p.Ms + (size_t)(row0 + rr) * m;
        for (int c = tid; c < CL_LD; c += CH_THREADS) s_rows[rr * CL_LD + c] = (rr < n_sm && c < m) ? __ldg(a + c) : 0.0;
    }
    double rrow[CL_REGROWS][4];
#pragma unroll
    for (int e = 0; e < CL_REGROWS; ++e)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = tid + CH_THREADS * u;
            rrow[e][u] = (n_sm + e < n_own && c < m) ? __ldg(p.Ms + (size_t)(row0 + n_sm + e) * m + c) : 0.0;
        }
    for (int c = tid; c < CL_LD; c += CH_THREADS) s_vec[c] = c < m ? p.V[(size_t)m + c] : 0.0;
    for (int i = tid; i < CL_MAXROWS * cap; i += CH_THREADS) Vs[i] = 0.0;
    __syncthreads();
    if (tid < 2 * n_own) { const int r = tid >> 1, i = tid & 1; Vs[(size_t)r * cap + i] = p.V[(size_t)i * m + row0 + r]; }
    double sigma = 0.0, beta_prev = 0.0, alpha1 = 0.0;
    unsigned int phA = 0u, phH = 0u, phQ = 0u;
    __syncthreads();
    CL_PROF(0);
    if (p.trace && b == 0 && tid == 0) p.trace[150 * 16 + 1] = wall_ns();

    int j = 1;
    int end_reason = 0;                                                      // 1 stop word, 2 breakdown, 3 max_steps, 4 abort
    for (; j <= p.max_steps; ++j) {
        if (b == 0 && tid == CH_THREADS - 1) s_sc[1] = ld_word(&ctl->stop) != 0u ? 1.0 : 0.0;   // consumed after the barriers of S0
        CL_TRACE(0);
        // ---- S0: w = Ms q_j - sigma q_j - beta_{j-1} q_{j-1} on the owned rows ----
        {
            double acc[CL_MAXROWS];
#pragma unroll
            for (int q = 0; q < CL_MAXROWS; ++q) acc[q] = 0.0;
            const double* rp = s_rows + tid;
            const double* vp = s_vec + tid;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double qv = vp[CH_THREADS * u];
                double t[CL_SMROWS];
#pragma unroll
                for (int q = 0; q < CL_SMROWS; ++q) t[q] = rp[q * CL_LD + CH_THREADS * u];
#pragma unroll
                for (int q = 0; q < CL_SMROWS; ++q) acc[q] = fma(t[q], qv, acc[q]);
#pragma unroll
                for (int e = 0; e < CL_REGROWS; ++e) acc[CL_SMROWS + e] = fma(rrow[e][u], qv, acc[CL_SMROWS + e]);
            }
            CL_PROF(19);
            CL_TRACE(1);
            const double t16 = wsum(acc[16]);
            double (&acc16)[16] = reinterpret_cast<double (&)[16]>(acc);
            const double t = multi_sum16(acc16, lane);
            if ((lane & 1) == 0) s_red[warp * CL_MAXROWS + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = t;
            if (lane == 0) s_red[warp * CL_MAXROWS + 16] = t16;
            CL_PROF(20);
            __syncthreads();
            CL_PROF(21);
            if (tid < n_own) {
                const int slot = tid < n_sm ? tid : (CL_MAXROWS - CL_REGROWS) + (tid - n_sm);
                double tot = 0.0, tot2 = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < CH_WARPS; w2 += 2) { tot += s_red[w2 * CL_MAXROWS + slot]; tot2 += s_red[(w2 + 1) * CL_MAXROWS + slot]; }
                tot += tot2;
                const double qr = Vs[(size_t)tid * cap + j];
                double x = fma(-sigma, qr, tot);
                if (j >= 2) x = fma(-beta_prev, Vs[(size_t)tid * cap + j - 1], x);
                s_w[tid] = x;
            }
            __syncthreads();
        }
        CL_PROF(1);

        CL_TRACE(2);
        double alpha_j = sigma, beta_j = 0.0;
        bool leave = false;
        for (int pass = 0; pass < 2; ++pass) {
            const unsigned int tag = 2u * (unsigned int)j + (unsigned int)pass;
            const int n_tot = j + 3;                                         // slot 0: |w|^2, slot 1 + i: V_i . w (i = 0..j), slot j + 2: stop word
            const int n_mine = rank < n_tot ? (n_tot - rank + W - 1) / W : 0;   // slots rank, rank + W, ... are added up by this CTA
            if (tid == CH_THREADS - 1) { mbar_expect(&s_bar[0], 8u * (unsigned int)(n_mine * W)); mbar_expect(&s_bar[1], 8u * (unsigned int)n_tot); }
            // ---- A1: partial coefficients of the owned rows -> the owner of the slot in this cluster ----
            if (tid < n_tot) {
                const int s = tid;
                double v = 0.0;
                if (s == n_tot - 1) v = (b == 0) ? s_sc[1] : 0.0;
                else {                                                       // rows beyond n_own hold zeros
                    const double* vcol = s == 0 ? s_w : Vs + (s - 1);
                    const int vstep = s == 0 ? 1 : cap;
                    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
#pragma unroll
                    for (int r = 0; r + 2 < CL_MAXROWS; r += 3) {
                        v0 = fma(vcol[r * vstep], s_w[r], v0);
                        v1 = fma(vcol[(r + 1) * vstep], s_w[r + 1], v1);
                        v2 = fma(vcol[(r + 2) * vstep], s_w[r + 2], v2);
                    }
#pragma unroll
                    for (int r = CL_MAXROWS - CL_MAXROWS % 3; r < CL_MAXROWS; ++r) v0 = fma(vcol[r * vstep], s_w[r], v0);
                    v = (v0 + v1) + v2;
                }
                const int idx = s / W, owner = s - idx * W;
                dsm_send(&s_part[idx * W + rank], owner, v, &s_bar[0]);
            }
            if (pass == 0) CL_TRACE(3);
            int bad = 0;
            // ---- A2: cluster totals of the owned slots -> packets ----
            if (!wait_mbar(&s_bar[0], phA, ctl)) bad = 1;
            phA ^= 1u;
            if (pass == 0) CL_PROF(16);
            if (pass == 0) CL_TRACE(4);
            if (tid < n_mine && !bad) {
                double tot = s_part[tid * W];
                for (int src = 1; src < W; ++src) tot += s_part[tid * W + src];
                st_ll(p.cbuf + (size_t)(tid * W + rank) * C + cid, tot, tag);
            }
            // ---- A3: totals over all clusters (G lanes per slot, lane c reads cluster c), broadcast inside the cluster ----
            if (pass == 0) CL_TRACE(5);
            // every thread first has all its packets in flight (up to CL_GPK), then the totals are formed group by group
            const int per_pass = CH_THREADS / G;
            const int c = tid & (G - 1);
            for (int i0 = 0; i0 < n_mine; i0 += per_pass * CL_GPK) {
                double x[CL_GPK];
                unsigned int pend = 0u;
#pragma unroll
                for (int q = 0; q < CL_GPK; ++q) { x[q] = 0.0; if (i0 + q * per_pass + tid / G < n_mine && c < C) pend |= 1u << q; }
                const uint4* src = p.cbuf + (size_t)((i0 + tid / G) * W + rank) * C + c;
                const size_t src_step = (size_t)per_pass * W * C;
                const long long t0 = now();
                for (unsigned int rounds = 1; pend != 0u; ++rounds) {
                    const unsigned int cur = pend;
#pragma unroll
                    for (int q = 0; q < CL_GPK; ++q)
                        if ((cur >> q) & 1u) { double v; if (ld_ll(src + q * src_step, tag, v)) { x[q] = v; pend &= ~(1u << q); } }
                    if (pend != 0u) {
                        relax_short();
                        if ((rounds & 63u) == 0u && (ld_word(&ctl->abort) != 0u || now() - t0 > CH_TIMEOUT)) { st_word(&ctl->abort, 1u); bad = 1; break; }
                    }
                }
#pragma unroll
                for (int q = 0; q < CL_GPK; ++q) {
                    if (i0 + q * per_pass >= n_mine) break;                  // uniform in the CTA
                    double v = x[q];
                    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    const int idx = i0 + q * per_pass + tid / G;
                    if (idx < n_mine)
                        for (int d = c; d < W; d += G) dsm_send(&s_h[idx * W + rank], d, v, &s_bar[1]);
                }
            }
            if (pass == 0) CL_TRACE(6);
            if (pass == 0) CL_PROF(17);
            if (!wait_mbar(&s_bar[1], phH, ctl)) bad = 1;
            phH ^= 1u;
            if (pass == 0) CL_TRACE(7);
            if (__syncthreads_or(bad)) { end_reason = 4; leave = true; break; }
            if (s_h[n_tot - 1] != 0.0) { end_reason = 1; leave = true; break; }
            CL_PROF(2 + pass * 4);
            // update of the owned rows (warp r) and sum of squares of the coefficients (last warp)
            {
                const int task = tid >> 4, l16 = tid & 15;                   // 32 half-warps: rows 0 .. n_own - 1, the last one the squares
                double a0 = 0.0, a1 = 0.0;
                if (task < n_own) {
                    const double* vr = Vs + (size_t)task * cap;
                    int i = l16;
                    for (; i + 16 <= j; i += 32) { a0 = fma(vr[i], s_h[1 + i], a0); a1 = fma(vr[i + 16], s_h[1 + i + 16], a1); }
                    if (i <= j) a0 = fma(vr[i], s_h[1 + i], a0);
                } else if (task == 31) {
                    int i = l16;
                    for (; i + 16 <= j; i += 32) { a0 = fma(s_h[1 + i], s_h[1 + i], a0); a1 = fma(s_h[1 + i + 16], s_h[1 + i + 16], a1); }
                    if (i <= j) a0 = fma(s_h[1 + i], s_h[1 + i], a0);
                }
                double tot = a0 + a1;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                if (l16 == 0) {
                    if (task < n_own) s_upd[task] = s_w[task] - tot;
                    else if (task == 31) s_sc[0] = tot;
                }
            }
            if (pass == 0) CL_TRACE(8);
            if (pass == 0) CL_PROF(22);
            __syncthreads();
            if (pass == 0) CL_PROF(23);
            const double ww = s_h[0], after = ww - s_sc[0];
            alpha_j += s_h[1 + j];
            if (tid < n_own) s_w[tid] = s_upd[tid];
            if (pass == 0 && after < 0.5 * ww) { __syncthreads(); continue; }   // DGKS: orthogonalise once more
            beta_j = after > 0.0 ? sqrt(after) : 0.0;
            break;
        }
        if (leave) break;
        if (j == 1) alpha1 = alpha_j;
        // ---- q_{j+1} on the owned rows -> Vs, global V, the CTAs of this cluster, packets for the other clusters ----
        const bool breakdown = !(beta_j > 1e-14 * fmax(fabs(alpha1), 1e-300));
        const bool last = breakdown || j == p.max_steps;                     // nobody waits for this vector: no sends to CTAs that leave
        const double inv = beta_j > 0.0 ? 1.0 / beta_j : 0.0;
        __syncthreads();
        CL_TRACE(9);
        if (tid == CH_THREADS - 1 && !last) mbar_expect(&s_bar[2], 8u * (unsigned int)m);
        if (tid < n_own) {
            const double x = s_w[tid] * inv;
            Vs[(size_t)tid * cap + j + 1] = x;
            p.V[(size_t)(j + 1) * m + row0 + tid] = x;
            if (!last) st_ll(p.qbuf + row0 + tid, x, 2u * (unsigned int)j);
        }
        if (!last && tid < n_own * W) {
            const int r = tid / W, d = tid - r * W;
            dsm_send(&s_vec[row0 + r], d, s_w[r] * inv, &s_bar[2]);
        }
        if (b == 0 && tid == 32) {
            p.alpha[j - 1] = alpha_j; p.beta[j - 1] = beta_j;
            st_ll(p.abuf + 2 * (size_t)j, alpha_j, 1u);
            st_ll(p.abuf + 2 * (size_t)j + 1, beta_j, 1u);
            st_word(&ctl->progress, (unsigned int)j);                        // a hint: the checkers wait for the packets themselves
        }
        CL_TRACE(10);
        CL_PROF(3);
        if (breakdown) { end_reason = 2; ++j; break; }
        if (j == p.max_steps) { end_reason = 3; ++j; break; }
        // ---- Q: this CTA's share of the other clusters' rows, forwarded to the CTAs of this cluster ----
        int bad = 0;
        const unsigned int qtag = 2u * (unsigned int)j;
        for (int c = f_lo + tid; c < f_hi; c += CH_THREADS) {
            if (c >= crow0 && c < crow1) continue;
            double x;
            if (!wait_ll_short(p.qbuf + c, qtag, x, ctl)) { bad = 1; continue; }
            for (int d = 0; d < W; ++d) dsm_send(&s_vec[c], d, x, &s_bar[2]);
        }
        CL_TRACE(11);
        CL_PROF(18);
        if (!wait_mbar(&s_bar[2], phQ, ctl)) bad = 1;
        phQ ^= 1u;
        CL_TRACE(12);
        if (__syncthreads_or(bad)) { end_reason = 4; ++j; break; }
        sigma = alpha_j; beta_prev = beta_j;
        CL_PROF(4);
    }
    if (p.trace && b == 0 && tid == 0) p.trace[150 * 16 + 2] = wall_ns();
    if (p.prof_on && b == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) { ctl->prof[i] += s_prof[i]; ctl->prof[16 + i] += s_prof[16 + i]; }
    if (b == 0 && tid == 0) {
        ctl->steps_run = j - 1;
        if (end_reason == 2 || end_reason == 3) {
            publish_fence();
            st_word(&ctl->final_n, 0x80000000u | (end_reason == 2 ? 0x40000000u : 0u) | (unsigned int)(j - 1));
        }
    }
    cluster_sync_all();                                                      // no CTA leaves while a peer may still store into it
#undef CL_PROF
#undef CL_TRACE
}

// ---------------------------------------------------------------------------------------------------------------------
// Merged schedule (lanczos_merged_kernel): ONE exchange per step instead of two.  The cluster kernel above exchanges the
// coefficients h = V^T w (reduce), orthogonalises, and then exchanges the new vector q_{j+1} (gather) before the next
// product.  Here the UNorthogonalised candidate w_j is gathered together with the reduction of its coefficients, and the
// product is taken of w_j: by linearity
//     q_{j+1} = (w_j - V h) / beta_j                      Ms q_{j+1} = (Ms w_j - (Ms V) h) / beta_j
// and Ms V is known in the basis itself: Ms v_0 = lambda0 v_0 (locked pair), Ms q_i = beta_i q_{i+1} + alpha_i q_i +
// beta_{i-1} q_{i-1} for the earlier steps (their tiny reorthogonalisation terms enter multiplied by the tiny h_i: second
// order), and for i = j EXACTLY  Ms q_j = beta_j q_{j+1} + sigma q_j + beta_{j-1} q_{j-1} + V h  (the definition of this
// step).  So  Ms q_{j+1} = (y - V g) / beta_j - h_j q_{j+1}  with  y = Ms w_j  and  g = T~ h + h_j h  (T~ = tridiag(beta, alpha,
// beta) with sigma in place of alpha_j, lambda0 in row 0) -- one more pass over the owned basis rows, no storage.  The next
// candidate w_{j+1} = Ms q_{j+1} - alpha_j q_{j+1} - beta_j q_j and its partial coefficients follow on the owned rows at once.
// Numerically this is the same iteration (numpy model on the benchmark matrix: alpha, beta within 1e-15, the recurrence for
// Ms q stays within 1e-13 of the direct product over 320 steps; the locked pair must be an eigenpair: lambda0 wrong -> unstable).
//   per step:  partial coefficients -> DSMEM | candidate rows -> DSMEM + packets | owners: cluster totals -> packets |
//              forward the other clusters' rows | owners: totals over the clusters -> DSMEM | vector complete: copy out, "buffer
//              free" arrive to the peers, product y = Ms w_j | coefficients complete: g, update, next candidate
// The vector buffer is single: peers write the next candidate into it only after every CTA of the cluster has copied the
// current one out (count-based mbarrier s_bar[3], remote arrives).  DGKS second pass (rare): coefficients of w - V h are
// exchanged alone, H = h + h2, and the update is redone from w_j and y with H.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t merged_worker_doubles(int cap) {
    return (size_t)2 * cl_round16(cap + 8) + (size_t)cl_round16(cap + 24) + (size_t)CL_LD * (1 + CL_SMROWS) + (size_t)CL_MAXROWS * cap;
}

// y = Ms[own rows, :] x for the vector x held as 4 values per thread (columns tid + 512 u): rows in shared memory and registers.
__device__ __forceinline__ void merged_symv(const double* s_rows, const double (&rrow)[CL_REGROWS][4], const double (&qv)[4], double* s_red, int tid, int warp, int lane) {
    double acc[CL_MAXROWS];
#pragma unroll
    for (int q = 0; q < CL_MAXROWS; ++q) acc[q] = 0.0;
    const double* rp = s_rows + tid;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        double t[CL_SMROWS];
#pragma unroll
        for (int q = 0; q < CL_SMROWS; ++q) t[q] = rp[q * CL_LD + CH_THREADS * u];
#pragma unroll
        for (int q = 0; q < CL_SMROWS; ++q) acc[q] = fma(t[q], qv[u], acc[q]);
#pragma unroll
        for (int e = 0; e < CL_REGROWS; ++e) acc[CL_SMROWS + e] = fma(rrow[e][u], qv[u], acc[CL_SMROWS + e]);
    }
    const double t16 = wsum(acc[16]);
    double (&acc16)[16] = reinterpret_cast<double (&)[16]>(acc);
    const double t = multi_sum16(acc16, lane);
    if ((lane & 1) == 0) s_red[warp * CL_MAXROWS + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = t;
    if (lane == 0) s_red[warp * CL_MAXROWS + 16] = t16;
}
// second half (after a __syncthreads): the row totals of the 16 warps -> s_y
__device__ __forceinline__ void merged_symv_totals(const double* s_red, double* s_y, int n_own, int n_sm, int tid) {
    if (tid < n_own) {
        const int slot = tid < n_sm ? tid : (CL_MAXROWS - CL_REGROWS) + (tid - n_sm);
        // a dependent FP64 add costs ~50 cycles here: pairwise tree (depth 4) instead of two chains of 8
        double t[CH_WARPS];
#pragma unroll
        for (int w2 = 0; w2 < CH_WARPS; ++w2) t[w2] = s_red[w2 * CL_MAXROWS + slot];
#pragma unroll
        for (int st = 1; st < CH_WARPS; st *= 2)
#pragma unroll
            for (int w2 = 0; w2 + st < CH_WARPS; w2 += 2 * st) t[w2] += t[w2 + st];
        s_y[tid] = t[0];
    }
}

__global__ void __launch_bounds__(CH_THREADS, 1) lanczos_merged_kernel(Params p) {
    extern __shared__ __align__(128) double ch_smem[];
    __shared__ Mbar s_bar[4];                                                // 0 partial coefficients | 1 coefficients | 2 vector | 3 vector buffer free
    __shared__ double s_red[CH_WARPS * CL_MAXROWS];
    __shared__ double s_w[CL_MAXROWS + 1];                                   // own rows of the gathered candidate w_j
    __shared__ double s_y[CL_MAXROWS + 1];                                   // own rows of Ms w_j
    __shared__ double s_upd[CL_MAXROWS + 1];                                 // w_j - V h
    __shared__ double s_z[CL_MAXROWS + 1];                                   // y - V g
    __shared__ double s_nw[CL_MAXROWS + 1];                                  // own rows of the next candidate
    __shared__ double s_sc[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int S = p.cl_size, C = p.cl_count, G = p.cl_group;
    const int rank = cluster_rank(), cid = b / S;
    Ctl* ctl = p.ctl;
    const int W = min(S, p.n_workers - cid * S);                             // workers in this cluster (<= 0 in a cluster of checkers)
    if (p.trace && b == 0 && tid == 0) p.trace[150 * 16] = wall_ns();
    if (tid == 0) { mbar_init(&s_bar[0]); mbar_init(&s_bar[1]); mbar_init(&s_bar[2]); mbar_init_count(&s_bar[3], (unsigned int)max(W, 1)); mbar_init_fence(); }
    __syncthreads();
    cluster_sync_all();
    if (b >= p.n_workers) {
        if (b - p.n_workers < p.n_checkers) checker_main(p, ch_smem, b - p.n_workers);
        if (p.trace && tid == 0 && b - p.n_workers < p.n_checkers) p.trace[150 * 16 + 4 + (b - p.n_workers)] = wall_ns();
        cluster_sync_all();
        return;
    }
    const int m = p.m, cap = p.vs_cap;
    const int row0 = b * p.chunk, n_own = min(p.chunk, m - row0);            // 1 .. CL_MAXROWS
    const int n_sm = max(0, n_own - CL_REGROWS);
    const int crow0 = cid * S * p.chunk, crow1 = min(m, crow0 + S * p.chunk);   // rows owned inside this cluster
    const int per = (m + W - 1) / W, f_lo = rank * per, f_hi = min(m, f_lo + per);   // share of the vector this CTA forwards
    double* s_h = ch_smem;                                                   // [cap + 8]   coefficients of the step   } written by the peers:
    double* s_part = s_h + cl_round16(cap + 8);                              // [cap + 24]  [owned slot][source rank]  } same offsets in
    double* s_vec = s_part + cl_round16(cap + 24);                           // [CL_LD]     gathered candidate w_j     } every CTA
    double* s_g = s_vec + CL_LD;                                             // [cap + 8]   g (local)
    double* s_rows = s_g + cl_round16(cap + 8);                              // [CL_SMROWS][CL_LD]
    double* Vs = s_rows + (size_t)CL_SMROWS * CL_LD;                         // [CL_MAXROWS][cap]
#define MG_TRACE(k) do { if (p.trace && tid == 0 && j == p.trace_step) p.trace[b * 16 + (k)] = wall_ns(); } while (0)
#define MG_TRACE_ANY(k) do { if (p.trace && j == p.trace_step) p.trace[b * 16 + (k)] = wall_ns(); } while (0)
    if (tid <= CL_MAXROWS) { s_w[tid] = 0.0; s_y[tid] = 0.0; s_upd[tid] = 0.0; s_z[tid] = 0.0; s_nw[tid] = 0.0; }

    for (int rr = 0; rr < CL_SMROWS; ++rr) {
        const double* a = p.Ms + (size_t)(row0 + rr) * m;
        for (int c = tid; c < CL_LD; c += CH_THREADS) s_rows[rr * CL_LD + c] = (rr < n_sm && c < m) ? __ldg(a + c) : 0.0;
    }
    double rrow[CL_REGROWS][4];
#pragma unroll
    for (int e = 0; e < CL_REGROWS; ++e)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = tid + CH_THREADS * u;
            rrow[e][u] = (n_sm + e < n_own && c < m) ? __ldg(p.Ms + (size_t)(row0 + n_sm + e) * m + c) : 0.0;
        }
    double qv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int c = tid + CH_THREADS * u; qv[u] = c < m ? p.V[(size_t)m + c] : 0.0; }
    for (int c = tid; c < CL_LD; c += CH_THREADS) s_vec[c] = 0.0;            // columns m .. CL_LD - 1 stay zero for the whole solve
    for (int i = tid; i < CL_MAXROWS * cap; i += CH_THREADS) Vs[i] = 0.0;
    __syncthreads();
    if (tid < 2 * n_own) { const int r = tid >> 1, i = tid & 1; Vs[(size_t)r * cap + i] = p.V[(size_t)i * m + row0 + r]; }
    if (tid < W) dsm_arrive(&s_bar[3], tid);                                 // this CTA's vector buffer may be written
    if (tid == CH_THREADS - 1) mbar_expect(&s_bar[2], 8u * (unsigned int)m);
    // first candidate: w_1 = Ms q_1 (sigma = 0)
    merged_symv(s_rows, rrow, qv, s_red, tid, warp, lane);
    __syncthreads();
    merged_symv_totals(s_red, s_y, n_own, n_sm, tid);
    if (b == 0 && tid == CH_THREADS - 1) s_sc[1] = ld_word(&ctl->stop) != 0u ? 1.0 : 0.0;
    __syncthreads();
    if (tid < n_own) s_nw[tid] = s_y[tid];
    double sigma = 0.0, alpha1 = 0.0;
    double my_alpha = 0.0, my_beta = 0.0, my_beta_prev = 0.0;                // alpha_i, beta_i, beta_{i-1} of step i = tid
    unsigned int phA = 0u, phH = 0u, phQ = 0u, phF = 0u;
    __syncthreads();
    if (p.trace && b == 0 && tid == 0) p.trace[150 * 16 + 1] = wall_ns();

    int j = 1;                                                               // the candidate w_j is exchanged; step j completes with its coefficients
    int end_reason = 0;                                                      // 1 stop word, 2 breakdown, 3 max_steps, 4 abort
    for (; j <= p.max_steps; ++j) {
        MG_TRACE(0);
        double alpha_j = sigma, beta_j = 0.0, inv_beta = 0.0, hj = 0.0, h_keep = 0.0, ww0 = 0.0;
        bool leave = false;
        for (int pass = 0; pass < 2; ++pass) {
            const unsigned int tag = 2u * (unsigned int)j + (unsigned int)pass;
            const int n_tot = j + 3;                                         // slot 0: |w|^2, slot 1 + i: V_i . w (i = 0..j), slot j + 2: stop word
            const int n_mine = rank < n_tot ? (n_tot - rank + W - 1) / W : 0;   // slots rank, rank + W, ... are added up by this CTA
            // A cluster may run one exchange ahead of a slow reader in another cluster (its next packets depend on the vector and on
            // its OWN gathers only): packets alternate between buffers by step parity (and pass), the vector packets by step parity.
            uint4* cb = p.cbuf + (size_t)(2 * (j & 1) + pass) * p.cb_stride;
            uint4* qb = p.qbuf + (size_t)(j & 1) * m;
            if (tid == CH_THREADS - 1) { mbar_expect(&s_bar[0], 8u * (unsigned int)(n_mine * W)); mbar_expect(&s_bar[1], 8u * (unsigned int)n_tot); }
            const double* cand = pass == 0 ? s_nw : s_upd;                   // second pass: coefficients of w_j - V h
            // ---- A1: partial coefficients of the owned rows -> the owner of the slot in this cluster ----
            if (tid < n_tot) {
                const int s = tid;
                double v = 0.0;
                if (s == n_tot - 1) v = (b == 0) ? s_sc[1] : 0.0;
                else {                                                       // rows beyond n_own hold zeros
                    const double* vcol = s == 0 ? cand : Vs + (s - 1);
                    const int vstep = s == 0 ? 1 : cap;
                    double vv[6];                                            // six chains of three (a dependent FMA costs ~50 cycles)
#pragma unroll
                    for (int q = 0; q < 6; ++q) vv[q] = vcol[q * vstep] * cand[q];
#pragma unroll
                    for (int r = 6; r < CL_MAXROWS; ++r) vv[r % 6] = fma(vcol[r * vstep], cand[r], vv[r % 6]);
                    v = ((vv[0] + vv[1]) + (vv[2] + vv[3])) + (vv[4] + vv[5]);
                }
                const int idx = s / W, owner = s - idx * W;
                if (pass == 0) MG_TRACE(12);
                dsm_send(&s_part[idx * W + rank], owner, v, &s_bar[0]);
            }
            int bad = 0;
            // ---- the candidate itself: own rows -> the CTAs of this cluster and packets for the other clusters ----
            if (pass == 0) {
                if (!wait_mbar(&s_bar[3], phF, ctl)) bad = 1;                // every CTA of the cluster has copied the previous vector out
                phF ^= 1u;
                if (!bad) {
                    if (tid < n_own) st_ll(qb + row0 + tid, s_nw[tid], 2u * (unsigned int)j);
                    if (tid < n_own * W) {
                        const int r = tid / W, d = tid - r * W;
                        dsm_send(&s_vec[row0 + r], d, s_nw[r], &s_bar[2]);
                    }
                }
                MG_TRACE(1);
            }
            // ---- A2: cluster totals of the owned slots -> packets ----
            if (!wait_mbar(&s_bar[0], phA, ctl)) bad = 1;
            phA ^= 1u;
            if (pass == 0) MG_TRACE(2);
            if (tid < n_mine && !bad) {
                double t[8];                                                 // pairwise tree in rank order (W <= 16: two per leaf)
#pragma unroll
                for (int q = 0; q < 8; ++q) { t[q] = q < W ? s_part[tid * W + q] : 0.0; if (q + 8 < W) t[q] += s_part[tid * W + q + 8]; }
                const double tot = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
                st_ll(cb + (size_t)(tid * W + rank) * C + cid, tot, tag);
            }
            if (pass == 0) MG_TRACE(3);
            // ---- Q: this CTA's share of the other clusters' rows, forwarded to the CTAs of this cluster (threads t0 .. t0 + nt - 1) ----
            auto forward_rows = [&](int t0, int nt) {
                const unsigned int qtag = 2u * (unsigned int)j;
                for (int c = f_lo + (tid - t0); c < f_hi; c += nt) {
                    if (c >= crow0 && c < crow1) continue;
                    double x;
                    if (!wait_ll_short(qb + c, qtag, x, ctl)) { bad = 1; continue; }
                    for (int d = 0; d < W; ++d) dsm_send(&s_vec[c], d, x, &s_bar[2]);
                }
            };
            // ---- A3: totals over all clusters (G lanes per slot, lane c reads cluster c), broadcast inside the cluster ----
            // (threads 0 .. gt - 1 take part, gt a multiple of 32)
            auto gather_coefficients = [&](int gt) {
                const int per_pass = gt / G;
                const int c = tid & (G - 1);
                for (int i0 = 0; i0 < n_mine; i0 += per_pass * CL_GPK) {
                    double x[CL_GPK];
                    unsigned int pend = 0u;
#pragma unroll
                    for (int q = 0; q < CL_GPK; ++q) { x[q] = 0.0; if (i0 + q * per_pass + tid / G < n_mine && c < C) pend |= 1u << q; }
                    const uint4* src = cb + (size_t)((i0 + tid / G) * W + rank) * C + c;
                    const size_t src_step = (size_t)per_pass * W * C;
                    const long long t0 = now();
                    for (unsigned int rounds = 1; pend != 0u; ++rounds) {
                        const unsigned int cur = pend;
#pragma unroll
                        for (int q = 0; q < CL_GPK; ++q)
                            if ((cur >> q) & 1u) { double v; if (ld_ll(src + q * src_step, tag, v)) { x[q] = v; pend &= ~(1u << q); } }
                        if (pend != 0u) {
                            relax_short();
                            if ((rounds & 63u) == 0u && (ld_word(&ctl->abort) != 0u || now() - t0 > CH_TIMEOUT)) { st_word(&ctl->abort, 1u); bad = 1; break; }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < CL_GPK; ++q) {
                        if (i0 + q * per_pass >= n_mine) break;              // uniform in the CTA
                        double v = x[q];
                        for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        const int idx = i0 + q * per_pass + tid / G;
                        if (idx < n_mine)
                            for (int d = c; d < W; d += G) dsm_send(&s_h[idx * W + rank], d, v, &s_bar[1]);
                    }
                }
            };
            // ---- the gathered candidate: copy out, release the buffer, y = Ms w_j on the owned rows ----
            // order 1 (default): forward the rows, the product, then the gather (it finds every packet in place): 1.51 ms at C2.
            // order 2 (ESPER_MERGED_ORDER=2): the last four warps forward the rows WHILE the others gather the coefficients, then the
            // product -- measured slower (1.70 ms): whenever the coefficient broadcast lands in shared memory during the product, the
            // product takes 2.6 instead of 1.9 us (the same was seen with the gather in front of the product, 1.58 ms).  Loading the
            // packets before the product and decoding them after it: the four live registers spill, 1.48 ms -- no gain either.
            bool vec_ok = true;
            if (pass == 0) {
                constexpr int FW = 128;
                if (p.mg_order == 2) {
                    if (!bad) { if (tid >= CH_THREADS - FW) forward_rows(CH_THREADS - FW, FW); else gather_coefficients(CH_THREADS - FW); }
                    if (tid == CH_THREADS - FW) MG_TRACE_ANY(4);
                    MG_TRACE(5);
                } else {
                    if (!bad) forward_rows(0, CH_THREADS);
                    MG_TRACE(4);
                }
                if (!bad && !wait_mbar(&s_bar[2], phQ, ctl)) bad = 1;
                phQ ^= 1u;
                MG_TRACE(6);
#pragma unroll
                for (int u = 0; u < 4; ++u) qv[u] = s_vec[tid + CH_THREADS * u];
                if (tid < n_own) s_w[tid] = s_vec[row0 + tid];
                vec_ok = __syncthreads_or(bad) == 0;
                if (vec_ok) {
                    if (tid < W) dsm_arrive(&s_bar[3], tid);
                    if (tid == CH_THREADS - 1) mbar_expect(&s_bar[2], 8u * (unsigned int)m);
                    merged_symv(s_rows, rrow, qv, s_red, tid, warp, lane);
                    __syncthreads();
                    merged_symv_totals(s_red, s_y, n_own, n_sm, tid);
                    MG_TRACE(7);
                    if (p.mg_order != 2) { gather_coefficients(CH_THREADS); MG_TRACE(5); }
                }
            } else gather_coefficients(CH_THREADS);
            if (!vec_ok) { end_reason = 4; leave = true; break; }
            // ---- the coefficients ----
            if (!bad && !wait_mbar(&s_bar[1], phH, ctl)) bad = 1;
            phH ^= 1u;
            if (pass == 0) MG_TRACE(8);
            if (__syncthreads_or(bad)) { end_reason = 4; leave = true; break; }          // also: s_y of all rows is visible
            if (s_h[n_tot - 1] != 0.0) { end_reason = 1; leave = true; break; }
            double ww = s_h[0], sum_h2 = 0.0;
            if (pass == 0) ww0 = ww;
            if (pass == 0) { if (tid <= j) h_keep = s_h[1 + tid]; }
            else {
                // second pass: |h2|^2 first, then H = h + h2 in place
                double a0 = 0.0;
                if (warp == 0) { for (int i = lane; i <= j; i += 32) a0 = fma(s_h[1 + i], s_h[1 + i], a0); a0 = wsum(a0); if (lane == 0) s_sc[2] = a0; }
                __syncthreads();
                sum_h2 = s_sc[2];
                if (tid <= j) s_h[1 + tid] += h_keep;
                __syncthreads();
            }
            // g = T~ H + H_j H (thread i: coefficient of V_i)
            hj = s_h[1 + j];
            if (tid <= j) {
                const int i = tid;
                const double hi = s_h[1 + i];
                // (diagonal + H_j) H_i and the two neighbour terms are independent: depth 2 instead of 4
                const double diag = (i == 0 ? p.lambda0 : (i < j ? my_alpha : sigma)) + hj;
                const double lo_t = (i >= 2) ? my_beta_prev * s_h[i] : 0.0;             // beta_{i-1} H_{i-1}
                const double up_t = (i >= 1 && i + 1 <= j) ? my_beta * s_h[2 + i] : 0.0;   // beta_i H_{i+1}
                s_g[1 + i] = fma(diag, hi, lo_t + up_t);
            }
            __syncthreads();
            if (pass == 0) MG_TRACE(9);
            // update of the owned rows (half-warp r): w_j - V H and y - V g; sum of squares of the coefficients (half-warp 31)
            {
                const int task = tid >> 4, l16 = tid & 15;
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                if (task < n_own) {
                    const double* vr = Vs + (size_t)task * cap;
                    int i = l16;
                    for (; i + 16 <= j; i += 32) {
                        const double x0 = vr[i], x1 = vr[i + 16];
                        a0 = fma(x0, s_h[1 + i], a0); a1 = fma(x1, s_h[1 + i + 16], a1);
                        b0 = fma(x0, s_g[1 + i], b0); b1 = fma(x1, s_g[1 + i + 16], b1);
                    }
                    if (i <= j) { const double x0 = vr[i]; a0 = fma(x0, s_h[1 + i], a0); b0 = fma(x0, s_g[1 + i], b0); }
                } else if (task == 31) {
                    int i = l16;
                    for (; i + 16 <= j; i += 32) { a0 = fma(s_h[1 + i], s_h[1 + i], a0); a1 = fma(s_h[1 + i + 16], s_h[1 + i + 16], a1); }
                    if (i <= j) a0 = fma(s_h[1 + i], s_h[1 + i], a0);
                }
                double ta = a0 + a1, tb = b0 + b1;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) { ta += __shfl_xor_sync(0xffffffffu, ta, o); tb += __shfl_xor_sync(0xffffffffu, tb, o); }
                if (l16 == 0) {
                    if (task < n_own) { s_upd[task] = s_w[task] - ta; s_z[task] = s_y[task] - tb; }
                    else if (task == 31) s_sc[0] = ta;
                }
            }
            __syncthreads();
            if (pass == 0) MG_TRACE(10);
            if (pass == 0) sum_h2 = s_sc[0];
            const double after = ww - sum_h2;
            alpha_j = sigma + hj;
            if (pass == 0 && after < 0.5 * ww) continue;                     // DGKS: orthogonalise once more (s_upd is the candidate)
            inv_beta = after > 0.0 ? rsqrt(after) : 0.0;                    // one reciprocal square root instead of sqrt + divide on the critical path
            beta_j = after * inv_beta;
            break;
        }
        if (leave) break;
        if (j == 1) alpha1 = alpha_j;
        // ---- q_{j+1}, Ms q_{j+1} and the next candidate on the owned rows ----
        const bool breakdown = !(beta_j > 1e-14 * fmax(fabs(alpha1), 1e-300));
        const double inv = inv_beta;
        if (tid < n_own) {
            const double q = s_upd[tid] * inv;
            const double z = fma(-hj, q, s_z[tid] * inv);
            Vs[(size_t)tid * cap + j + 1] = q;
            p.V[(size_t)(j + 1) * m + row0 + tid] = q;
            s_nw[tid] = fma(-beta_j, Vs[(size_t)tid * cap + j], fma(-alpha_j, q, z));
        }
        if (tid == j) { my_alpha = alpha_j; my_beta = beta_j; }
        if (tid == j + 1) my_beta_prev = beta_j;
        if (b == 0 && tid == 32) {
            p.alpha[j - 1] = alpha_j; p.beta[j - 1] = beta_j;
            st_ll(p.abuf + 2 * (size_t)j, alpha_j, 1u);
            st_ll(p.abuf + 2 * (size_t)j + 1, beta_j, 1u);
            st_word(&ctl->progress, (unsigned int)j);                        // a hint: the checkers wait for the packets themselves
        }
        if (b == 0 && tid == CH_THREADS - 1) s_sc[1] = ld_word(&ctl->stop) != 0u ? 1.0 : 0.0;
        MG_TRACE(11);
        if (breakdown) { end_reason = 2; ++j; break; }
        if (j == p.max_steps) { end_reason = 3; ++j; break; }
        // Ms q_{j+1} came from a recurrence that divides by beta_j: when the candidate lost eight digits in the orthogonalisation
        // (numerically exhausted Krylov space that the breakdown threshold has not caught yet: duplicated images, rank-deficient Ms)
        // its error is no longer at rounding level and would be amplified step after step.  T_j is still good: stop here like at
        // max_steps -- checker 0 tests T_j, and the host falls back to the barrier kernels if that is not enough.
        if (!(beta_j * beta_j > 1e-16 * fmax(ww0, alpha1 * alpha1))) { end_reason = 3; ++j; break; }
        sigma = alpha_j;
        __syncthreads();
    }
    if (p.trace && b == 0 && tid == 0) p.trace[150 * 16 + 2] = wall_ns();
    if (b == 0 && tid == 0) {
        ctl->steps_run = j - 1;
        if (end_reason == 2 || end_reason == 3) {
            publish_fence();
            st_word(&ctl->final_n, 0x80000000u | (end_reason == 2 ? 0x40000000u : 0u) | (unsigned int)(j - 1));
        }
    }
    cluster_sync_all();                                                      // no CTA leaves while a peer may still store into it
#undef MG_TRACE
#undef MG_TRACE_ANY
}

// ---- end chain kernel text ----

}  // namespace chain

// Ritz vectors U = [V_0 | (V_1 .. V_n) S] with the step count and S of the winning checker read on the device; unit columns, largest-magnitude
// entry positive.  Two launches that use the whole device (a single launch with one CTA per column took 0.095 ms at C2: 10 CTAs):
// ritz_rows_kernel -- a CTA forms ALL columns for 32 rows (lane = row, warp = every 8th basis vector, partial sums added over the warps in
// a fixed order), stores them unscaled and leaves per-CTA column statistics (sum of squares, largest magnitude and its row);
// ritz_scale_kernel -- CTA i reduces the statistics of column i in CTA order and scales / orients the column.
constexpr int RZ_ROWS = 32, RZ_SLICES = 8;
__global__ void __launch_bounds__(RZ_ROWS * RZ_SLICES) ritz_rows_kernel(const double* __restrict__ V, int m, const chain::Ctl* __restrict__ ctl,
                                                                          const double* __restrict__ S_base, int s_stride, int kw,
                                                                          double* __restrict__ U, int k, double* __restrict__ part) {
    __shared__ double s_acc[RZ_SLICES][RZ_ROWS + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.x * RZ_ROWS + lane;
    const int steps = ctl->n_final, win = ctl->winner;
    const bool dead = (win < 0 || steps < 1);
    const double* S = S_base + (size_t)max(win, 0) * s_stride;
    double acc[chain::CH_MAXWANT];
#pragma unroll
    for (int i = 0; i < chain::CH_MAXWANT; ++i) acc[i] = 0.0;
    if (!dead && r < m) {
        const double* v = V + (size_t)m + r;
        for (int l = w; l < steps; l += RZ_SLICES) {
            const double x = v[(size_t)l * m];
            const double* sl = S + (size_t)l * kw;
#pragma unroll
            for (int i = 0; i < chain::CH_MAXWANT; ++i) if (i < kw) acc[i] = fma(x, __ldg(sl + i), acc[i]);
        }
    }
    double* pb = part + (size_t)blockIdx.x * k * 3;
    // column i of U (i = 0 the locked vector, else Ritz vector i - 1): store, statistics of this CTA's 32 rows
    auto emit = [&](int i, double x) {
        if (r < m) U[(size_t)r * k + i] = x;
        double sq = x * x, best = r < m ? fabs(x) : -1.0; int best_r = r < m ? r : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int orr = __shfl_xor_sync(0xffffffffu, best_r, o);
            if (ob > best || (ob == best && orr < best_r)) { best = ob; best_r = orr; }
        }
        if (lane == 0) { pb[i * 3] = sq; pb[i * 3 + 1] = best; pb[i * 3 + 2] = (double)best_r; }
    };
    if (w == 0) emit(0, r < m ? V[r] : 0.0);
#pragma unroll
    for (int q = 0; q < chain::CH_MAXWANT; ++q) {
        if (q + 1 < k) {                                                     // uniform
            s_acc[w][lane] = acc[q];
            __syncthreads();
            if (w == 0) {
                double x = 0.0;
#pragma unroll
                for (int t = 0; t < RZ_SLICES; ++t) x += s_acc[t][lane];
                emit(q + 1, x);
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) ritz_scale_kernel(double* __restrict__ U, int m, int k, const double* __restrict__ part, int nblk) {
    __shared__ double s_sq[8], s_bb[8], s_scale;
    __shared__ int s_br[8];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // statistics of column i over the CTAs of ritz_rows_kernel: thread t takes CTAs t, t + 256, ... (one load each at the usual sizes),
    // then a fixed shuffle tree and the eight warp results in warp order -- the same bits in every run
    double sq = 0.0, bb = -1.0; int br = 0x7fffffff;
    for (int b = tid; b < nblk; b += 256) {
        const double* pb = part + ((size_t)b * k + i) * 3;
        sq += pb[0];
        const int rr = (int)pb[2];
        if (pb[1] > bb || (pb[1] == bb && rr < br)) { bb = pb[1]; br = rr; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const double ob = __shfl_xor_sync(0xffffffffu, bb, o);
        const int orr = __shfl_xor_sync(0xffffffffu, br, o);
        if (ob > bb || (ob == bb && orr < br)) { bb = ob; br = orr; }
    }
    if (lane == 0) { s_sq[warp] = sq; s_bb[warp] = bb; s_br[warp] = br; }
    __syncthreads();
    if (tid == 0) {
        sq = 0.0; bb = -1.0; br = 0x7fffffff;
        for (int w = 0; w < 8; ++w) {
            sq += s_sq[w];
            if (s_bb[w] > bb || (s_bb[w] == bb && s_br[w] < br)) { bb = s_bb[w]; br = s_br[w]; }
        }
        const double nrm = sqrt(sq);
        double sc = nrm > 0.0 ? 1.0 / nrm : 0.0;
        if (br < m && U[(size_t)br * k + i] < 0.0) sc = -sc;
        s_scale = sc;
    }
    __syncthreads();
    const double sc = s_scale;
    for (int r = tid; r < m; r += 256) U[(size_t)r * k + i] *= sc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side.  Preconditions (checked by chain_plan): 32 <= m <= 2048, ceil(m / (SMs - 1)) <= 14, kw + guard <= 16.
// V[0] (locked vector) and V[1] (start vector) are already on the device.  On success U_d [m, k] holds the Ritz vectors (column 0 =
// V[0]; unit columns, largest-magnitude entry positive) and the result block is on its way to the host; the caller synchronises.
// Returns ESPER_OK, an error, or +1 = "not handled / not converged here: use the barrier kernels".
// ---------------------------------------------------------------------------------------------------------------------
struct ChainPlan {
    bool ok = false;
    int chunk = 0, n_workers = 0, n_checkers = 0, n_sm = 0, vs_cap = 0, max_steps = 0;
    size_t smem = 0;
};

static ChainPlan chain_plan(esper_ctx* c, int m, int kw, int guard, int max_iter, long long smem_limit) {
    using namespace chain;
    ChainPlan pl;
    if (m < 32 || m > CH_MAXM || kw < 1 || kw + guard > CH_MAXWANT || c->sm_count < 2 || c->sm_count > 32 * CH_PK) return pl;
    pl.chunk = ceil_div(m, c->sm_count - 1);
    if (pl.chunk > CH_MAXROWS) return pl;
    pl.n_workers = ceil_div(m, pl.chunk);
    pl.n_checkers = std::min(c->sm_count - pl.n_workers, CH_MAXCHK);
    if (pl.n_checkers < 1) return pl;
    pl.n_sm = std::max(0, pl.chunk - CH_REGROWS);
    const size_t fixed = sizeof(double) * ((size_t)m + (size_t)pl.n_sm * m);
    if ((long long)fixed + 4096 > smem_limit) return pl;
    // workers: Vs [chunk][vs_cap]; checkers: 3 cap + 3*16 + 18 + 4 want cap doubles + counters + swap flags
    const int want = kw + guard;
    int steps = std::min(max_iter, CH_THREADS - 2);                            // all coefficients of a step: one packet per thread
    while (steps >= 32 && (fixed + sizeof(double) * (size_t)pl.chunk * (steps + 2) > (size_t)smem_limit ||
                           sizeof(double) * checker_doubles(want, steps + 2) > (size_t)smem_limit)) steps -= 2;
    if (steps < std::min(max_iter, 96)) return pl;
    pl.max_steps = steps;
    pl.vs_cap = steps + 2;
    pl.smem = std::max(fixed + sizeof(double) * (size_t)pl.chunk * pl.vs_cap, sizeof(double) * checker_doubles(want, steps + 2));
    pl.ok = (long long)pl.smem <= smem_limit;
    return pl;
}

// Plan of the cluster schedule: clusters of S CTAs, as many as the device keeps resident at once (cooperative launch: the CTAs
// spin on each other), CL_MAXROWS rows per worker at most, at least one CTA left over for a checker.
struct ClusterPlan {
    bool ok = false;
    int S = 0, total = 0, chunk = 0, n_workers = 0, n_checkers = 0, cl_count = 0, cl_group = 1, n_sm = 0, vs_cap = 0, max_steps = 0;
    size_t smem = 0;
};

static ClusterPlan cluster_plan(int S, int n_clusters, int m, int kw, int guard, int max_iter, long long smem_limit, bool merged = false) {
    using namespace chain;
    ClusterPlan pl;
    auto cluster_worker_doubles = [merged](int chunk, int cap) { return merged ? merged_worker_doubles(cap) : chain::cluster_worker_doubles(chunk, cap); };
    pl.S = S; pl.total = S * n_clusters;
    if (m < 32 || m > CH_MAXM || kw < 1 || kw + guard > CH_MAXWANT || pl.total < 2) return pl;
    pl.chunk = ceil_div(m, pl.total - 1);
    if (pl.chunk > CL_MAXROWS) return pl;
    pl.n_workers = ceil_div(m, pl.chunk);
    pl.n_checkers = std::min(pl.total - pl.n_workers, CH_MAXCHK);
    if (pl.n_checkers < 1) return pl;
    pl.cl_count = ceil_div(pl.n_workers, S);
    if (pl.cl_count > CL_MAXCLUSTERS) return pl;
    while (pl.cl_group < pl.cl_count) pl.cl_group *= 2;
    pl.n_sm = std::max(0, pl.chunk - CL_REGROWS);
    const int want = kw + guard;
    int steps = std::min(max_iter, CH_THREADS - 3);                            // j + 3 slots of a step: one thread each
    while (steps >= 32 && (sizeof(double) * cluster_worker_doubles(pl.chunk, steps + 2) > (size_t)smem_limit ||
                           sizeof(double) * checker_doubles(want, steps + 2) > (size_t)smem_limit)) steps -= 2;
    if (steps < std::min(max_iter, 96)) return pl;
    pl.max_steps = steps;
    pl.vs_cap = steps + 2;
    pl.smem = std::max(sizeof(double) * cluster_worker_doubles(pl.chunk, pl.vs_cap), sizeof(double) * checker_doubles(want, steps + 2));
    pl.ok = (long long)pl.smem <= smem_limit;
    return pl;
}

// Clusters of the cluster kernel the device holds at once (0: clusters of this size cannot be launched); cached per ctx.
static int cluster_capacity(esper_ctx* c, int S, bool merged = false) {
    using namespace chain;
    int& cap_size = merged ? c->merged_cap_size : c->cluster_cap_size;
    int& cap = merged ? c->merged_cap : c->cluster_cap;
    long long& smem = merged ? c->merged_smem : c->cluster_smem;
    void (*kernel)(Params) = merged ? lanczos_merged_kernel : lanczos_cluster_kernel;
    if (cap_size == S) return cap;
    int optin = 0, n = 0;
    cudaFuncAttributes fa;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, kernel) != cudaSuccess ||
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes) != cudaSuccess) {
        cudaGetLastError();
        cap_size = S; cap = 0;
        return 0;
    }
    smem = optin - (long long)fa.sharedSizeBytes;
    if (S > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S * 64); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, (void*)kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    cap_size = S; cap = n;
    return n;
}

int run_lanczos_chain(esper_ctx* c, int m, int k, int guard, double tol, int max_iter, double* V, double* U_d, bool unit_locked) {
    using namespace chain;
    const int kw = k - 1;
    cudaStream_t st = c->stream;
    if (!c->chain_smem_set) {
        int optin = 0;
        cudaFuncAttributes fa;
        ESPER_CUDA(c, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        ESPER_CUDA(c, cudaFuncGetAttributes(&fa, lanczos_chain_kernel));
        ESPER_CUDA(c, cudaFuncSetAttribute(lanczos_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        c->chain_smem_set = optin - (long long)fa.sharedSizeBytes;
    }
    // cluster schedules first (ESPER_CLUSTER_SIZE = 0 switches them off, 2 / 4 / 8 / 16 choose the cluster size; default 8).  The merged
    // schedule (one exchange per step; ESPER_CHAIN_MERGED = 0 switches it off) needs the locked pair to be an eigenpair with a known
    // eigenvalue (unit_locked: the Markov matrix, lambda0 = 1); its basis capacity is a little smaller (one more coefficient array).
    ClusterPlan cp;
    bool merged = false;
    {
        const char* e = getenv("ESPER_CLUSTER_SIZE");
        const int S = e ? atoi(e) : 8;
        const char* em = getenv("ESPER_CHAIN_MERGED");
        const bool want_merged = unit_locked && !(em && atoi(em) == 0);
        if (S == 2 || S == 4 || S == 8 || S == 16) {
            if (want_merged) {
                const int cap = cluster_capacity(c, S, true);
                if (cap > 0) cp = cluster_plan(S, cap, m, kw, guard, max_iter, c->merged_smem, true);
                merged = cp.ok;
            }
            if (!cp.ok) {
                const int cap = cluster_capacity(c, S);
                if (cap > 0) cp = cluster_plan(S, cap, m, kw, guard, max_iter, c->cluster_smem);
            }
            if (getenv("ESPER_DEBUG_LANCZOS"))
                fprintf(stderr, "[chain] clusters of %d (%s schedule): %d resident -> plan %s: %d rows per CTA, %d workers in %d clusters, %d checkers, %d steps, %zu B of shared memory\n", S,
                        merged ? "merged" : "two-exchange", merged ? c->merged_cap : c->cluster_cap, cp.ok ? "ok" : "rejected", cp.chunk, cp.n_workers, cp.cl_count, cp.n_checkers, cp.max_steps, cp.smem);
        }
    }
    const ChainPlan pl = chain_plan(c, m, kw, guard, max_iter, c->chain_smem_set);
    if (!cp.ok && !pl.ok) return 1;
    const bool cluster = cp.ok;
    if (!cluster) {
        int per_sm = 0;
        ESPER_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_chain_kernel, CH_THREADS, pl.smem));
        if (per_sm < 1) return 1;
    }
    const int max_steps = cluster ? cp.max_steps : pl.max_steps;
    const int n_workers = cluster ? cp.n_workers : pl.n_workers;

    // workspace: packets | ctl | theta | S | alpha, beta
    const int slots = max_steps + 4;
    const size_t cb_stride = (size_t)slots * (size_t)std::max(cp.cl_count, 1);
    const size_t n_p = merged ? 4 * cb_stride : cluster ? cb_stride : (size_t)slots * n_workers;      // cbuf (4 buffers: merged schedule) or pbuf
    const size_t n_h = (size_t)slots + 1, n_q = (merged ? 2 : 1) * (size_t)m, n_a = 2 * ((size_t)max_steps + 2);
    const size_t pk_bytes = sizeof(uint4) * (n_p + n_h + n_q + n_a);
    const int s_stride = max_steps * kw;
    const size_t n_ritz_part = (size_t)ceil_div(m, RZ_ROWS) * k * 3;
    const size_t bytes = pk_bytes + sizeof(Ctl) + 64 + sizeof(double) * ((size_t)CH_MAXCHK * CH_MAXWANT + (size_t)CH_MAXCHK * s_stride + 2 * ((size_t)max_steps + 8) + n_ritz_part);
    ESPER_RESERVE(c, c->chain_ws, bytes);
    char* base = c->chain_ws.as<char>();
    uint4* pbuf = reinterpret_cast<uint4*>(base);
    uint4* hbuf = pbuf + n_p; uint4* qbuf = hbuf + n_h; uint4* abuf = qbuf + n_q;
    Ctl* ctl = reinterpret_cast<Ctl*>(base + pk_bytes);
    double* theta_d = reinterpret_cast<double*>(base + pk_bytes + ((sizeof(Ctl) + 63) / 64) * 64);
    double* S_d = theta_d + (size_t)CH_MAXCHK * CH_MAXWANT;
    double* alpha_d = S_d + (size_t)CH_MAXCHK * s_stride;
    double* beta_d = alpha_d + max_steps + 4;
    double* ritz_part_d = beta_d + max_steps + 4;
    // tags repeat from one solve to the next: the packet buffers and the control block start from zero every time
    ESPER_CUDA(c, cudaMemsetAsync(base, 0, pk_bytes + sizeof(Ctl), st));

    Params p;
    p.Ms = c->markov.as<double>(); p.m = m; p.V = V; p.alpha = alpha_d; p.beta = beta_d;
    p.chunk = cluster ? cp.chunk : pl.chunk; p.n_workers = n_workers; p.n_checkers = cluster ? cp.n_checkers : pl.n_checkers;
    p.max_steps = max_steps; p.vs_cap = cluster ? cp.vs_cap : pl.vs_cap;
    p.kw = kw; p.guard = guard; p.cap_steps = m - 1; p.tol = tol;
    p.first_check = std::min(max_steps, std::max(3 * (kw + guard) + 8, 32));
    p.check_back = 8;                        // growth over 8 steps: the gate opens when the eigenvector test is likely to pass at once
    p.check_stride = (p.chunk >= 12 ? 4 : 6) * (p.n_checkers >= 2 ? 1 : 2);       // an eigenvalue test takes about the time of 4 steps (6 short ones)
    p.pbuf = pbuf; p.hbuf = hbuf; p.qbuf = qbuf; p.abuf = abuf; p.ctl = ctl;
    p.theta_out = theta_d; p.S_out = S_d; p.s_stride = s_stride;
    p.prof_on = getenv("ESPER_DEBUG_LANCZOS") ? 1 : 0;
    p.cl_size = cluster ? cp.S : 0; p.cl_count = cp.cl_count; p.cl_group = cp.cl_group; p.cbuf = pbuf;
    p.merged = merged ? 1 : 0; p.cb_stride = cb_stride; p.lambda0 = 1.0;
    p.mg_order = getenv("ESPER_MERGED_ORDER") ? atoi(getenv("ESPER_MERGED_ORDER")) : 1;
    p.trace = nullptr; p.trace_step = 0;
    if (cluster && getenv("ESPER_CLUSTER_RELAX_NS")) { const unsigned int ns = (unsigned int)atoi(getenv("ESPER_CLUSTER_RELAX_NS")); cudaMemcpyToSymbolAsync(g_relax_ns, &ns, sizeof ns, 0, cudaMemcpyHostToDevice, st); }
    if (cluster && p.prof_on) {
        ESPER_RESERVE(c, c->chain_trace, sizeof(long long) * 16 * 160);
        ESPER_CUDA(c, cudaMemsetAsync(c->chain_trace.p, 0, sizeof(long long) * 16 * 160, st));
        p.trace = c->chain_trace.as<long long>();
        p.trace_step = getenv("ESPER_TRACE_STEP") ? atoi(getenv("ESPER_TRACE_STEP")) : 100;
        c->chain_trace_ctas = cp.n_workers;
    }
    if (cluster) {
        LaunchScope ls(c, "lanczos");
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cp.total); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = cp.smem; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cp.S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative;
        at[1].val.cooperative = 1;
        // ESPER_CHAIN_NOCOOP=1 (profiling only): without the cooperative attribute -- Nsight Compute 2025.2 fails to launch a kernel that is
        // both cooperative and clustered; on an otherwise idle device the CTAs are co-resident anyway (the plan never exceeds the occupancy query)
        cfg.attrs = at; cfg.numAttrs = getenv("ESPER_CHAIN_NOCOOP") ? 1 : 2;
        cudaError_t e = merged ? cudaLaunchKernelEx(&cfg, lanczos_merged_kernel, p) : cudaLaunchKernelEx(&cfg, lanczos_cluster_kernel, p);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(c, ESPER_E_CUDA, "cooperative cluster launch (%d x %d): %s", cp.S, cp.total / cp.S, cudaGetErrorString(e)); }
        c->cluster_launches++;
        if (merged) c->merged_launches++;
        c->chain_trace_merged = merged;
    } else {
        LaunchScope ls(c, "lanczos");
        void* args[] = {&p};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)lanczos_chain_kernel, dim3(c->sm_count), dim3(CH_THREADS), args, pl.smem, st);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(c, ESPER_E_CUDA, "cooperative launch (chain): %s", cudaGetErrorString(e)); }
    }
    {
        const int nblk = ceil_div(m, RZ_ROWS);
        { LaunchScope ls(c, "lanczos"); ritz_rows_kernel<<<nblk, RZ_ROWS * RZ_SLICES, 0, st>>>(V, m, ctl, S_d, s_stride, kw, U_d, k, ritz_part_d); }
        { LaunchScope ls(c, "lanczos"); ritz_scale_kernel<<<k, 256, 0, st>>>(U_d, m, k, ritz_part_d, nblk); }
    }
    ESPER_CUDA(c, cudaGetLastError());
    // result block -> pinned host memory (same stream; the caller's final synchronise covers it)
    if (c->pin_chain.reserve(sizeof(Ctl) + sizeof(double) * CH_MAXCHK * CH_MAXWANT) != 0) return fail(c, ESPER_E_NOMEM, "pinned alloc failed");
    ESPER_CUDA(c, cudaMemcpyAsync(c->pin_chain.p, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    ESPER_CUDA(c, cudaMemcpyAsync(c->pin_chain.as<char>() + sizeof(Ctl), theta_d, sizeof(double) * CH_MAXCHK * CH_MAXWANT, cudaMemcpyDeviceToHost, st));
    return ESPER_OK;
}

// After the stream has been synchronised: decode the result block.  Returns ESPER_OK, +1 (fall back) or an error.
int chain_result(esper_ctx* c, int k, double* theta_out, int* iters_out) {
    using namespace chain;
    const Ctl* h = c->pin_chain.as<Ctl>();
    const double* th = reinterpret_cast<const double*>(c->pin_chain.as<char>() + sizeof(Ctl));
    if (getenv("ESPER_DEBUG_LANCZOS")) {
        fprintf(stderr, "[chain] status=%d n_final=%d winner=%d steps_run=%d abort=%u stop=%08x\n", h->status, h->n_final, h->winner, h->steps_run, h->abort, h->stop);
        const double n = h->steps_run > 0 ? (double)h->steps_run : 1.0;
        fprintf(stderr, "[chain] CTA0 cycles/step: S0 %.0f | P1-P3 %.0f | norm+publish %.0f | P4 %.0f | second pass %.0f ; prologue %lld\n",
                h->prof[1] / n, h->prof[2] / n, h->prof[3] / n, h->prof[4] / n, h->prof[6] / n, h->prof[0]);
        if (h->prof[16]) fprintf(stderr, "[chain] cluster schedule, CTA0 cycles/step: partials in DSMEM %.0f | packets + gather %.0f | coefficients in DSMEM %.0f ; forward packets %.0f | vector in DSMEM %.0f\n",
                                 h->prof[16] / n, h->prof[17] / n, h->prof[2] / n, h->prof[18] / n, h->prof[4] / n);
        if (h->prof[16]) fprintf(stderr, "[chain] cluster schedule, thread 0 of CTA0: S0 = products %.0f | warp sums %.0f | barrier %.0f | row totals + barrier %.0f ; update: own row %.0f | barrier %.0f | rest %.0f\n",
                                 h->prof[19] / n, h->prof[20] / n, h->prof[21] / n, h->prof[1] / n, h->prof[22] / n, h->prof[23] / n, h->prof[3] / n);
        { double mw, mg; memcpy(&mw, &h->prof[24], 8); memcpy(&mg, &h->prof[25], 8); fprintf(stderr, "[chain] recurrence residuals |rho| / (|s| |T|): wanted pairs %.2e, guard pairs %.2e\n", mw, mg); }
        fprintf(stderr, "[chain] recurrence phases (cycles, summed over the vector tests of all checkers): reciprocals %lld | factors %lld | rows %lld | scaling %lld\n", h->prof[26], h->prof[27], h->prof[28], h->prof[29]);
        fprintf(stderr, "[chain] Ritz vectors of the winner: %s\n", h->prof[7] == 1 ? "backward recurrence" : h->prof[7] >= 1000 ? "inverse iteration (recurrence residual too large)" : "inverse iteration");
        for (int cc = 0; cc < CH_MAXCHK; ++cc) {
            const long long e = h->prof[8 + 2 * cc], v = h->prof[9 + 2 * cc];
            if (e >> 32) fprintf(stderr, "[chain] checker %d: %lld eigenvalue tests, %.0f cycles each; %lld vector tests, %.0f cycles each up to the decision\n", cc, e >> 32,
                                 (double)(e & 0xffffffffLL) / (double)(e >> 32), v >> 32, (v >> 32) ? (double)(v & 0xffffffffLL) / (double)(v >> 32) : 0.0);
        }
    }
    if (getenv("ESPER_DEBUG_LANCZOS") && c->chain_trace_ctas > 0 && c->chain_trace.p) {
        std::vector<long long> tr((size_t)16 * 160);
        cudaMemcpy(tr.data(), c->chain_trace.p, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost);
        long long t0 = -1;
        for (int b = 0; b < c->chain_trace_ctas; ++b) if (tr[(size_t)b * 16] && (t0 < 0 || tr[(size_t)b * 16] < t0)) t0 = tr[(size_t)b * 16];
        static const char* names2[13] = {"step start", "products done", "w ready", "partials sent", "partials complete", "packets published", "gather + broadcast sent",
                                         "coefficients complete", "update done", "norm known", "vector published", "forwarded", "vector complete"};
        static const char* names1[13] = {"step start", "partials + rows sent", "partials complete", "packets published", "forwarded", "gather + broadcast sent", "vector complete",
                                         "product done", "coefficients complete", "g done", "update done", "step end", "own partial computed"};
        const char** names = c->chain_trace_merged ? names1 : names2;
        fprintf(stderr, "[chain] timeline of one step over %d worker CTAs (thread 0 each), ns after the first CTA started it: min / median / max\n", c->chain_trace_ctas);
        for (int k = 0; k < 13; ++k) {
            std::vector<long long> v;
            for (int b = 0; b < c->chain_trace_ctas; ++b) if (tr[(size_t)b * 16 + k]) v.push_back(tr[(size_t)b * 16 + k] - t0);
            if (v.empty()) continue;
            std::sort(v.begin(), v.end());
            fprintf(stderr, "[chain]   %-24s %6lld %6lld %6lld\n", names[k], v.front(), v[v.size() / 2], v.back());
        }
        if (getenv("ESPER_TRACE_CLUSTERS")) {
            const int S = atoi(getenv("ESPER_TRACE_CLUSTERS"));
            fprintf(stderr, "[chain] the same per cluster of %d CTAs (median of its CTAs)\n", S);
            for (int k = 0; k < 13; ++k) {
                if (!names[k][0]) continue;
                fprintf(stderr, "[chain]   %-24s", names[k]);
                for (int b0 = 0; b0 < c->chain_trace_ctas; b0 += S) {
                    std::vector<long long> v;
                    for (int b = b0; b < std::min(b0 + S, c->chain_trace_ctas); ++b) if (tr[(size_t)b * 16 + k]) v.push_back(tr[(size_t)b * 16 + k] - t0);
                    std::sort(v.begin(), v.end());
                    fprintf(stderr, " %5lld", v.empty() ? -1LL : v[v.size() / 2]);
                }
                fprintf(stderr, "\n");
            }
        }
        const long long* e = tr.data() + 150 * 16;
        fprintf(stderr, "[chain] kernel timeline (ns after the start of CTA 0): prologue done %lld | workers leave %lld | checkers leave %lld %lld %lld %lld\n", e[1] - e[0], e[2] - e[0],
                e[4] ? e[4] - e[0] : 0, e[5] ? e[5] - e[0] : 0, e[6] ? e[6] - e[0] : 0, e[7] ? e[7] - e[0] : 0);
        c->chain_trace_ctas = 0;
    }
    if (h->abort != 0u) return fail(c, ESPER_E_CUDA, "lanczos (chain): a packet wait timed out (CTAs not co-resident?)");
    if (h->status != 1 || h->winner < 0) return 1;
    const int kw = k - 1;
    for (int i = 0; i < kw; ++i) theta_out[i] = th[(size_t)h->winner * CH_MAXWANT + i];
    if (iters_out) *iters_out = h->n_final;
    return ESPER_OK;
}

}  // namespace esper
