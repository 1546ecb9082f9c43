// chain_emu.cpp -- TEST INFRASTRUCTURE ONLY: the barrier-free Lanczos kernel of csrc/k4_chain.cu (cut into chain.inc by
// tests/test_emu.py) on the CPU.  ALL blocks run concurrently, one OS thread per CUDA thread; the packet primitives
// (st_ll / ld_ll: 16-byte {lo, tag, hi, tag} stores whose 8-byte halves are atomic) become two 64-bit atomics.
//   chain_emu Ms.bin m V0.bin kw guard tol grid max_steps out.bin
//   V0.bin: [2, m] = locked vector v0 and normalised start q1.
//   out.bin: ctl {status, n_final, winner, steps_run, abort} as 5 doubles | alpha[max_steps] | beta[max_steps] |
//            theta[CH_MAXWANT] of the winner | S[max_steps * kw] of the winner | V[(max_steps + 2), m]
#define EMU_ARENA_SHARED
#define EMU_NO_WARP_SUM
#define EMU_STATIC_BYTES_DEF (32 * 1024)
#define ESPER_CHAIN_EMU
#include "cuda_emu.h"
#include <algorithm>

using std::fmax; using std::fmin; using std::fabs; using std::sqrt; using std::frexp; using std::ldexp; using std::fma;

namespace esper {
namespace chain {
constexpr int CH_THREADS = 512;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr int CH_MAXROWS = 14;
constexpr int CH_REGROWS = 4;
constexpr int CH_MAXM = 2048;
constexpr int CH_MAXWANT = 24;
constexpr int CH_MAXCHK = 4;
constexpr int CH_POINTS = 16;
constexpr int CH_HWARP = 8;
constexpr int CH_PK = 8;
constexpr long long CH_TIMEOUT = 4000000000LL;
#include "chain_structs.inc"

inline void st_ll(uint4* p, double v, unsigned int tag) {
    uint64_t bits; std::memcpy(&bits, &v, 8);
    const uint64_t lo = (bits & 0xffffffffull) | ((uint64_t)tag << 32), hi = (bits >> 32) | ((uint64_t)tag << 32);
    uint64_t* q = reinterpret_cast<uint64_t*>(p);
    __atomic_store_n(q, lo, __ATOMIC_RELAXED);
    __atomic_store_n(q + 1, hi, __ATOMIC_RELAXED);
}
inline bool ld_ll(const uint4* p, unsigned int tag, double& v) {
    const uint64_t* q = reinterpret_cast<const uint64_t*>(p);
    const uint64_t lo = __atomic_load_n(q, __ATOMIC_RELAXED), hi = __atomic_load_n(q + 1, __ATOMIC_RELAXED);
    const uint64_t bits = (lo & 0xffffffffull) | (hi << 32);
    std::memcpy(&v, &bits, 8);
    const bool ok = (uint32_t)(lo >> 32) == tag && (uint32_t)(hi >> 32) == tag;
    if (!ok) std::this_thread::yield();
    return ok;
}
inline unsigned int ld_word(const unsigned int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_word(unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned int cas_word(unsigned int* p, unsigned int cmp, unsigned int v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline void publish_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long now() { return 0; }
inline void relax() { std::this_thread::yield(); }
#include "chain.inc"
}  // namespace chain
}  // namespace esper

template <class T> static std::vector<T> read_bin(const char* path, size_t n) {
    std::vector<T> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    using namespace esper::chain;
    if (argc < 10) return 2;
    const int m = atoi(argv[2]), kw = atoi(argv[4]), guard = atoi(argv[5]);
    const double tol = atof(argv[6]);
    const int grid = atoi(argv[7]), max_steps = atoi(argv[8]);
    std::vector<double> Ms = read_bin<double>(argv[1], (size_t)m * m), V0 = read_bin<double>(argv[3], (size_t)2 * m);
    std::vector<double> V((size_t)(max_steps + 3) * m, 0.0);
    std::copy(V0.begin(), V0.end(), V.begin());
    Params p;
    p.Ms = Ms.data(); p.m = m; p.V = V.data();
    p.chunk = (m + grid - 2) / (grid - 1);
    if (p.chunk > CH_MAXROWS) { fprintf(stderr, "too many rows per CTA\n"); return 3; }
    p.n_workers = (m + p.chunk - 1) / p.chunk;
    p.n_checkers = std::min(grid - p.n_workers, CH_MAXCHK);
    p.max_steps = max_steps; p.vs_cap = max_steps + 2;
    p.kw = kw; p.guard = guard; p.cap_steps = m - 1; p.tol = tol;
    p.first_check = std::min(max_steps, std::max(3 * (kw + guard) + 8, 12));
    const int slots = max_steps + 3;
    std::vector<uint4> pk((size_t)slots * p.n_workers + slots + 1 + m + 2 * (max_steps + 2), uint4{0, 0, 0, 0});
    p.pbuf = pk.data(); p.hbuf = p.pbuf + (size_t)slots * p.n_workers; p.qbuf = p.hbuf + slots + 1; p.abuf = p.qbuf + m;
    Ctl ctl; std::memset(&ctl, 0, sizeof ctl);
    p.ctl = &ctl;
    const int s_stride = max_steps * kw;
    std::vector<double> theta((size_t)CH_MAXCHK * CH_MAXWANT, 0.0), S((size_t)CH_MAXCHK * s_stride, 0.0), alpha(max_steps + 4, 0.0), beta(max_steps + 4, 0.0);
    p.theta_out = theta.data(); p.S_out = S.data(); p.s_stride = s_stride; p.alpha = alpha.data(); p.beta = beta.data();
    p.prof_on = 0;
    const int n_sm = std::max(0, p.chunk - CH_REGROWS);
    const size_t worker = sizeof(double) * ((size_t)m + (size_t)n_sm * m + (size_t)p.chunk * p.vs_cap);
    const size_t dyn = std::max(worker, sizeof(double) * checker_doubles(kw + guard, max_steps + 2));
    emu_launch_cluster(grid, CH_THREADS, grid, dyn, lanczos_chain_kernel, p);
    // Ritz vectors with the device kernel text as well (serial launch)
    FILE* f = fopen(argv[9], "wb");
    const double head[5] = {(double)ctl.status, (double)ctl.n_final, (double)ctl.winner, (double)ctl.steps_run, (double)ctl.abort};
    fwrite(head, 8, 5, f);
    fwrite(alpha.data(), 8, max_steps, f); fwrite(beta.data(), 8, max_steps, f);
    const int win = std::max(ctl.winner, 0);
    fwrite(theta.data() + (size_t)win * CH_MAXWANT, 8, CH_MAXWANT, f);
    fwrite(S.data() + (size_t)win * s_stride, 8, s_stride, f);
    fwrite(V.data(), 8, (size_t)(max_steps + 2) * m, f);
    fclose(f);
    return 0;
}
