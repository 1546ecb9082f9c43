// chain_emu.cpp -- TEST INFRASTRUCTURE ONLY: the barrier-free Lanczos kernel of csrc/k4_chain.cu (cut into chain.inc by
// tests/test_emu.py) on the CPU.  ALL blocks run concurrently, one OS thread per CUDA thread; the packet primitives
// (st_ll / ld_ll: 16-byte {lo, tag, hi, tag} stores whose 8-byte halves are atomic) become two 64-bit atomics.
//   chain_emu Ms.bin m V0.bin kw guard tol grid max_steps out.bin [S] [merged]
//   merged = 1: lanczos_merged_kernel (one exchange per step; count-based mbarrier with remote arrives for the vector buffer)
//   S > 0: the cluster schedule (lanczos_cluster_kernel) with clusters of S consecutive blocks; distributed shared memory is the
//   peer block's arena, an mbarrier a {pending arrivals, transaction bytes, phase} record, st.async a store + complete_tx.
//   V0.bin: [2, m] = locked vector v0 and normalised start q1.
//   out.bin: ctl {status, n_final, winner, steps_run, abort} as 5 doubles | alpha[max_steps] | beta[max_steps] |
//            theta[CH_MAXWANT] of the winner | S[max_steps * kw] of the winner | V[(max_steps + 2), m]
#define EMU_ARENA_SHARED
#define EMU_NO_WARP_SUM
#define EMU_STATIC_BYTES_DEF (32 * 1024)
#define ESPER_CHAIN_EMU
#include "cuda_emu.h"
#include <algorithm>
#include <chrono>

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
static long long CH_TIMEOUT = 4000000000LL;      // nanoseconds here; EMU_TIMEOUT_S shortens it (then a timed-out wait reports itself)
constexpr int CH_MAXTESTS = 192;
constexpr int CL_MAXROWS = 17;
constexpr int CL_REGROWS = 7;
constexpr int CL_GPK = 4;
constexpr int CL_SMROWS = CL_MAXROWS - CL_REGROWS;
constexpr int CL_LD = 2048;
constexpr int CL_MAXCLUSTERS = 32;
#include "chain_structs.inc"

static thread_local const void* emu_last_wait = nullptr;          // packet or mbarrier this thread polled last
static thread_local unsigned emu_last_tag = 0;
static bool emu_clock = false;
static int emu_S = 1;                                              // cluster size of the running launch
static std::vector<std::unique_ptr<std::barrier<>>> emu_cl_bars;   // one per cluster
struct Mbar { std::atomic<long long> tx; std::atomic<int> pending; std::atomic<unsigned> phase; std::atomic<int> lock; std::atomic<int> count; };
inline void mbar_lock(Mbar* b) { int z = 0; while (!b->lock.compare_exchange_weak(z, 1, std::memory_order_acquire)) { z = 0; std::this_thread::yield(); } }
inline void mbar_unlock(Mbar* b) { b->lock.store(0, std::memory_order_release); }
inline void mbar_settle(Mbar* b) {                                 // phase completes when no arrival and no byte is pending
    if (b->pending.load() == 0 && b->tx.load() == 0) { b->pending.store(b->count.load()); b->phase.store(b->phase.load() ^ 1u, std::memory_order_release); }
}
inline int cluster_rank() { return (int)(blockIdx.x % (unsigned)emu_S); }
inline void mbar_init(Mbar* b) { b->tx.store(0); b->pending.store(1); b->count.store(1); b->phase.store(0); b->lock.store(0); }
inline void mbar_init_count(Mbar* b, unsigned int count) { b->tx.store(0); b->pending.store((int)count); b->count.store((int)count); b->phase.store(0); b->lock.store(0); }
inline void mbar_init_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void mbar_expect(Mbar* b, unsigned int bytes) {
    mbar_lock(b);
    if (b->pending.load() != 1) { fprintf(stderr, "emu: mbarrier armed twice in one phase\n"); abort(); }
    b->tx.fetch_add((long long)bytes); b->pending.store(0); mbar_settle(b);
    mbar_unlock(b);
}
inline bool mbar_test(Mbar* b, unsigned int parity) {
    const bool ok = (b->phase.load(std::memory_order_acquire) & 1u) != parity;
    if (!ok) { emu_last_wait = b; emu_last_tag = 1000u + parity + 10u * (unsigned)(b->tx.load() & 0xffff); std::this_thread::yield(); }
    return ok;
}
inline void dsm_send(double* dst_local, int rank, double v, Mbar* bar_local) {
    const unsigned peer = blockIdx.x - (blockIdx.x % (unsigned)emu_S) + (unsigned)rank;
    char* base = emu_arenas[peer].data();
    double* dst = reinterpret_cast<double*>(base + (reinterpret_cast<char*>(dst_local) - emu_arena));
    Mbar* bar = reinterpret_cast<Mbar*>(base + (reinterpret_cast<char*>(bar_local) - emu_arena));
    uint64_t bits; std::memcpy(&bits, &v, 8);
    __atomic_store_n(reinterpret_cast<uint64_t*>(dst), bits, __ATOMIC_RELAXED);
    mbar_lock(bar);
    if (bar->tx.fetch_sub(8) - 8 < -(1LL << 20)) { fprintf(stderr, "emu: transaction count underflow\n"); abort(); }
    mbar_settle(bar);
    mbar_unlock(bar);
}
inline void dsm_arrive(Mbar* bar_local, int rank) {                // remote arrive on a count-based mbarrier
    const unsigned peer = blockIdx.x - (blockIdx.x % (unsigned)emu_S) + (unsigned)rank;
    Mbar* bar = reinterpret_cast<Mbar*>(emu_arenas[peer].data() + (reinterpret_cast<char*>(bar_local) - emu_arena));
    mbar_lock(bar);
    if (bar->pending.fetch_sub(1) - 1 < 0) { fprintf(stderr, "emu: more arrivals than the mbarrier counts in one phase\n"); abort(); }
    mbar_settle(bar);
    mbar_unlock(bar);
}
inline long long wall_ns() { return 0; }
inline void cluster_sync_all() { emu_cl_bars[blockIdx.x / (unsigned)emu_S]->arrive_and_wait(); }

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
    if (!ok) { emu_last_wait = p; emu_last_tag = tag; std::this_thread::yield(); }
    return ok;
}
inline uint4 ld_ll_raw(const uint4* p) {
    const uint64_t* q = reinterpret_cast<const uint64_t*>(p);
    const uint64_t lo = __atomic_load_n(q, __ATOMIC_RELAXED), hi = __atomic_load_n(q + 1, __ATOMIC_RELAXED);
    return uint4{(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
}
inline bool ll_decode(const uint4& x, unsigned int tag, double& v) {
    const uint64_t bits = (uint64_t)x.x | ((uint64_t)x.z << 32);
    std::memcpy(&v, &bits, 8);
    return x.y == tag && x.w == tag;
}
static const uint4* emu_pk_base = nullptr; static size_t emu_off_h = 0, emu_off_q = 0, emu_off_a = 0;
static Ctl* emu_ctl = nullptr;
inline void emu_report(const char* why) {
    const char* w = reinterpret_cast<const char*>(emu_last_wait);
    const char* pk = reinterpret_cast<const char*>(emu_pk_base);
    const long long off = (w - pk) / 16;
    if (w >= emu_arena && w < emu_arena + (1 << 20))
        fprintf(stderr, "emu: block %u %s: mbarrier %lld parity %u tx %u\n", blockIdx.x, why, (long long)(w - emu_arena) / (long long)sizeof(Mbar), emu_last_tag % 10u, (emu_last_tag - 1000u) / 10u);
    else
        fprintf(stderr, "emu: block %u %s: %s index %lld tag %u\n", blockIdx.x, why,
                (size_t)off < emu_off_h ? "pbuf/cbuf" : (size_t)off < emu_off_q ? "hbuf" : (size_t)off < emu_off_a ? "qbuf" : "abuf",
                (size_t)off < emu_off_h ? off : (size_t)off < emu_off_q ? off - (long long)emu_off_h : (size_t)off < emu_off_a ? off - (long long)emu_off_q : off - (long long)emu_off_a, emu_last_tag);
}
inline unsigned int ld_word(const unsigned int* p) {
    const unsigned int v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    if (emu_clock && emu_ctl && p == &emu_ctl->abort && v != 0u && emu_last_wait) { emu_report("sees abort"); emu_last_wait = nullptr; }
    return v;
}
inline void st_word(unsigned int* p, unsigned int v) {
    if (emu_ctl && p == &emu_ctl->abort && v == 1u && emu_last_wait) emu_report("gave up");
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
inline unsigned int cas_word(unsigned int* p, unsigned int cmp, unsigned int v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline void publish_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long now() { return emu_clock ? (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0; }
inline void relax() { std::this_thread::yield(); }
inline void relax_short() { std::this_thread::yield(); }
#define CH_DEBUG_TEST(c, n, mode, track, stagnant) do { if (getenv("EMU_DEBUG_TESTS") && threadIdx.x == 0) { \
    fprintf(stderr, "checker %d: n %d mode %d track %d stagnant %d tn %.3e:", c, n, mode, (int)(track), (int)(stagnant), tn); \
    for (int i_ = 0; i_ < min(p.kw, want); ++i_) fprintf(stderr, " [%.1e %.1e]", (cs.lo[i_] - cs.prev[i_]) / tn, (cs.hi[i_] - cs.prev[i_]) / tn); \
    fprintf(stderr, "\n"); } } while (0)
#define CH_RR_BIG 1e6                   // (device: 1e100) small, so that the rescaling events of the recurrences are exercised at the test sizes
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
    const int CS = argc > 10 ? atoi(argv[10]) : 0;
    const bool merged = argc > 11 && atoi(argv[11]) != 0 && CS > 0;
    const int m = atoi(argv[2]), kw = atoi(argv[4]), guard = atoi(argv[5]);
    const double tol = atof(argv[6]);
    const int grid = atoi(argv[7]), max_steps = atoi(argv[8]);
    std::vector<double> Ms = read_bin<double>(argv[1], (size_t)m * m), V0 = read_bin<double>(argv[3], (size_t)2 * m);
    std::vector<double> V((size_t)(max_steps + 3) * m, 0.0);
    std::copy(V0.begin(), V0.end(), V.begin());
    Params p;
    p.Ms = Ms.data(); p.m = m; p.V = V.data();
    p.chunk = (m + grid - 2) / (grid - 1);
    if (p.chunk > (CS > 0 ? CL_MAXROWS : CH_MAXROWS)) { fprintf(stderr, "too many rows per CTA\n"); return 3; }
    p.n_workers = (m + p.chunk - 1) / p.chunk;
    p.n_checkers = std::min(grid - p.n_workers, CH_MAXCHK);
    p.max_steps = max_steps; p.vs_cap = max_steps + 2;
    p.kw = kw; p.guard = guard; p.cap_steps = m - 1; p.tol = tol;
    p.first_check = std::min(max_steps, std::max(3 * (kw + guard) + 8, 12));
    p.check_back = 4;
    p.check_stride = getenv("EMU_CHECK_STRIDE") ? atoi(getenv("EMU_CHECK_STRIDE")) : (p.n_checkers >= 2 ? 2 : 3);
    const int slots = max_steps + 4;
    p.cl_size = CS; p.cl_count = CS > 0 ? (p.n_workers + CS - 1) / CS : 0; p.cl_group = 1;
    while (p.cl_group < p.cl_count) p.cl_group *= 2;
    std::vector<uint4> pk((size_t)4 * slots * p.n_workers + slots + 1 + 2 * m + 2 * (max_steps + 2), uint4{0, 0, 0, 0});
    p.cbuf = pk.data();
    p.pbuf = pk.data(); p.hbuf = p.pbuf + (size_t)4 * slots * p.n_workers; p.qbuf = p.hbuf + slots + 1; p.abuf = p.qbuf + 2 * m;
    p.merged = merged ? 1 : 0; p.cb_stride = (size_t)slots * (size_t)std::max(p.cl_count, 1); p.lambda0 = 1.0; p.mg_order = getenv("EMU_MERGED_ORDER") ? atoi(getenv("EMU_MERGED_ORDER")) : 1;
    Ctl ctl; std::memset(&ctl, 0, sizeof ctl);
    p.ctl = &ctl;
    emu_ctl = &ctl; emu_pk_base = pk.data(); emu_off_h = p.hbuf - pk.data(); emu_off_q = p.qbuf - pk.data(); emu_off_a = p.abuf - pk.data();
    if (getenv("EMU_TIMEOUT_S")) { emu_clock = true; CH_TIMEOUT = (long long)(atof(getenv("EMU_TIMEOUT_S")) * 1e9); }
    const int s_stride = max_steps * kw;
    std::vector<double> theta((size_t)CH_MAXCHK * CH_MAXWANT, 0.0), S((size_t)CH_MAXCHK * s_stride, 0.0), alpha(max_steps + 4, 0.0), beta(max_steps + 4, 0.0);
    p.theta_out = theta.data(); p.S_out = S.data(); p.s_stride = s_stride; p.alpha = alpha.data(); p.beta = beta.data();
    p.prof_on = 0; p.trace = nullptr; p.trace_step = 0;
    const int n_sm = std::max(0, p.chunk - (CS > 0 ? CL_REGROWS : CH_REGROWS));
    const size_t worker = merged ? sizeof(double) * merged_worker_doubles(p.vs_cap) : CS > 0 ? sizeof(double) * cluster_worker_doubles(p.chunk, p.vs_cap) : sizeof(double) * ((size_t)m + (size_t)n_sm * m + (size_t)p.chunk * p.vs_cap);
    const size_t dyn = std::max(worker, sizeof(double) * checker_doubles(kw + guard, max_steps + 2));
    if (CS > 0) {
        if (grid % CS) { fprintf(stderr, "grid must be a multiple of the cluster size\n"); return 3; }
        emu_S = CS;
        for (int i = 0; i < grid / CS; ++i) emu_cl_bars.push_back(std::make_unique<std::barrier<>>(CS * CH_THREADS));
        if (merged) emu_launch_cluster(grid, CH_THREADS, grid, dyn, lanczos_merged_kernel, p);
        else emu_launch_cluster(grid, CH_THREADS, grid, dyn, lanczos_cluster_kernel, p);
    } else {
        emu_launch_cluster(grid, CH_THREADS, grid, dyn, lanczos_chain_kernel, p);
    }
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
