// lanczos_emu.cpp -- TEST INFRASTRUCTURE ONLY: the persistent cooperative Lanczos kernels of csrc/k4_eigen.cu (cut into
// lanczos.inc by tests/test_emu.py; grid barrier PTX replaced by the equivalent atomics, `__shared__` by the per-block arena)
// on the CPU: ALL blocks of the grid run concurrently (they synchronise through the grid barrier), one OS thread per CUDA thread.
//   lanczos_emu Ms.bin m V0.bin steps variant grid batch out.bin     variant 0 = lanczos_steps_kernel,
//   1 = lanczos_resident_kernel<true> (rows in shared memory), 2 = lanczos_resident_kernel<false> (streamed)
//   V0.bin: [2, m] = locked vector v0 and normalised start q1.  out.bin: alpha[steps], beta[steps], V[(steps+2), m]
#define EMU_ARENA_SHARED
#define EMU_NO_WARP_SUM
#define EMU_STATIC_BYTES_DEF (96 * 1024)
#include "cuda_emu.h"
#include <algorithm>

namespace esper {
inline double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
namespace lanczos {
#include "lanczos.inc"
}
}  // namespace esper

template <class T> static std::vector<T> read_bin(const char* path, size_t n) {
    std::vector<T> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    using namespace esper::lanczos;
    if (argc < 9) return 2;
    const int m = atoi(argv[2]), steps = atoi(argv[4]), variant = atoi(argv[5]), grid = atoi(argv[6]), batch = atoi(argv[7]);
    std::vector<double> Ms = read_bin<double>(argv[1], (size_t)m * m), V0 = read_bin<double>(argv[3], (size_t)2 * m);
    const int stride = steps + 8;
    std::vector<double> V((size_t)(steps + 3) * m, 0.0), w((size_t)2 * m, 0.0), h((size_t)stride * 3 + 6 * grid + 64, 0.0);
    std::copy(V0.begin(), V0.end(), V.begin());
    double* alpha = h.data() + stride; double* beta = alpha + stride; double* part = beta + stride;
    unsigned int bar[4] = {0, 0, 0, 0};
    const int chunk_rows = (m + grid - 1) / grid;
    int rows_cached = 0; size_t dyn = 0;
    if (variant == 1) { rows_cached = std::min(chunk_rows, 14); if (chunk_rows > rows_cached + 2) { fprintf(stderr, "too many rows per CTA\n"); return 3; } dyn = sizeof(double) * ((size_t)LAN_CH + (size_t)rows_cached * m); }
    if (variant == 2) dyn = sizeof(double) * (size_t)LAN_CH;
    int done = 0;
    while (done < steps) {
        const int n = std::min(batch, steps - done);
        bar[0] = 0; bar[1] = 0;
        StepParams sp{Ms.data(), m, V.data(), w.data(), h.data(), part, alpha, beta, done + 1, n, bar, rows_cached, nullptr};
        if (variant == 0) emu_launch_cluster(grid, LAN_THREADS, grid, 0, lanczos_steps_kernel, sp);
        else if (variant == 1) emu_launch_cluster(grid, LAN_THREADS, grid, dyn, lanczos_resident_kernel<true>, sp);
        else emu_launch_cluster(grid, LAN_THREADS, grid, dyn, lanczos_resident_kernel<false>, sp);
        if (bar[1] != 0) { fprintf(stderr, "grid barrier aborted\n"); return 4; }
        done += n;
    }
    FILE* f = fopen(argv[8], "wb");
    fwrite(alpha, 8, steps, f); fwrite(beta, 8, steps, f); fwrite(V.data(), 8, (size_t)(steps + 2) * m, f);
    fclose(f);
    return 0;
}
