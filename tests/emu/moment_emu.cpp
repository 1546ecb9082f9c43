// moment_emu.cpp -- TEST INFRASTRUCTURE ONLY: runs the moment-sweep kernels of csrc/k2_sweep.cu (cut into
// moment_kernels.inc by tests/test_emu.py) on the CPU through cuda_emu.h, with the launch chain of run_sweep_moments /
// run_moment_{stats,accum}_rows.   usage: moment_emu D.bin m rows sym coef.bin n_eps thr out.bin sm_count
//   D.bin [rows, m] float64, coef.bin [n_eps] float64 -> out.bin: [status, bins, logsum[n_eps], table[MB*MV]]
#include "cuda_emu.h"

namespace esper {
namespace sweep {
#include "moment_kernels.inc"
}  // namespace sweep
}  // namespace esper

static std::vector<double> read_bin(const char* path, size_t n) {
    std::vector<double> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 8, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    using namespace esper::sweep;
    if (argc < 10) return 2;
    const int m = atoi(argv[2]), rows = atoi(argv[3]), sym = atoi(argv[4]), n_eps = atoi(argv[6]), sms = atoi(argv[9]);
    const double thr = atof(argv[7]);
    std::vector<double> D = read_bin(argv[1], (size_t)rows * m), coef = read_bin(argv[5], n_eps);
    int grid = sms * 2;
    if (grid > 480) grid = 480;
    if (grid > (rows + 1) / 2) grid = (rows + 1) / 2;
    if (grid < 1) grid = 1;
    std::vector<double> stat((size_t)4 * grid), plan(PLAN_LEN, 0.0), part((size_t)grid * MB * MV, 0.0), out(n_eps, 0.0), table(MB * MV, 0.0);
    emu_launch(grid, MTHREADS, moment_stats_kernel, (const double*)D.data(), rows, m, sym, stat.data());
    emu_launch(1, 32, moment_plan_kernel, (const double*)stat.data(), grid, (const double*)coef.data(), n_eps, thr, plan.data());
    emu_launch(grid, MTHREADS, moment_accum_kernel, (const double*)D.data(), rows, m, sym, (const double*)coef.data(), thr,
               (const double*)plan.data(), part.data());
    emu_launch(1, 1024, moment_final_kernel, (const double*)part.data(), grid, (const double*)coef.data(), n_eps, thr,
               (const double*)plan.data(), out.data());
    emu_launch(1, 1024, moment_table_kernel, (const double*)part.data(), grid, (int)plan[1], table.data());
    FILE* f = fopen(argv[8], "wb");
    fwrite(&plan[0], 8, 1, f); fwrite(&plan[1], 8, 1, f);
    fwrite(out.data(), 8, n_eps, f);
    fwrite(table.data(), 8, table.size(), f);
    fwrite(plan.data(), 8, plan.size(), f);
    fclose(f);
    return 0;
}
