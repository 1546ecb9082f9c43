// finalize_emu.cpp -- TEST INFRASTRUCTURE ONLY: dist_finalize_kernel / gram_finalize_kernel of csrc/k1_gram.cu (cut into
// finalize.inc by tests/test_emu.py) on the host, fed with split-K partial tiles in the tile-native layout the Gram kernels
// write.   usage: finalize_emu part.bin stat.bin n P nt n_tiles mode split_k gscale gram out.bin
//   mode: -1 upper triangle of 128x128 tiles, -3 wide tiles (virtual tiles 2w, 2w+1); gram: 0 distances, 1 Gram matrix
#include "cuda_emu.h"
constexpr int BM = 128, BN = 128;
namespace esper {
struct RowStat { float mean; float stdv; double norm2; };
namespace gram {
#include "finalize.inc"
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
    using namespace esper; using namespace esper::gram;
    if (argc < 12) return 2;
    const int n = atoi(argv[3]), P = atoi(argv[4]), nt = atoi(argv[5]), n_tiles = atoi(argv[6]), mode = atoi(argv[7]), split_k = atoi(argv[8]);
    const double gscale = atof(argv[9]);
    const int want_gram = atoi(argv[10]);
    std::vector<float> part = read_bin<float>(argv[1], (size_t)split_k * n_tiles * BM * BN);
    std::vector<double> nrm = read_bin<double>(argv[2], n);
    std::vector<RowStat> st(n);
    for (int i = 0; i < n; ++i) { st[i].mean = 0.f; st[i].stdv = 1.f; st[i].norm2 = nrm[i]; }
    std::vector<double> out((size_t)n * n, -7.0);                     // poison: every entry must be written
    if (want_gram) emu_launch_serial(n_tiles, 8, 256, gram_finalize_kernel, (const float*)part.data(), nt, n_tiles, 0, mode, split_k, n, (const RowStat*)st.data(), out.data());
    else emu_launch_serial(n_tiles, 8, 256, dist_finalize_kernel, (const float*)part.data(), nt, n_tiles, 0, mode, split_k, n, P, (const RowStat*)st.data(), gscale, out.data());
    FILE* f = fopen(argv[11], "wb"); fwrite(out.data(), 8, out.size(), f); fclose(f);
    return 0;
}
