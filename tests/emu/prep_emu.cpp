// prep_emu.cpp -- TEST INFRASTRUCTURE ONLY: runs the standardise-and-split kernels of csrc/k0_prep.cu (cut into
// prep_kernels.inc by tests/test_emu.py, `__shared__` declarations rewritten to the per-block arena) on the CPU:
// stats_split_kernel<HALF> (one CTA per image); the L2 cache-policy loads / stores are plain accesses here.
//   prep_emu raw.bin n P ld standardize xbar.bin|- half scale cs hi.bin lo.bin stat.bin
#define EMU_ARENA_SHARED
#define EMU_NO_WARP_SUM
#include "cuda_emu.h"

inline float tf32_rna(float x) {                 // cvt.rna.tf32.f32: nearest, ties away from zero, 10 mantissa bits
    return emu_u2f((emu_f2u(x) + 0x1000u) & 0xFFFFE000u);
}

namespace esper {
inline float4 ld_keep(const float4* p) { return *p; }
inline float4 ld_last_use(const float4* p) { return *p; }
inline void st_stream(uint2* p, uint2 v) { *p = v; }
inline void st_stream(float4* p, float4 v) { *p = v; }
#include "devutil_body.inc"
#include "prep_kernels.inc"
}  // namespace esper

template <class T> static std::vector<T> read_bin(const char* path, size_t n) {
    std::vector<T> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}
static void write_bin(const char* path, const void* p, size_t bytes) { FILE* f = fopen(path, "wb"); fwrite(p, 1, bytes, f); fclose(f); }

int main(int argc, char** argv) {
    using namespace esper;
    if (argc < 13) return 2;
    const int n = atoi(argv[2]), P = atoi(argv[3]), ld = atoi(argv[4]), standardize = atoi(argv[5]);
    const bool center = argv[6][0] != '-';
    const int half = atoi(argv[7]);
    const float scale = (float)atof(argv[8]);
    const int cs = atoi(argv[9]);
    std::vector<float> raw = read_bin<float>(argv[1], (size_t)n * P), xbar;
    if (center) xbar = read_bin<float>(argv[6], P);
    const size_t esz = half ? 2 : 4;
    std::vector<char> hi((size_t)n * ld * esz, 0x55), lo((size_t)n * ld * esz, 0x55);     // poison: every element must be written
    std::vector<RowStat> st(n);
    const float* xb = center ? xbar.data() : nullptr;
    if (cs == 0) {
        if (half) emu_launch_cluster(n, 512, 1, 0, stats_split_kernel<true>, (const float*)raw.data(), n, P, ld, standardize, st.data(), xb, (int)center, scale, (void*)hi.data(), (void*)lo.data());
        else emu_launch_cluster(n, 512, 1, 0, stats_split_kernel<false>, (const float*)raw.data(), n, P, ld, standardize, st.data(), xb, (int)center, scale, (void*)hi.data(), (void*)lo.data());
    } else {
        fprintf(stderr, "prep_emu: cs must be 0\n");
        return 2;
    }
    write_bin(argv[10], hi.data(), hi.size());
    write_bin(argv[11], lo.data(), lo.size());
    std::vector<double> so((size_t)n * 3);
    for (int i = 0; i < n; ++i) { so[3 * i] = st[i].mean; so[3 * i + 1] = st[i].stdv; so[3 * i + 2] = st[i].norm2; }
    write_bin(argv[12], so.data(), so.size() * 8);
    return 0;
}
