// tilemap_emu.cpp -- TEST INFRASTRUCTURE ONLY: the tile index maps of csrc/k1_gram.cu (cut into tilemap.inc by
// tests/test_emu.py) checked exhaustively on the host: every 128x128 tile of the upper triangle is produced exactly once by
// the wide-tile enumeration (virtual tiles 2w, 2w+1 of wide tile w), wide_coords agrees with tile_map, and the row-block map
// covers its rectangle with the mirrored tiles marked transposed.   usage: tilemap_emu max_nt
#include <cstdio>
#include <cstdlib>
#include <vector>
#define __device__
#define __host__
#define __forceinline__ inline
constexpr int BM = 128, BN = 128;
#include "tilemap.inc"

int main(int argc, char** argv) {
    const int max_nt = argc > 1 ? atoi(argv[1]) : 64;
    for (int nt = 1; nt <= max_nt; ++nt) {
        // wide enumeration
        const int nw = wide_tile_count(nt);
        std::vector<int> seen((size_t)nt * nt, 0);
        for (int w = 0; w < nw; ++w) {
            int ti, tc; wide_coords(w, nt, ti, tc);
            if (ti < 0 || ti >= nt || tc < ti / 2 || 2 * tc >= nt + 1) { printf("nt=%d wide %d -> bad (%d,%d)\n", nt, w, ti, tc); return 1; }
            for (int h = 0; h < 2; ++h) {
                int ta, tb; bool tr; tile_map(2 * w + h, nt, WIDE_TRI, ta, tb, tr);
                const int col = 2 * tc + h;
                const bool valid = col >= ti && col < nt;
                if (valid != (ta >= 0)) { printf("nt=%d wide %d h=%d validity mismatch\n", nt, w, h); return 1; }
                if (ta >= 0) {
                    if (ta != ti || tb != col || tr) { printf("nt=%d wide %d h=%d -> (%d,%d) expected (%d,%d)\n", nt, w, h, ta, tb, ti, col); return 1; }
                    seen[(size_t)ta * nt + tb]++;
                }
            }
        }
        for (int i = 0; i < nt; ++i)
            for (int j = 0; j < nt; ++j)
                if (seen[(size_t)i * nt + j] != (j >= i ? 1 : 0)) { printf("nt=%d tile (%d,%d) produced %d times\n", nt, i, j, seen[(size_t)i * nt + j]); return 1; }
        // plain upper triangle
        std::vector<int> seen2((size_t)nt * nt, 0);
        for (int t = 0; t < nt * (nt + 1) / 2; ++t) { int ta, tb; bool tr; tile_map(t, nt, -1, ta, tb, tr); if (ta > tb || tr) return 2; seen2[(size_t)ta * nt + tb]++; }
        for (int i = 0; i < nt; ++i) for (int j = i; j < nt; ++j) if (seen2[(size_t)i * nt + j] != 1) return 3;
        // row block starting at every tile row: rectangle coverage, mirrored tiles transposed
        for (int ti0 = 0; ti0 < nt; ++ti0)
            for (int t = 0; t < (nt - ti0) * nt; ++t) {
                int ta, tb; bool tr; tile_map(t, nt, ti0, ta, tb, tr);
                const int ti = ti0 + t / nt, tj = t % nt;
                if (tj >= ti ? (ta != ti || tb != tj || tr) : (ta != tj || tb != ti || !tr)) return 4;
            }
    }
    printf("ok %d\n", max_nt);
    return 0;
}
