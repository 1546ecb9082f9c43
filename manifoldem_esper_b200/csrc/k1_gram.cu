// k1_gram.cu -- K1: symmetric 3xTF32 Gram contraction on tcgen05 tensor cores (TMA -> smem -> UMMA ->
// TMEM), split-K partial tiles combined in FP64 by the distance epilogue, plus the direct-difference
// refinement for cancelling pairs.  Replaces the double loop of 1_Embedding/DM/Distances.py:96-104.
//
// Work decomposition
//   G = C C^T with C [n, K] the centred standardised images (K = ld, zero padded), only the upper
//   triangle of 128x128 output tiles is computed.  A work unit is (k-chunk, tile); units are ordered
//   chunk-major so that the CTAs resident at any time stream the SAME K range of the planes and the
//   operand working set (n x chunk x 8 B) stays in L2.  Each unit accumulates
//       hi*hi + hi*lo + lo*hi          (3xTF32: operands split into two TF32 values, FP32 accumulate)
//   over its chunk in TMEM and writes an FP32 partial tile; chains are kept short (<= ~2048 pixels)
//   because tensor-core FP32 accumulation is not round-to-nearest; partials are summed in FP64.
//
// Kernel structure (one CTA per SM, persistent over units, 320 threads):
//   warp 0     TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx; elect.sync lane)
//   warp 1     MMA issuer     (tcgen05.mma.cta_group::1.kind::tf32, elect.sync lane, warp-uniform loop) + TMEM alloc
//   warps 2-9  epilogue       (tcgen05.ld 32x32b -> registers -> coalesced FP32 partial tile); warps w and
//                             w+4 share a TMEM lane quadrant and take 64 of the 128 columns each
//   smem ring: 3 stages x {A_hi, A_lo, B_hi, B_lo} x (128 rows x 32 fp32) = 3 x 64 KB
//   TMEM: 4 accumulator buffers x 128 columns (the MMA warp runs up to 3 groups ahead of the drains)
#include <cuda.h>
#include "common.cuh"
#include "devutil.cuh"

namespace esper {
namespace gram {

constexpr int BM = 128;                 // tile rows  (UMMA M)
constexpr int BN = 128;                 // tile cols  (UMMA N)
constexpr int BK = 32;                  // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UK = 8;                   // UMMA K for kind::tf32 (32 bytes)
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4; // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;
constexpr int ACC_STAGES = 4;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512: the whole tensor memory of the SM (one CTA per SM)
constexpr int EPI_WARPS = 8;                // two warps per TMEM lane quadrant, 64 columns each
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// One elected lane of a converged warp (warp-uniform predicate: the tensor-core / TMA instructions guarded by it take
// their operands from uniform registers; under an `if (lane == 0)` branch ptxas emits a per-instruction
// ELECT / R2UR waterfall instead, ~12 instructions per tcgen05.mma, and the issuing thread becomes the bottleneck).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                     // layout: SWIZZLE_128B
    return d;
}
// Instruction descriptor: D=F32, A=B=TF32, both K-major, N=128, M=128.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Upper-triangular tile index t -> (ti, tj), tj >= ti, row-major over the triangle.
__device__ __forceinline__ void tile_coords(int t, int nt, int& ti, int& tj) {
    int i = 0, rem = t;
    while (rem >= nt - i) { rem -= nt - i; ++i; }
    ti = i; tj = i + rem;
}
// Work tile t -> canonical upper-triangle tile (ta <= tb) whose Gram block is computed.
//   rect_ti0 < 0 : the whole upper triangle (single-GPU path)
//   rect_ti0 >= 0: row-block path, tiles (rect_ti0 + t / nt, t % nt) of a block of tile rows; a tile below the
//                  diagonal is computed as its mirror image (same operands, same bits) and stored transposed.
//   rect_ti0 == -2: general product A B^T (two operand matrices): tile (t / nt, t % nt), nt = tile columns (B tiles)
constexpr int FULL_RECT = -2;
__device__ __forceinline__ void tile_map(int t, int nt, int rect_ti0, int& ta, int& tb, bool& transposed) {
    if (rect_ti0 == FULL_RECT) { ta = t / nt; tb = t % nt; transposed = false; return; }
    if (rect_ti0 == -3) {                      // WIDE_TRI: t = virtual 128x128 tile 2 w + h of wide tile w (see gram_wide_kernel)
        const int nt2 = (nt + 1) / 2;
        int i = 0, rem = t >> 1;
        while (rem >= nt2 - i / 2) { rem -= nt2 - i / 2; ++i; }
        ta = i; tb = 2 * (i / 2 + rem) + (t & 1); transposed = false;
        if (tb < ta || tb >= nt) ta = -1;       // redundant (below the diagonal) or beyond the matrix: nothing to store
        return;
    }
    if (rect_ti0 < 0) { tile_coords(t, nt, ta, tb); transposed = false; return; }
    const int ti = rect_ti0 + t / nt, tj = t % nt;
    if (tj >= ti) { ta = ti; tb = tj; transposed = false; } else { ta = tj; tb = ti; transposed = true; }
}

// ---- wide (128 x 256) tiles: the default for the full symmetric matrix (ESPER_GRAM_WIDE=0 switches back) ---------------
// Measured at the benchmark shape: 1.107 ms against 1.259 ms for the 128x128 kernel on the same box, distances bit-identical
// (same per-element accumulation order); `pytest -m gpu` distance / PCA / full-size parity tests green with it.
// Motivation (profiles/r1_gram_chain_cut_cost.md): the 128x128 kernel is bound by operand feed from L2 (60 kB per k-block
// per SM) and pays ~330 cycles per accumulation-chain cut.  A 128x256 tile reads 96 kB for twice the MMA work (25 % fewer
// operand bytes per flop) and halves the cuts per flop.  The upper triangle is covered by tiles (ti, tc): rows
// [128 ti, 128 ti + 128), columns [256 tc, 256 tc + 256), tc >= ti / 2; a wide tile is stored as two 128x128 partial tiles
// ("virtual tiles" 2 t and 2 t + 1, columns tj = 2 tc + h), so the epilogue kernels keep their layout; virtual tiles below
// the diagonal (odd ti, h = 0: computed but redundant, +6 % work) or beyond the matrix are skipped there.
constexpr int BN_W = 256;
constexpr int STAGES_W = 2;
constexpr int STAGE_W_BYTES = 2 * TILE_BYTES + 2 * (2 * TILE_BYTES);     // A hi, lo (16 kB each) + B hi, lo (32 kB each) = 96 kB
constexpr int ACC_W = 2;                                                  // 2 x 256 TMEM columns
constexpr int SMEM_W_BYTES = STAGES_W * STAGE_W_BYTES + 1024 + 256;
constexpr uint32_t IDESC_W = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN_W >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int WIDE_TRI = -3;

__host__ __device__ inline int wide_tile_count(int nt) {
    const int nt2 = (nt + 1) / 2;
    int tot = 0;
    for (int i = 0; i < nt; ++i) tot += nt2 - i / 2;
    return tot;
}
__device__ __forceinline__ void wide_coords(int t, int nt, int& ti, int& tc) {
    const int nt2 = (nt + 1) / 2;
    int i = 0, rem = t;
    while (rem >= nt2 - i / 2) { rem -= nt2 - i / 2; ++i; }
    ti = i; tc = i / 2 + rem;
}

__device__ __forceinline__ void umma_tf32_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC_W), "r"(accumulate)
        : "memory");
}
// 3xFP16 variant (opt-in, esper_gram_config(ctx, 1) / ESPER_GRAM_F16=1): operands are FP16 hi/lo planes (same 11-bit
// significands as TF32, see stats_split_kernel<true>), D = F32, A = B = F16 (format 0), K = 16 per instruction: the same
// 32 bytes per operand row per MMA as kind::tf32, so descriptors, swizzle and the k-step offsets are unchanged while a
// 128-byte row now holds 64 pixels -- half the MMA issue slots and half the operand bytes per pixel.
constexpr uint32_t IDESC_WH = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN_W >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int BK_H = 64;                    // FP16 elements per 128-byte swizzle row
__device__ __forceinline__ void umma_f16_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC_WH), "r"(accumulate)
        : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    if (F16) umma_f16_w(tmem_d, adesc, bdesc, accumulate); else umma_tf32_w(tmem_d, adesc, bdesc, accumulate);
}

template <bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_wide_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                 const __grid_constant__ CUtensorMap mapw_hi, const __grid_constant__ CUtensorMap mapw_lo,
                 int nt, int n_tiles, int tile0, int nkb, int split_k, int drain_kb, float* __restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES_W * STAGE_W_BYTES;
    const uint32_t full_bar = bars;                                     // STAGES_W x 8 B
    const uint32_t empty_bar = bars + 8 * STAGES_W;
    const uint32_t tfull_bar = bars + 16 * STAGES_W;                    // ACC_W x 8 B
    const uint32_t tempty_bar = tfull_bar + 8 * ACC_W;
    const uint32_t tmem_slot = tempty_bar + 8 * ACC_W;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_units = n_tiles * split_k;
    constexpr int BKE = F16 ? BK_H : BK;                                // elements per k-block (one 128-byte row)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_hi); tma_prefetch_desc(&map_lo); tma_prefetch_desc(&mapw_hi); tma_prefetch_desc(&mapw_lo);
        for (int s = 0; s < STAGES_W; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int a = 0; a < ACC_W; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(ACC_W * BN_W));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int chunk = u / n_tiles;
            int ti, tc; wide_coords(tile0 + (u % n_tiles), nt, ti, tc);
            const int kb0 = (int)((long long)nkb * chunk / split_k);
            const int kb1 = (int)((long long)nkb * (chunk + 1) / split_k);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                if (elect_one_sync()) {
                    const uint32_t sa = base + stage * STAGE_W_BYTES;
                    const uint32_t fb = full_bar + 8 * stage;
                    mbar_expect_tx(fb, STAGE_W_BYTES);
                    tma_load_2d(sa, &map_hi, fb, kb * BKE, ti * BM);
                    tma_load_2d(sa + TILE_BYTES, &map_lo, fb, kb * BKE, ti * BM);
                    tma_load_2d(sa + 2 * TILE_BYTES, &mapw_hi, fb, kb * BKE, tc * BN_W);
                    tma_load_2d(sa + 4 * TILE_BYTES, &mapw_lo, fb, kb * BKE, tc * BN_W);
                }
                __syncwarp();
                if (++stage == STAGES_W) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int chunk = u / n_tiles;
            const int kb0 = (int)((long long)nkb * chunk / split_k);
            const int kb1 = (int)((long long)nkb * (chunk + 1) / split_k);
            for (int g0 = kb0; g0 < kb1; g0 += drain_kb) {
                const int g1 = min(g0 + drain_kb, kb1);
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                const uint32_t tmem_d = tmem_base + acc * BN_W;
                for (int kb = g0; kb < g1; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * STAGE_W_BYTES;
                    const uint64_t dah0 = make_smem_desc(sa), dal0 = make_smem_desc(sa + TILE_BYTES);
                    const uint64_t dbh0 = make_smem_desc(sa + 2 * TILE_BYTES), dbl0 = make_smem_desc(sa + 4 * TILE_BYTES);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint64_t off = (uint64_t)((k * UK * 4) >> 4);
                            umma_w<F16>(tmem_d, dal0 + off, dbh0 + off, (kb > g0 || k > 0) ? 1u : 0u);
                            umma_w<F16>(tmem_d, dah0 + off, dbl0 + off, 1);
                            umma_w<F16>(tmem_d, dah0 + off, dbh0 + off, 1);
                        }
                        umma_commit(empty_bar + 8 * stage);
                        if (kb == g1 - 1) umma_commit(tfull_bar + 8 * acc);
                    }
                    __syncwarp();
                    if (++stage == STAGES_W) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_W) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // epilogue warps 2..9: quadrant = warp & 3 (TMEM lanes), half = which 128 of the 256 accumulator columns
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        float sum[BN];
#pragma unroll
        for (int q = 0; q < BN; ++q) sum[q] = 0.f;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int chunk = u / n_tiles;
            const int kb0 = (int)((long long)nkb * chunk / split_k);
            const int kb1 = (int)((long long)nkb * (chunk + 1) / split_k);
            for (int g0 = kb0; g0 < kb1; g0 += drain_kb) {
                mbar_wait(tfull_bar + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN_W + half * BN);
#pragma unroll
                for (int cc = 0; cc < BN / 32; ++cc) {
                    uint32_t r[32];
                    tmem_ld32(ta + cc * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) sum[cc * 32 + q] = __fadd_rn(sum[cc * 32 + q], __uint_as_float(r[q]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
                if (++acc == ACC_W) { acc = 0; acc_phase ^= 1; }
            }
            float4* out = reinterpret_cast<float4*>(part + ((size_t)u * 2 + half) * (BM * BN));
#pragma unroll
            for (int q4 = 0; q4 < BN / 4; ++q4) {
                out[(size_t)q4 * BM + row] = make_float4(sum[4 * q4], sum[4 * q4 + 1], sum[4 * q4 + 2], sum[4 * q4 + 3]);
                sum[4 * q4] = 0.f; sum[4 * q4 + 1] = 0.f; sum[4 * q4 + 2] = 0.f; sum[4 * q4 + 3] = 0.f;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ACC_W * BN_W));
    }
}

// ------------------------------------------------------------------------------------------------
// The Gram kernel
//   part: [split_k * n_tiles] tiles of 128x128 FP32 in "tile-native" order: element (r, c) of a
//         tile lives at ((c>>2)*128 + r)*4 + (c&3)  (so that a warp's 32 rows store 512 contiguous B)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t IDESC_H = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC_H), "r"(accumulate)
        : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_sq(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    if (F16) umma_f16(tmem_d, adesc, bdesc, accumulate); else umma_tf32(tmem_d, adesc, bdesc, accumulate);
}

// F16 = true: FP16 operand planes (see gram_wide_kernel<true>); used by the row-block path when esper_gram_config(ctx, 1).
template <bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
            const __grid_constant__ CUtensorMap mapb_hi, const __grid_constant__ CUtensorMap mapb_lo,
            int nt, int n_tiles, int tile0, int rect_ti0, int nkb, int split_k, int drain_kb_arg, float* __restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    const bool skip_drain_loads = drain_kb_arg < 0;      // diagnostics only (ESPER_DRAIN_KB=-d): groups without TMEM reads
    const int drain_kb = drain_kb_arg < 0 ? -drain_kb_arg : drain_kb_arg;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // 1024-byte aligned operand ring
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    const uint32_t full_bar = bars;                                     // STAGES x 8 B
    const uint32_t empty_bar = bars + 8 * STAGES;                       // STAGES x 8 B
    const uint32_t tfull_bar = bars + 16 * STAGES;                      // ACC_STAGES x 8 B
    const uint32_t tempty_bar = tfull_bar + 8 * ACC_STAGES;             // ACC_STAGES x 8 B
    const uint32_t tmem_slot = tempty_bar + 8 * ACC_STAGES;             // 4 B
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_units = n_tiles * split_k;
    constexpr int BKE = F16 ? BK_H : BK;                                // elements per k-block (one 128-byte row)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_hi);
        tma_prefetch_desc(&map_lo);
        tma_prefetch_desc(&mapb_hi);
        tma_prefetch_desc(&mapb_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer (whole warp in the loop, one elected lane issues) =====================
        int stage = 0; uint32_t phase = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int chunk = u / n_tiles, t = tile0 + (u % n_tiles);
            int ti, tj; bool tr_; tile_map(t, nt, rect_ti0, ti, tj, tr_);
            const bool diag = (ti == tj) && rect_ti0 != FULL_RECT;
            const int kb0 = (int)((long long)nkb * chunk / split_k);
            const int kb1 = (int)((long long)nkb * (chunk + 1) / split_k);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                if (elect_one_sync()) {
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint32_t fb = full_bar + 8 * stage;
                    mbar_expect_tx(fb, diag ? 2 * TILE_BYTES : 4 * TILE_BYTES);
                    tma_load_2d(sa, &map_hi, fb, kb * BKE, ti * BM);
                    tma_load_2d(sa + TILE_BYTES, &map_lo, fb, kb * BKE, ti * BM);
                    if (!diag) {
                        tma_load_2d(sa + 2 * TILE_BYTES, &mapb_hi, fb, kb * BKE, tj * BN);
                        tma_load_2d(sa + 3 * TILE_BYTES, &mapb_lo, fb, kb * BKE, tj * BN);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp in the loop, one elected lane issues) =====================
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int chunk = u / n_tiles, t = tile0 + (u % n_tiles);
            int ti, tj; bool tr_; tile_map(t, nt, rect_ti0, ti, tj, tr_);
            const bool diag = (ti == tj) && rect_ti0 != FULL_RECT;
            const int kb0 = (int)((long long)nkb * chunk / split_k);
            const int kb1 = (int)((long long)nkb * (chunk + 1) / split_k);
            // Tensor-core FP32 accumulation rounds toward zero, which biases a long chain in proportion to the
            // accumulated magnitude (i.e. to the signal).  The chain is therefore cut every `drain_kb` k-blocks:
            // the accumulator buffer is handed to the epilogue warps, which add it in round-to-nearest FP32 into
            // register accumulators while the next groups run in the other TMEM buffers.
            for (int g0 = kb0; g0 < kb1; g0 += drain_kb) {
                const int g1 = min(g0 + drain_kb, kb1);
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = g0; kb < g1; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint32_t a_hi = sa, a_lo = sa + TILE_BYTES;
                    const uint32_t b_hi = diag ? a_hi : sa + 2 * TILE_BYTES;
                    const uint32_t b_lo = diag ? a_lo : sa + 3 * TILE_BYTES;
                    const uint64_t dah0 = make_smem_desc(a_hi), dal0 = make_smem_desc(a_lo);
                    const uint64_t dbh0 = make_smem_desc(b_hi), dbl0 = make_smem_desc(b_lo);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint64_t off = (uint64_t)((k * UK * 4) >> 4);   // 32 bytes per UMMA K step, in 16-byte units
                            umma_sq<F16>(tmem_d, dal0 + off, dbh0 + off, (kb > g0 || k > 0) ? 1u : 0u);   // small cross terms first
                            umma_sq<F16>(tmem_d, dah0 + off, dbl0 + off, 1);
                            umma_sq<F16>(tmem_d, dah0 + off, dbh0 + off, 1);
                        }
                        umma_commit(empty_bar + 8 * stage);            // frees the smem stage when the MMAs retire
                        if (kb == g1 - 1) umma_commit(tfull_bar + 8 * acc);   // group complete -> epilogue drains it
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;          // which 64 of the 128 accumulator columns
        const int row = quad * 32 + lane;          // tile row held by this thread
        constexpr int HC = BN / 2;
        int acc = 0; uint32_t acc_phase = 0;
        float sum[HC];                              // round-to-nearest FP32 accumulators of this thread's half row
#pragma unroll
        for (int q = 0; q < HC; ++q) sum[q] = 0.f;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int chunk = u / n_tiles;
            const int kb0 = (int)((long long)nkb * chunk / split_k);
            const int kb1 = (int)((long long)nkb * (chunk + 1) / split_k);
            for (int g0 = kb0; g0 < kb1; g0 += drain_kb) {
                mbar_wait(tfull_bar + 8 * acc, acc_phase);
                tc_fence_after();
                uint32_t r0[32], r1[32];
                const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * HC);
                if (!skip_drain_loads) {
                    tmem_ld32(ta, r0);              // both loads in flight before the wait
                    tmem_ld32(ta + 32, r1);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) { r0[q] = 0u; r1[q] = 0u; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);     // values are in registers: release the buffer
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    sum[q] = __fadd_rn(sum[q], __uint_as_float(r0[q]));
                    sum[32 + q] = __fadd_rn(sum[32 + q], __uint_as_float(r1[q]));
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
            float4* out = reinterpret_cast<float4*>(part + (size_t)u * (BM * BN)) + (size_t)half * (HC / 4) * BM;
#pragma unroll
            for (int q4 = 0; q4 < HC / 4; ++q4) {
                out[(size_t)q4 * BM + row] = make_float4(sum[4 * q4], sum[4 * q4 + 1], sum[4 * q4 + 2], sum[4 * q4 + 3]);
                sum[4 * q4] = 0.f; sum[4 * q4 + 1] = 0.f; sum[4 * q4 + 2] = 0.f; sum[4 * q4 + 3] = 0.f;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------
// Distance epilogue: D_ij = sqrt(max(|c_i|^2 + |c_j|^2 - 2 sum_chunks G_ij, 0) / P), mirrored, diag 0.
// One block per tile; thread -> (row r, 4 columns).  FP64 combination in fixed chunk order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dist_finalize_kernel(const float* __restrict__ part, int nt, int n_tiles, int tile0, int rect_ti0, int split_k, int n, int P,
                     const RowStat* __restrict__ st, double gscale /*exact power of two; 1 unless the operands were scaled*/,
                     double* __restrict__ D) {
    const int lt = blockIdx.x;
    const int e_per = (BM * (BN / 4)) / gridDim.y;          // slice of the tile handled by this block
    int ti, tj; bool transposed; tile_map(tile0 + lt, nt, rect_ti0, ti, tj, transposed);
    if (ti < 0) return;                                     // wide-tile mode: redundant / out-of-range virtual tile
    const bool diag = (ti == tj);
    const bool rect = rect_ti0 >= 0;
    const size_t row_off = rect ? (size_t)rect_ti0 * BM : 0;   // first global row held by D in the row-block path
    const double invP = 1.0 / (double)P;
    for (int e = blockIdx.y * e_per + threadIdx.x; e < (blockIdx.y + 1) * e_per; e += blockDim.x) {
        const int r = e & (BM - 1), c4 = e >> 7;
        const int i = ti * BM + r;
        if (i >= n) continue;
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        for (int ch = 0; ch < split_k; ++ch) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)(ch * n_tiles + lt) * (BM * BN)) + e);
            g[0] += (double)v.x; g[1] += (double)v.y; g[2] += (double)v.z; g[3] += (double)v.w;
        }
        const double ni = st[i].norm2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c4 * 4 + q;
            const int j = tj * BN + c;
            if (j >= n) continue;
            if (diag && c <= r) { if (c == r) D[((size_t)i - row_off) * n + i] = 0.0; continue; }
            double d2 = (ni + st[j].norm2 - 2.0 * (g[q] * gscale)) * invP;
            d2 = d2 > 0.0 ? d2 : 0.0;
            const double d = sqrt(d2);
            if (!rect || diag) {                    // both (i, j) and (j, i) live in this D
                D[((size_t)i - row_off) * n + j] = d;
                D[((size_t)j - row_off) * n + i] = d;
            } else if (!transposed) {
                D[((size_t)i - row_off) * n + j] = d;
            } else {
                D[((size_t)j - row_off) * n + i] = d;
            }
        }
    }
}

// Gram epilogue of the PCA path (PCA.py:58: the right singular vectors of Y are the eigenvectors of Y^T Y):
// G_ij = sum_chunks partial, mirrored; the diagonal is the FP64 |c_i|^2 of the prep pass.
__global__ void __launch_bounds__(256)
gram_finalize_kernel(const float* __restrict__ part, int nt, int n_tiles, int tile0, int mode, int split_k, int n,
                     const RowStat* __restrict__ st, double* __restrict__ G) {
    const int lt = blockIdx.x;
    const int e_per = (BM * (BN / 4)) / gridDim.y;
    int ti, tj; bool transposed; tile_map(tile0 + lt, nt, mode, ti, tj, transposed);
    if (ti < 0) return;
    const bool diag = (ti == tj);
    for (int e = blockIdx.y * e_per + threadIdx.x; e < (blockIdx.y + 1) * e_per; e += blockDim.x) {
        const int r = e & (BM - 1), c4 = e >> 7;
        const int i = ti * BM + r;
        if (i >= n) continue;
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        for (int ch = 0; ch < split_k; ++ch) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)(ch * n_tiles + lt) * (BM * BN)) + e);
            g[0] += (double)v.x; g[1] += (double)v.y; g[2] += (double)v.z; g[3] += (double)v.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c4 * 4 + q;
            const int j = tj * BN + c;
            if (j >= n) continue;
            if (diag && c <= r) { if (c == r) G[(size_t)i * n + i] = st[i].norm2; continue; }
            G[(size_t)i * n + j] = g[q];
            G[(size_t)j * n + i] = g[q];
        }
    }
}

// General product epilogue (CTF branch): M = sum of the split-K partials, FP64.  sym: upper-triangle tiles of a
// symmetric product, mirrored, diagonal from the partials; else tile (t / nt, t % nt) of an [na, nb] product.
__global__ void __launch_bounds__(256)
product_store_kernel(const float* __restrict__ part, int nt, int n_tiles, int tile0, int sym, int split_k, int na, int nb,
                     double* __restrict__ M) {
    const int lt = blockIdx.x;
    const int e_per = (BM * (BN / 4)) / gridDim.y;
    int ti, tj; bool transposed; tile_map(tile0 + lt, nt, sym ? -1 : FULL_RECT, ti, tj, transposed);
    const bool diag = sym && (ti == tj);
    for (int e = blockIdx.y * e_per + threadIdx.x; e < (blockIdx.y + 1) * e_per; e += blockDim.x) {
        const int r = e & (BM - 1), c4 = e >> 7;
        const int i = ti * BM + r;
        if (i >= na) continue;
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        for (int ch = 0; ch < split_k; ++ch) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)(ch * n_tiles + lt) * (BM * BN)) + e);
            g[0] += (double)v.x; g[1] += (double)v.y; g[2] += (double)v.z; g[3] += (double)v.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c4 * 4 + q;
            const int j = tj * BN + c;
            if (j >= nb) continue;
            if (diag && c < r) continue;
            M[(size_t)i * nb + j] = g[q];
            if (sym && j != i) M[(size_t)j * nb + i] = g[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Refinement of cancelling pairs by direct differences (the reference's own arithmetic:
// float64 sum of squared differences of the float32 standardised images).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flag_pairs_kernel(const double* __restrict__ D, int n, int row0, int nrows, int rect, int P,
                  const RowStat* __restrict__ st, double tau,
                  int2* __restrict__ pairs, unsigned int* __restrict__ count, unsigned int cap) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nrows * n) return;
    const int i = row0 + (int)(idx / n), j = (int)(idx % n);
    if (rect ? (j == i) : (j <= i)) return;
    const double d = D[idx];
    if (d * d * (double)P < tau * (st[i].norm2 + st[j].norm2)) {
        const unsigned int k = atomicAdd(count, 1u);
        if (k < cap) pairs[k] = make_int2(i, j);
    }
}

__global__ void __launch_bounds__(256)
refine_pairs_kernel(const float* __restrict__ raw, int n, int row0, int rect, int P, const RowStat* __restrict__ st,
                    const int2* __restrict__ pairs, const unsigned int* __restrict__ count, unsigned int cap,
                    double* __restrict__ D) {
    __shared__ double sh[32];
    unsigned int total = *count;
    if (total > cap) total = cap;
    for (unsigned int k = blockIdx.x; k < total; k += gridDim.x) {
        const int2 pr = pairs[k];
        const float* a = raw + (size_t)pr.x * P;
        const float* b = raw + (size_t)pr.y * P;
        const float ma = st[pr.x].mean, sa = st[pr.x].stdv, mb = st[pr.y].mean, sb = st[pr.y].stdv;
        double acc = 0.0;
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const double df = (double)zscore(__ldg(a + p), ma, sa) - (double)zscore(__ldg(b + p), mb, sb);
            acc += df * df;
        }
        acc = block_sum(acc, sh);
        if (threadIdx.x == 0) {
            const double d = sqrt(acc / (double)P);
            D[((size_t)pr.x - row0) * n + pr.y] = d;
            if (!rect) D[(size_t)pr.y * n + pr.x] = d;
        }
    }
}

}  // namespace gram

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

static int make_plane_map(esper_ctx* c, CUtensorMap* map, float* plane, int rows_pad, int ld, int box_rows = gram::BM) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(c, ESPER_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows_pad};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)gram::BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, ESPER_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return ESPER_OK;
}

static int make_plane_map_h(esper_ctx* c, CUtensorMap* map, void* plane, int rows_pad, int ldh, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(c, ESPER_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)ldh, (cuuint64_t)rows_pad};
    cuuint64_t gstride[1] = {(cuuint64_t)ldh * 2};
    cuuint32_t box[2] = {(cuuint32_t)gram::BK_H, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, plane, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, ESPER_E_CUDA, "cuTensorMapEncodeTiled (FP16 planes) failed (%d)", (int)r);
    return ESPER_OK;
}

// Power-of-two scale of the FP16 operand planes for z-scored images of P pixels: |c| <= 2 sqrt(P) (a z-score is at most
// sqrt(P - 1), so is the common image subtracted from it), scaled values stay below 2^15 < 65504, and the lo parts of
// all but the smallest pixels leave the FP16 subnormal range.  -1: P too large for the FP16 path.
int gram_f16_scale_log2(int P) {
    const double top = 2.0 * sqrt((double)P);
    int s = (int)floor(log2(32768.0 / top));
    if (s < 0) return -1;
    return s > 8 ? 8 : s;
}

// Choose the number of K chunks: chains of <= ~2048 pixels, and a unit count that fills the SMs.
static int choose_split_k(int requested, int nkb, int n_tiles, int sms) {
    if (requested > 0) return requested < nkb ? requested : nkb;
    int s = ceil_div(nkb, 64);                     // 64 k-blocks = 2048 pixels per chain
    if (s < 1) s = 1;
    if (n_tiles * s < sms) s = ceil_div(sms, n_tiles);
    // nudge s upward (at most +25%) to the value with the best last-wave fill
    int best = s; double best_eff = 0.0;
    for (int cand = s; cand <= s + s / 4 + 1; ++cand) {
        long long units = (long long)n_tiles * cand;
        long long waves = (units + sms - 1) / sms;
        double eff = (double)units / (double)(waves * sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = cand; }
    }
    if (best > nkb) best = nkb;
    return best < 1 ? 1 : best;
}

// M [na, nb] (FP64) = A B^T with A, B given as TF32 hi/lo planes of row stride ld (rows padded to 128, zero filled).
// b_hi == nullptr: symmetric product A A^T (upper-triangle tiles, mirrored).  Used by the CTF branch.
int run_product(esper_ctx* c, float* a_hi, float* a_lo, int na, float* b_hi, float* b_lo, int nb, int ld, double* M) {
    using namespace gram;
    const bool sym = (b_hi == nullptr);
    if (sym) { b_hi = a_hi; b_lo = a_lo; nb = na; }
    const int a_pad = (int)round_up((size_t)na, BM), b_pad = (int)round_up((size_t)nb, BN);
    const int nta = a_pad / BM, ntb = b_pad / BN;
    const int tiles_total = sym ? nta * (nta + 1) / 2 : nta * ntb;
    const int nkb = ld / BK;
    const int sms = c->sm_count;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_plane_map(c, &ma_hi, a_hi, a_pad, ld))) return rc;
    if ((rc = make_plane_map(c, &ma_lo, a_lo, a_pad, ld))) return rc;
    if ((rc = make_plane_map(c, &mb_hi, b_hi, b_pad, ld))) return rc;
    if ((rc = make_plane_map(c, &mb_lo, b_lo, b_pad, ld))) return rc;
    if (!c->attr_gram) {
        ESPER_CUDA(c, cudaFuncSetAttribute(gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        ESPER_CUDA(c, cudaFuncSetAttribute(gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        c->attr_gram = true;
    }
    const size_t tile_bytes = (size_t)BM * BN * sizeof(float);
    const size_t ws_cap = (size_t)3 << 30;
    int batch = tiles_total;
    int split_k = choose_split_k(c->split_k, nkb, batch, sms);
    while ((size_t)batch * split_k * tile_bytes > ws_cap && batch > sms) {
        batch = (batch + 1) / 2;
        split_k = choose_split_k(c->split_k, nkb, batch, sms);
    }
    ESPER_RESERVE(c, c->gram_part, (size_t)batch * split_k * tile_bytes);
    float* part = c->gram_part.as<float>();
    const int drain_kb = c->drain_kb > 0 ? c->drain_kb : 2;
    for (int tile0 = 0; tile0 < tiles_total; tile0 += batch) {
        const int nbt = (tiles_total - tile0 < batch) ? tiles_total - tile0 : batch;
        int sk = choose_split_k(c->split_k, nkb, nbt, sms);
        if (sk > split_k) sk = split_k;
        const int units = nbt * sk;
        const int grid = units < sms ? units : sms;
        {
            LaunchScope ls(c, "gram");
            gram_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, sym ? nta : ntb, nbt, tile0,
                                                                       sym ? -1 : FULL_RECT, nkb, sk, drain_kb, part);
        }
        {
            LaunchScope ls(c, "dist_finalize");
            product_store_kernel<<<dim3(nbt, 8), 256, 0, c->stream>>>(part, sym ? nta : ntb, nbt, tile0, sym ? 1 : 0, sk, na, nb, M);
        }
    }
    ESPER_CUDA(c, cudaGetLastError());
    return ESPER_OK;
}

int run_gram(esper_ctx* c, int n, int P, int rows_pad, int ld, unsigned flags, int row0, int nrows, double* gram_out,
             int half_scale_log2) {
    using namespace gram;
    // row0 < 0: full symmetric matrix (upper triangle of tiles); else the row block [row0, row0+nrows) x [0, n)
    const int nt = rows_pad / BM;
    const bool rect = row0 >= 0;
    const int rect_ti0 = rect ? row0 / BM : -1;
    const int tiles_total = rect ? ceil_div(nrows, BM) * nt : nt * (nt + 1) / 2;
    if (!rect) { row0 = 0; nrows = n; }
    const bool f16 = half_scale_log2 >= 0;            // FP16 planes (ld a multiple of 64 halves), see run_prep
    const int nkb = f16 ? ld / BK_H : ld / BK;
    const int sms = c->sm_count;
    const RowStat* st = c->row_stats.as<RowStat>();
    double* D = c->dist.as<double>();
    const double gscale = f16 ? ldexp(1.0, -2 * half_scale_log2) : 1.0;
    if (f16 && (gram_out || (!rect && nt < 2)))
        return fail(c, ESPER_E_STATE, "gram: the FP16 operand path covers distance matrices (full symmetric, or row blocks) only");

    CUtensorMap map_hi, map_lo;
    int rc = f16 ? make_plane_map_h(c, &map_hi, c->plane_hi.p, rows_pad, ld, BM) : make_plane_map(c, &map_hi, c->plane_hi.as<float>(), rows_pad, ld);
    if (rc) return rc;
    rc = f16 ? make_plane_map_h(c, &map_lo, c->plane_lo.p, rows_pad, ld, BM) : make_plane_map(c, &map_lo, c->plane_lo.as<float>(), rows_pad, ld);
    if (rc) return rc;

    if (!c->attr_gram) {
        ESPER_CUDA(c, cudaFuncSetAttribute(gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        ESPER_CUDA(c, cudaFuncSetAttribute(gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        c->attr_gram = true;
    }

    int drain_kb = c->drain_kb > 0 ? c->drain_kb : 2;
    if (const char* e = getenv("ESPER_DRAIN_KB")) { const int v = atoi(e); if (v != 0) drain_kb = v; }   // experiments only
    // 128x256 tiles for the full symmetric matrix (gram_wide_kernel): same bits as the 128x128 kernel, 12 % faster at the
    // benchmark shape.  ESPER_GRAM_WIDE=0 selects the 128x128 kernel (which the row-block and CTF products always use).
    const char* wide_env = getenv("ESPER_GRAM_WIDE");
    // The PCA Gram output (gram_out) keeps the 128x128 kernel until gram_finalize_kernel has run with wide tiles on hardware
    // (the wide kernel became the default in the last GPU minutes of round 1; only the distance epilogue was exercised);
    // ESPER_GRAM_WIDE=2 enables it there too.
    const bool wide_gram_ok = !gram_out || (wide_env && atoi(wide_env) == 2);
    const bool wide = (f16 && !rect) || (!rect && nt >= 2 && !(wide_env && atoi(wide_env) == 0) && drain_kb > 0 && wide_gram_ok);
    if (f16 && drain_kb < 1) drain_kb = 2;

    // Bound the split-K workspace: process the triangle in batches of tiles.
    const size_t ws_cap = (size_t)3 << 30;   // 3 GiB
    if (wide) {
        CUtensorMap mapw_hi, mapw_lo;
        if (f16) {
            if ((rc = make_plane_map_h(c, &mapw_hi, c->plane_hi.p, rows_pad, ld, BN_W))) return rc;
            if ((rc = make_plane_map_h(c, &mapw_lo, c->plane_lo.p, rows_pad, ld, BN_W))) return rc;
        } else {
            if ((rc = make_plane_map(c, &mapw_hi, c->plane_hi.as<float>(), rows_pad, ld, BN_W))) return rc;
            if ((rc = make_plane_map(c, &mapw_lo, c->plane_lo.as<float>(), rows_pad, ld, BN_W))) return rc;
        }
        if (!c->attr_gram_wide) {
            ESPER_CUDA(c, cudaFuncSetAttribute(gram_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_W_BYTES));
            ESPER_CUDA(c, cudaFuncSetAttribute(gram_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_W_BYTES));
            c->attr_gram_wide = true;
        }
        const int n_wide = wide_tile_count(nt);
        const size_t wtile_bytes = (size_t)2 * BM * BN * sizeof(float);
        int batch = n_wide;
        int split_k = choose_split_k(c->split_k, nkb, batch, sms);
        while ((size_t)batch * split_k * wtile_bytes > ws_cap && batch > sms) {
            batch = (batch + 1) / 2;
            split_k = choose_split_k(c->split_k, nkb, batch, sms);
        }
        ESPER_RESERVE(c, c->gram_part, (size_t)batch * split_k * wtile_bytes);
        float* part = c->gram_part.as<float>();
        for (int tile0 = 0; tile0 < n_wide; tile0 += batch) {
            const int nb = (n_wide - tile0 < batch) ? n_wide - tile0 : batch;
            int sk = choose_split_k(c->split_k, nkb, nb, sms);
            if (sk > split_k) sk = split_k;
            const int units = nb * sk;
            {
                LaunchScope ls(c, "gram");
                if (f16)
                    gram_wide_kernel<true><<<units < sms ? units : sms, NUM_THREADS, SMEM_W_BYTES, c->stream>>>(map_hi, map_lo, mapw_hi, mapw_lo, nt, nb,
                                                                                                              tile0, nkb, sk, drain_kb, part);
                else
                    gram_wide_kernel<false><<<units < sms ? units : sms, NUM_THREADS, SMEM_W_BYTES, c->stream>>>(map_hi, map_lo, mapw_hi, mapw_lo, nt, nb,
                                                                                                               tile0, nkb, sk, drain_kb, part);
            }
            {
                LaunchScope ls(c, "dist_finalize");
                if (gram_out) gram_finalize_kernel<<<dim3(2 * nb, 8), 256, 0, c->stream>>>(part, nt, 2 * nb, 2 * tile0, WIDE_TRI, sk, n, st, gram_out);
                else dist_finalize_kernel<<<dim3(2 * nb, 8), 256, 0, c->stream>>>(part, nt, 2 * nb, 2 * tile0, WIDE_TRI, sk, n, P, st, gscale, D);
            }
        }
        ESPER_CUDA(c, cudaGetLastError());
    } else {
    const size_t tile_bytes = (size_t)BM * BN * sizeof(float);
    int batch = tiles_total;
    int split_k = choose_split_k(c->split_k, nkb, batch, sms);
    while ((size_t)batch * split_k * tile_bytes > ws_cap && batch > sms) {
        batch = (batch + 1) / 2;
        split_k = choose_split_k(c->split_k, nkb, batch, sms);
    }
    ESPER_RESERVE(c, c->gram_part, (size_t)batch * split_k * tile_bytes);
    float* part = c->gram_part.as<float>();

    for (int tile0 = 0; tile0 < tiles_total; tile0 += batch) {
        const int nb = (tiles_total - tile0 < batch) ? tiles_total - tile0 : batch;
        int sk = choose_split_k(c->split_k, nkb, nb, sms);
        if (sk > split_k) sk = split_k;
        const int units = nb * sk;
        const int grid = units < sms ? units : sms;
        {
            LaunchScope ls(c, "gram");
            if (f16) gram_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(map_hi, map_lo, map_hi, map_lo, nt, nb, tile0, rect_ti0, nkb, sk, drain_kb, part);
            else gram_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(map_hi, map_lo, map_hi, map_lo, nt, nb, tile0, rect_ti0, nkb, sk, drain_kb, part);
        }
        {
            LaunchScope ls(c, "dist_finalize");
            if (gram_out) gram_finalize_kernel<<<dim3(nb, 8), 256, 0, c->stream>>>(part, nt, nb, tile0, -1, sk, n, st, gram_out);
            else dist_finalize_kernel<<<dim3(nb, 8), 256, 0, c->stream>>>(part, nt, nb, tile0, rect_ti0, sk, n, P, st, gscale, D);
        }
    }
    ESPER_CUDA(c, cudaGetLastError());
    }

    if (!gram_out && (flags & ESPER_DIST_REFINE) && c->refine_tau > 0.0 && n > 1) {
        const size_t max_pairs = rect ? (size_t)nrows * n : (size_t)n * (n - 1) / 2;
        const unsigned int cap = (unsigned int)(max_pairs < ((size_t)1 << 26) ? max_pairs : (size_t)1 << 26);
        ESPER_RESERVE(c, c->pair_list, sizeof(int2) * (size_t)cap + 256);
        unsigned int* count = c->pair_list.as<unsigned int>();
        int2* pairs = reinterpret_cast<int2*>(c->pair_list.as<char>() + 256);
        ESPER_CUDA(c, cudaMemsetAsync(count, 0, sizeof(unsigned int), c->stream));
        {
            LaunchScope ls(c, "refine");
            const long long tot = (long long)nrows * n;
            // Small stacks (the reference's own CPU-runnable inputs: 400 pristine states, config 1) are noise-free, every pair is
            // "close" and the 1e-5 error budget of D is amplified ~20x by exp(-D^2 / 2 eps) downstream: there all pairs are
            // recomputed in the reference's arithmetic (|c_i - c_j|^2 < 2 (|c_i|^2 + |c_j|^2) always holds).
            const double tau = (n <= ESPER_EXACT_PAIRS_MAX_N) ? fmax(c->refine_tau, 2.0) : c->refine_tau;
            flag_pairs_kernel<<<ceil_div(tot, 256), 256, 0, c->stream>>>(D, n, row0, nrows, rect ? 1 : 0, P, st, tau, pairs, count, cap);
        }
        {
            LaunchScope ls(c, "refine");
            refine_pairs_kernel<<<sms * 8, 256, 0, c->stream>>>(c->imgs.as<float>(), n, row0, rect ? 1 : 0, P, st, pairs, count, cap, D);
        }
        ESPER_CUDA(c, cudaGetLastError());
    }
    return ESPER_OK;
}

}  // namespace esper
