// W4A16 "skinny" GEMM for 9 <= M <= 128 activation rows (batched decode steps, 128-token prefill
// chunks): swap-AB tcgen05 with split reduction, so that EVERY SM streams weights.
//
//   out[m, n] = sum_k a[m, k] * T(code[n, k] * scale[n, k/128] + bias[n, k/128])       (m = token, n = feature)
//
// Why (round-1 VERDICT, missing #5 / row N4): with M <= 128 the 128 x 128-tile GEMM has K_out/128
// CTAs (20 for the o / down projections of Qwen3-4B) and the streaming kernel re-reads the weights
// once per 32 rows; a 64-slot decode step took 16 ms where the weights are worth 0.33 ms of HBM time.
// Reference's answer on its hardware: quantized_matmul_splitk (quantized_matmul.metal:251-293, policy
// quantized_matmul.cpp:136-150).  Here:
//
//   * the WEIGHTS are the 128-row M operand of the MMA (one CTA tile = 128 output features), the tokens
//     are the N operand (16..128 columns): D[feature, token] lives in TMEM, lane = feature;
//   * the reduction is split over `splits` CTAs per feature tile so that tiles x splits ~ #SMs; partial
//     sums go to a workspace in fp32 and a small second launch adds the planes in split order (deterministic)
//     and runs the epilogue (plain / + residual / SwiGLU of interleaved gate|up rows);
//   * packed weights arrive by TMA (one 128-row x 128-byte box = two quantisation groups of 128 reduction elements per
//     copy, three or four boxes in flight); 256 dequantiser threads in two groups of four warps - group g owns the
//     128-wide reduction blocks g, g + 2, ... - turn them into bf16 (LOP3 magic -> exact code -> one HFMA2, the rounding
//     point of the reference's tiled kernel, quantized_matmul.metal:183-194) and write them to TENSOR memory
//     (tcgen05.st, lane = feature row): the MMA reads its A operand from there; activations arrive by TMA (tokens
//     beyond M zero-filled); the MMA warp issues tcgen05.mma M128 N{16..128} K16 from an elected lane.
// Two CTAs fit per SM (96-113 KB of shared memory, 256 TMEM columns each) at every column count.
//
// How it got here (lm_head at 64 rows, 2560 -> 151936, 218 MB: profiles/r02_skinny_history.md): 256 threads on the SAME
// 64-wide block with the tile in shared memory 156 us -> two alternating groups, thread = row 143 us -> tile in tensor
// memory (6 stages instead of 2) 136 us -> MMA warp on elect.sync 112 us -> 128-wide blocks (one barrier round trip per
// quantisation group) 94 us (36.5 % of the HBM rate; 16 rows: 79 us, 41 %).  The per-block cycle trace (tools/skinny_blocks.py) showed each time which actor the
// others were waiting for; the arithmetic of the dequantisers (~2.9 instructions per weight) is the floor.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"
#include "trace.cuh"

namespace tl {

constexpr int SK_FEAT = 128;      // features per CTA tile (UMMA M)
constexpr int SK_KB = 64;         // reduction elements per activation tile / MMA descriptor (one 128-byte swizzle atom)
constexpr int SK_GB = 128;        // reduction elements per pipeline step: one quantisation group, two activation tiles, 8 MMAs
constexpr int SK_PG = 2;          // group blocks per packed-weight TMA box: 128 rows x 128 B (a 32-byte-wide box cost ~3 us per block)
constexpr int SK_PACKED_BYTES = SK_PG * SK_FEAT * SK_GB / 2;  // 16 KiB per box, 128-byte swizzle
constexpr int SK_DEQ_THREADS = 256;
constexpr int SK_TRC_BLOCKS = 40;
constexpr int SK_GROUP_WARPS = 4;   // dequantiser warps per group block (two groups alternate blocks)
constexpr int SK_THREADS = 128 + SK_DEQ_THREADS;
enum { SK_EPI_NONE = 0, SK_EPI_RESIDUAL = 1, SK_EPI_SWIGLU_PAIRS = 2 };

// Shared memory: activation ring (BSTAGES group blocks of two [NT x 64] tiles) + packed-weight ring (PSTAGES boxes) +
// barriers.  Tensor memory: D[feature, token] in the first D_COLS columns, then ASTAGES dequantised tiles of 64
// columns (128 reduction elements as bf16 pairs, lane = feature row).
// WIDE1 (NT = 128 only): the one-CTA-per-SM geometry of 128 token columns (deeper rings, six tile stages in all 512
// TMEM columns); the default packs two CTAs per SM there too (two activation stages, two tile stages): the dequantiser
// throughput of an SM is what bounds this kernel, and one CTA has half the warps.
template <int NT, bool WIDE1 = false>
struct SkSmem {
    static constexpr int B_BYTES = NT * SK_KB * 2;                 // one [NT x 64] activation tile
    static constexpr int BSTAGES = NT >= 128 ? (WIDE1 ? 3 : 2) : (NT >= 64 ? 3 : 4);  // group blocks in flight (L2 round trip ~1 us)
    static constexpr int PSTAGES = NT >= 128 && WIDE1 ? 4 : 3;     // packed boxes in flight (two group blocks each)
    static constexpr int B_OFF = 0;
    static constexpr int P_OFF = B_OFF + BSTAGES * 2 * B_BYTES;
    static constexpr int BAR_OFF = P_OFF + PSTAGES * SK_PACKED_BYTES;
    static constexpr int BYTES = BAR_OFF + 512;
    static constexpr int D_COLS = NT < 32 ? 32 : NT;               // accumulator columns (fp32)
    static constexpr int TMEM_COLS = NT >= 128 && WIDE1 ? 512 : 256;  // two CTAs per SM share the 512 columns
    static constexpr int CTAS_PER_SM = NT >= 128 && WIDE1 ? 1 : 2;
    static constexpr int ASTAGES = (TMEM_COLS - D_COLS) / 64;      // 3 (NT <= 64) or 6 (NT = 128)
    static_assert(BYTES <= 227 * 1024 && ASTAGES >= 2, "budget");
};

struct SkArgs {
    const void *scales, *biases, *residual;
    void *out;
    float *partials;   // [splits][M][K] fp32 (splits > 1)
    int M, N, K;       // tokens, reduction, features
    int splits, gb_per_split;  // reduction split in units of 128-wide group blocks
    int epilogue;
};

template <typename T>
struct SkNum;
template <>
struct SkNum<__nv_bfloat16> {
    using V2 = __nv_bfloat162;
    static constexpr uint32_t MAGIC = 0x43004300u;
    static constexpr uint32_t FMT = 1u;
};
template <>
struct SkNum<__half> {
    using V2 = __half2;
    static constexpr uint32_t MAGIC = 0x64006400u;
    static constexpr uint32_t FMT = 0u;
};

template <typename T, int NT>
__host__ __device__ constexpr uint32_t sk_instr_desc() {
    return (1u << 4) | (SkNum<T>::FMT << 7) | (SkNum<T>::FMT << 10) | (0u << 15) | (0u << 16) | (static_cast<uint32_t>(NT >> 3) << 17) |
           (static_cast<uint32_t>(SK_FEAT >> 4) << 24);
}

template <typename T, int NT, bool WIDE1>
__global__ void __launch_bounds__(SK_THREADS, SkSmem<NT, WIDE1>::CTAS_PER_SM)
w4a16_skinny_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const SkArgs args) {
    using Smem = SkSmem<NT, WIDE1>;
    constexpr int ASTAGES = Smem::ASTAGES;
    constexpr int PSTAGES = Smem::PSTAGES;
    constexpr int BSTAGES = Smem::BSTAGES;
    extern __shared__ __align__(1024) unsigned char ssm[];
#if TL_TRACE
    // per-block cycle stamps of CTA 0 (tools/skinny_blocks.py): role 0 = MMA warp (k: A tile ready, B tiles ready, MMAs
    // issued), roles 1 / 2 = first thread of dequantiser group 0 / 1 (k: box ready, math done, stage free, handed over)
    __shared__ unsigned long long trc[3][SK_TRC_BLOCKS][4];
#define SK_TRC(role, i, k)                                                                                      \
    do {                                                                                                        \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (i) < SK_TRC_BLOCKS) trc[role][i][k] = clock64();     \
    } while (0)
#else
#define SK_TRC(role, i, k) do { } while (0)
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / args.splits, split = blockIdx.x - tile * args.splits;
    const int G = args.N / SK_GB;  // group blocks of the whole reduction = quantisation groups per row
    const int gb0 = min(split * args.gb_per_split, G), gb1 = min(gb0 + args.gb_per_split, G);
    const int n_gb = gb1 - gb0;
    TL_TRACE_STAMP(30);
    // Programmatic dependent launch: the next kernel of the stream may become resident now (its own prologue and weight
    // pipeline do not depend on this grid).  Of THIS kernel only the activation loads and the epilogue depend on the
    // predecessor: barrier init, TMEM allocation, the packed-weight TMA ring and the dequantisers run ahead of
    // griddep_wait(), i.e. under the predecessor's tail.
    griddep_launch();

    const uint32_t b_base = g_smem_u32(ssm + Smem::B_OFF), p_base = g_smem_u32(ssm + Smem::P_OFF);
    const uint32_t bar = g_smem_u32(ssm + Smem::BAR_OFF);
    const uint32_t full_a = bar, empty = bar + 8 * ASTAGES, p_full = bar + 16 * ASTAGES, p_empty = p_full + 8 * PSTAGES;
    const uint32_t full_b = p_empty + 8 * PSTAGES, b_empty = full_b + 8 * BSTAGES, tmem_full = b_empty + 8 * BSTAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ssm + Smem::BAR_OFF + 16 * ASTAGES + 16 * PSTAGES + 16 * BSTAGES + 8);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < ASTAGES; ++i) {
            g_mbar_init(full_a + 8 * i, SK_GROUP_WARPS);  // one arrival per warp of the dequantiser group that owns the block
            g_mbar_init(empty + 8 * i, 1);
        }
        for (int i = 0; i < BSTAGES; ++i) {
            g_mbar_init(full_b + 8 * i, 1);
            g_mbar_init(b_empty + 8 * i, 1);
        }
        for (int i = 0; i < PSTAGES; ++i) {
            g_mbar_init(p_full + 8 * i, 1);
            g_mbar_init(p_empty + 8 * i, SK_DEQ_THREADS / 32);
        }
        g_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "n"(Smem::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    g_tc_fence_before();
    __syncthreads();
    g_tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    TL_TRACE_STAMP(31);

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: packed weights, one 128-row x 128-byte box per two group blocks
        const int n_box = (n_gb + SK_PG - 1) / SK_PG;
        int ps = 0;
        uint32_t ph = 1;  // producer side: the first pass through the ring does not wait
        for (int i = 0; i < n_box; ++i) {
            g_mbar_wait(p_empty + 8 * ps, ph);
            if (g_elect_one()) {
                g_mbar_expect_tx(p_full + 8 * ps, SK_PACKED_BYTES);  // columns past the row end are zero-filled and still counted
                g_tma_load_2d(p_base + ps * SK_PACKED_BYTES, &tmap_w, (gb0 + i * SK_PG) * (SK_GB / 2), tile * SK_FEAT, p_full + 8 * ps);
            }
            __syncwarp();
            if (++ps == PSTAGES) ps = 0, ph ^= 1u;
        }
    } else if (warp == 3) {
        // ------------------------------------------------ TMA producer: activations, two [NT x 64] tiles per group block
        griddep_wait();  // the activations are the predecessor's output (every lane: whichever one is elected below has waited)
        int bs = 0;
        uint32_t ph = 1;
        for (int i = 0; i < n_gb; ++i) {
            g_mbar_wait(b_empty + 8 * bs, ph);
            if (g_elect_one()) {
                g_mbar_expect_tx(full_b + 8 * bs, 2 * Smem::B_BYTES);
                g_tma_load_2d(b_base + (2 * bs) * Smem::B_BYTES, &tmap_a, (gb0 + i) * SK_GB, 0, full_b + 8 * bs);
                g_tma_load_2d(b_base + (2 * bs + 1) * Smem::B_BYTES, &tmap_a, (gb0 + i) * SK_GB + SK_KB, 0, full_b + 8 * bs);
            }
            __syncwarp();
            if (++bs == BSTAGES) bs = 0, ph ^= 1u;
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer: D[feature, token] += W_block . A_block^T, 8 K-steps per group block
        // The whole warp walks the loop (uniform control flow: barrier addresses, descriptors and phases stay in
        // uniform registers, no division per block), one elected lane issues.
        constexpr uint32_t idesc = sk_instr_desc<T, NT>();
        const uint64_t bdesc0 = g_smem_desc_sw128(b_base, 0, 1024);
        int s = 0, bs = 0;
        uint32_t pha = 0, phb = 0;
        for (int i = 0; i < n_gb; ++i) {
            g_mbar_wait(full_a + 8 * s, pha);
            SK_TRC(0, i, 0);
            g_mbar_wait(full_b + 8 * bs, phb);
            SK_TRC(0, i, 1);
            g_tc_fence_after();
            if (g_elect_one()) {
                const uint32_t a_tmem = tmem_d + Smem::D_COLS + s * 64;  // 16 reduction elements = 8 columns per K step
                const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(bs * (2 * Smem::B_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < SK_GB / 16; ++k)
                    g_tc_mma_ts(tmem_d, a_tmem + 8 * k, bdesc + (k >> 2) * (Smem::B_BYTES >> 4) + 2 * (k & 3), idesc, (i > 0 || k > 0) ? 1u : 0u);
                g_tc_commit(empty + 8 * s);
                g_tc_commit(b_empty + 8 * bs);
            }
            __syncwarp();
            SK_TRC(0, i, 2);
            if (++s == ASTAGES) s = 0, pha ^= 1u;
            if (++bs == BSTAGES) bs = 0, phb ^= 1u;
        }
        if (g_elect_one()) g_tc_commit(tmem_full);
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------ dequantisers
        // Two groups of four warps; group g owns the group blocks i = g, g + 2, ... = bytes [64 g, 64 g + 64) of its row
        // in every packed box, so the two groups work on different blocks and meet only in the MMA queue.  Thread = one
        // feature row of the tile = the TMEM lane it reads in the epilogue.  Per block: four 16-byte chunks of the box
        // row -> 128 bf16 (scale and bias of the row's quantisation group: one pair per block, read from global memory
        // one own block ahead) -> two tcgen05.st of 32 columns -> ONE wait / fence / barrier arrival.
        const int grp = (warp - 4) >> 2;
        const int row = (warp & 3) * 32 + lane;
        using V2 = typename SkNum<T>::V2;
        const uint32_t magic = SkNum<T>::MAGIC;
        const V2 offset2 = *reinterpret_cast<const V2 *>(&magic);
        const size_t srow = static_cast<size_t>(min(tile * SK_FEAT + row, args.K - 1)) * G;
        const unsigned short *sc = reinterpret_cast<const unsigned short *>(args.scales) + srow + gb0;
        const unsigned short *bi = reinterpret_cast<const unsigned short *>(args.biases) + srow + gb0;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const unsigned char *p_row = ssm + Smem::P_OFF + row * 128;
        const uint32_t lane_base = tmem_d + (static_cast<uint32_t>((warp & 3) * 32) << 16) + Smem::D_COLS;
        unsigned short s_next = 0, b_next = 0;
        if (grp < n_gb) s_next = __ldg(sc + grp), b_next = __ldg(bi + grp);
        int ps = 0, s = grp % ASTAGES;
        uint32_t php = 0, phe = 1;  // consumer of the packed ring; producer side of the tile ring (first pass does not wait)
        if (grp >= ASTAGES) phe ^= 1u;
        for (int i = grp; i < n_gb; i += 2) {
            const unsigned short s16 = s_next, b16 = b_next;
            if (i + 2 < n_gb) s_next = __ldg(sc + i + 2), b_next = __ldg(bi + i + 2);
            g_mbar_wait(p_full + 8 * ps, php);
            if (i == grp) TL_TRACE_STAMP_T(32, 128);  // first packed box has landed
            if ((warp & 3) == 0) SK_TRC(1 + grp, i, 0);
            V2 s2, b2;
            s2.x = s2.y = *reinterpret_cast<const T *>(&s16);
            b2.x = b2.y = *reinterpret_cast<const T *>(&b16);
            // row `row` of the box: 128 bytes = 8 chunks of 16 B, chunk c stored at (c ^ (row & 7)) by the TMA swizzle
            const unsigned char *src = p_row + ps * SK_PACKED_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // the two 64-wide halves of the block: 32 tensor-memory columns each
                const uint4 lo = *reinterpret_cast<const uint4 *>(src + ((static_cast<uint32_t>(4 * grp + 2 * h) ^ swz) << 4));
                const uint4 hi = *reinterpret_cast<const uint4 *>(src + ((static_cast<uint32_t>(4 * grp + 2 * h + 1) ^ swz) << 4));
                uint32_t outw[32];
                const uint32_t wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t p[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t bits;  // (128 + code_q, 128 + code_{q+4}) in one LOP3: ((w >> 4q) & 0x000F000F) | magic
                        asm("lop3.b32 %0, %1, 0x000F000F, %2, 0xEA;" : "=r"(bits) : "r"(wv[j] >> (4 * q)), "r"(magic));
                        V2 code = __hsub2(*reinterpret_cast<V2 *>(&bits), offset2);
                        V2 v = __hfma2(code, s2, b2);                                   // code * scale + bias, one rounding
                        p[q] = *reinterpret_cast<uint32_t *>(&v);
                    }
                    // column c of the tile = elements (2c, 2c + 1) of the block
                    outw[4 * j + 0] = __byte_perm(p[0], p[1], 0x5410);
                    outw[4 * j + 1] = __byte_perm(p[2], p[3], 0x5410);
                    outw[4 * j + 2] = __byte_perm(p[0], p[1], 0x7632);
                    outw[4 * j + 3] = __byte_perm(p[2], p[3], 0x7632);
                }
                if (h == 0) {
                    if ((warp & 3) == 0) SK_TRC(1 + grp, i, 1);
                    g_mbar_wait(empty + 8 * s, phe);
                    if ((warp & 3) == 0) SK_TRC(1 + grp, i, 2);
                    g_tc_fence_after();
                }
                g_tmem_st32(lane_base + s * 64 + h * 32, outw);
            }
            g_tmem_st_wait();
            g_tc_fence_before();
            // one arrival per WARP (an arrival per thread is a serialised shared-memory atomic each: ~0.4 us per block in
            // the first version); __syncwarp orders the lanes' stores before it
            __syncwarp();
            if (lane == 0) {
                g_mbar_arrive(full_a + 8 * s);
                // Release the packed box only now: the tensor-memory stores above consumed the loaded words, so this
                // warp's shared-memory reads of the box have COMPLETED (an arrive issued right behind the load let the TMA
                // refill the box under a load still in flight - mbarrier ops are not ordered behind the load/store
                // unit - and single feature rows came out wrong in ~1 of 600 CTA-loops of 40 blocks).
                g_mbar_arrive(p_empty + 8 * ps);
            }
            if ((warp & 3) == 0) SK_TRC(1 + grp, i, 3);
            if (++ps == PSTAGES) ps = 0, php ^= 1u;
            s += 2;
            if (s >= ASTAGES) s -= ASTAGES, phe ^= 1u;
        }
        // ------------------------------------------------ epilogue: TMEM lane = feature; warps 4-7 take token columns [0, NT/2), 8-11 the rest
        TL_TRACE_STAMP_T(33, 128);  // last weight tile handed to the MMA thread
        // the epilogue reads the residual and overwrites buffers (output, partial planes) that the predecessor - the
        // reduction kernel of the previous projection - may still be reading
        griddep_wait();
        if (n_gb > 0) {
            g_mbar_wait(tmem_full, 0);
            g_tc_fence_after();
        }
        TL_TRACE_STAMP_T(34, 128);  // accumulators complete
        const int q = warp & 3;
        const int f = q * 32 + lane;            // feature row inside the tile
        const int n = tile * SK_FEAT + f;       // global feature
        constexpr int HALF_COLS = NT >= 64 ? NT / 2 : NT;   // with fewer than 64 tokens warps 8-11 have nothing to read
        const int col0 = ((warp - 4) >> 2) * HALF_COLS;
        const bool reader = NT >= 64 || warp < 8;
        constexpr int CH = HALF_COLS < 32 ? HALF_COLS : 32;  // columns per tcgen05.ld (x16 or x32)
        float *part = args.partials;
        const size_t plane = static_cast<size_t>(args.M) * args.K;
        // ---- splits > 1: park the fp32 partial tile; w4a16_skinny_reduce_kernel adds the planes in split order
        if (args.splits > 1) {
            if (reader) {
#pragma unroll
                for (int c0 = 0; c0 < HALF_COLS; c0 += 32) {
                    uint32_t v[32];
                    if (n_gb > 0) {
                        g_tmem_ld32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + col0 + c0, v);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = 0u;
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int m = col0 + c0 + c;
                        if (c < CH && m < args.M && n < args.K) part[split * plane + static_cast<size_t>(m) * args.K + n] = __uint_as_float(v[c]);
                    }
                }
            }
        }
        if (args.splits == 1 && reader) {
            T *out = static_cast<T *>(args.out);
            const T *res = static_cast<const T *>(args.residual);
#pragma unroll
            for (int c0 = 0; c0 < HALF_COLS; c0 += 32) {
                float acc[32];
                {
                    uint32_t v[32];
                    if (n_gb > 0) {
                        g_tmem_ld32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + col0 + c0, v);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = 0u;
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[c] = __uint_as_float(v[c]);
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int m = col0 + c0 + c;
                    const bool live = c < CH && m < args.M;  // warp-uniform in m; the shuffle below needs every lane
                    if (args.epilogue == SK_EPI_SWIGLU_PAIRS) {
                        // rows 16j + r (gate) and 16j + 8 + r (up) of the interleaved weight -> activation 8j + r
                        const float mine = to_f(from_f<T>(acc[c]));
                        const float other = __shfl_xor_sync(0xffffffffu, mine, 8);
                        if (live && (f & 8) == 0 && n < args.K) {
                            const int feat = (n >> 4) * 8 + (n & 7);
                            out[static_cast<size_t>(m) * (args.K / 2) + feat] = from_f<T>((mine / (1.0f + expf(-mine))) * other);
                        }
                    } else if (live && n < args.K) {
                        T vb = from_f<T>(acc[c]);
                        if (args.epilogue == SK_EPI_RESIDUAL) vb = from_f<T>(to_f(ld_cg(res + static_cast<size_t>(m) * args.K + n)) + to_f(vb));
                        out[static_cast<size_t>(m) * args.K + n] = vb;
                    }
                }
            }
        }
    }
    TL_TRACE_STAMP_T(35, 128);  // epilogue stored
    g_tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        g_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(Smem::TMEM_COLS) : "memory");
    }
    TL_TRACE_STAMP(36);
#if TL_TRACE
    if (blockIdx.x == 0 && g_trace_buf != nullptr) {  // dump the per-block stamps: tag = 10000 + role * 1000 + block * 4 + k
        for (int e = threadIdx.x; e < 3 * SK_TRC_BLOCKS * 4; e += SK_THREADS) {
            const int role = e / (SK_TRC_BLOCKS * 4), rest = e - role * SK_TRC_BLOCKS * 4;
            if (rest / 4 < n_gb && (role == 0 || (rest / 4) % 2 == role - 1)) {
                const unsigned at = atomicAdd(g_trace_n, 1u);
                if (at < g_trace_cap) g_trace_buf[2 * at] = 10000 + role * 1000 + rest, g_trace_buf[2 * at + 1] = trc[role][rest / 4][rest & 3];
            }
        }
    }
#endif
}

#if TL_TRACE
void trace_bind_skinny(unsigned long long *buf, unsigned int *n, unsigned int cap) { trace_bind(buf, n, cap); }
#endif

// Adds the fp32 partial planes of a split reduction in split order (deterministic: same bits on every run and for
// every CTA schedule) and applies the epilogue; one thread per output.  All `splits` loads of a thread are independent
// and in flight together (the first version let the last CTA of a tile do this with dependent round trips: 40 us).
template <typename T>
__global__ void __launch_bounds__(256) w4a16_skinny_reduce_kernel(const float *part, const T *res, T *out, int M, int K, int splits, int epilogue) {
    const size_t plane = static_cast<size_t>(M) * K;
    griddep_launch();
    griddep_wait();  // the planes are the GEMM's output; they are rewritten by every projection, so read them through L2
    // four planes per round trip, added in split order
    auto plane_sum = [&](size_t at) {
        float sum = 0.f;
        for (int sp0 = 0; sp0 < splits; sp0 += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = sp0 + j < splits ? ld_cg(part + (sp0 + j) * plane + at) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (sp0 + j < splits) sum += v[j];
        }
        return sum;
    };
    if (epilogue == SK_EPI_SWIGLU_PAIRS) {
        const int half = K / 2;
        const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (idx >= static_cast<size_t>(M) * half) return;
        const int m = static_cast<int>(idx / half), feat = static_cast<int>(idx - static_cast<size_t>(m) * half);
        const int n_gate = (feat >> 3) * 16 + (feat & 7);  // rows 16j + r (gate) and 16j + 8 + r (up) -> activation 8j + r
        const float g = plane_sum(static_cast<size_t>(m) * K + n_gate), u = plane_sum(static_cast<size_t>(m) * K + n_gate + 8);
        const float gate = to_f(from_f<T>(g)), up = to_f(from_f<T>(u));
        out[idx] = from_f<T>((gate / (1.0f + expf(-gate))) * up);
        return;
    }
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= plane) return;
    T vb = from_f<T>(plane_sum(idx));
    if (epilogue == SK_EPI_RESIDUAL) vb = from_f<T>(to_f(ld_cg(res + idx)) + to_f(vb));
    out[idx] = vb;
}

// The same reduction for the residual epilogue when the RMSNorm of the NEXT projection follows (o -> post-attention norm,
// down -> next layer's input norm): one CTA per token row adds the planes, adds the residual, writes the residual
// stream and - the row being complete in its registers - the normalised row too (week2_kernels.metal:41-47 arithmetic:
// T(x * rsqrt(mean(x^2) + eps) * w) on the ROUNDED residual sum).  Saves the rms_norm launch and its read of the row.
constexpr int SK_NORM_PER = 16;  // elements per thread: K <= 4096
template <typename T>
__global__ void __launch_bounds__(256) w4a16_skinny_reduce_norm_kernel(const float *part, const T *res, const T *__restrict__ norm_w, T *out, T *normed,
                                                                       int M, int K, int splits, float eps) {
    __shared__ float warp_part[8];
    griddep_launch();
    griddep_wait();
    const size_t plane = static_cast<size_t>(M) * K, row = static_cast<size_t>(blockIdx.x) * K;
    // all loads of a round (four planes x the thread's elements, then the residual) are issued before the first is used:
    // ld_cg is a volatile asm, so a load-then-add loop per element would pay one L2 round trip per batch (measured: 14 us
    // for this kernel against 3 + 2 us for the two it replaces)
    float xs[SK_NORM_PER];
#pragma unroll
    for (int j = 0; j < SK_NORM_PER; ++j) xs[j] = 0.f;
    for (int sp0 = 0; sp0 < splits; sp0 += 4) {
        float v[SK_NORM_PER][4];
#pragma unroll
        for (int j = 0; j < SK_NORM_PER; ++j) {
            const int n = threadIdx.x + j * 256;
#pragma unroll
            for (int q = 0; q < 4; ++q) v[j][q] = (n < K && sp0 + q < splits) ? ld_cg(part + (sp0 + q) * plane + row + n) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < SK_NORM_PER; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (sp0 + q < splits) xs[j] += v[j][q];  // split order
        }
    }
    T rs[SK_NORM_PER];
#pragma unroll
    for (int j = 0; j < SK_NORM_PER; ++j) {
        const int n = threadIdx.x + j * 256;
        rs[j] = n < K ? ld_cg(res + row + n) : from_f<T>(0.f);
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < SK_NORM_PER; ++j) {
        const int n = threadIdx.x + j * 256;
        if (n < K) {
            const T vb = from_f<T>(to_f(rs[j]) + to_f(from_f<T>(xs[j])));
            out[row + n] = vb;
            xs[j] = to_f(vb);
            ss += xs[j] * xs[j];
        }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = ss;
    __syncthreads();
    ss = warp_sum((threadIdx.x & 31) < 8 ? warp_part[threadIdx.x & 31] : 0.f);
    const float inv = rsqrtf(ss / static_cast<float>(K) + eps);
#pragma unroll
    for (int j = 0; j < SK_NORM_PER; ++j) {
        const int n = threadIdx.x + j * 256;
        if (n < K) normed[row + n] = from_f<T>(xs[j] * inv * to_f(norm_w[n]));
    }
}

// ---------------------------------------------------------------- host side --
bool w4a16_skinny_supported(int M, int N, int K, int dtype) {
    static const bool off = [] { const char *e = getenv("TL_SKINNY"); return e != nullptr && e[0] == '0'; }();
    static const int min_rows = [] { const char *e = getenv("TL_SKINNY_MIN_ROWS"); return e ? atoi(e) : 9; }();
    if (off) return false;
    return (dtype == TL_BF16 || dtype == TL_F16) && M >= min_rows && M <= 128 && K > 0 && N % 128 == 0;
}

// Split policy (ours; the reference's constants are M4-Pro tuning, quantized_matmul.cpp:138-150): the split count that
// minimises waves x (group blocks per CTA + fixed cost) + reduce launch, with `slots` CTAs resident at once (two per
// SM), a fixed cost per CTA worth ~10 group blocks (TMEM allocation, first TMA
// round trips, epilogue: ~3 us against ~0.3 us per 128-wide block) and ~8 for the extra reduce launch.
static bool skinny_wide1() {  // TL_SKINNY_WIDE1=1: one CTA per SM at 65..128 rows (A/B runs)
    static const bool on = [] { const char *e = getenv("TL_SKINNY_WIDE1"); return e != nullptr && e[0] == '1'; }();
    return on;
}
static int skinny_slots(int M) { return (M <= 64 || !skinny_wide1() ? 2 : 1) * sm_count(); }
int w4a16_skinny_splits(int M, int N, int K) {
    const int tiles = (K + SK_FEAT - 1) / SK_FEAT;
    const int num_gb = N / SK_GB;
    const int slots = skinny_slots(M);
    int best = 1;
    long long best_cost = -1;
    // (a sweep of the cap (4, 8, 16) and of the fixed cost (4, 10, 20) moved single shapes by +-10 % stand-alone and the
    // 64-slot step by nothing or for the worse: tools/gpu_call27.sh)
    for (int s = 1; s <= 16 && s <= num_gb / 2 + (num_gb < 2); ++s) {
        const int gbps = (num_gb + s - 1) / s;
        const int real = (num_gb + gbps - 1) / gbps;  // splits that actually get blocks
        const long long waves = (static_cast<long long>(tiles) * real + slots - 1) / slots;
        const long long cost = waves * (gbps + 10) + (real > 1 ? 8 : 0);  // in units of one group block
        if (best_cost < 0 || cost < best_cost) best_cost = cost, best = real;
    }
    return best;
}

size_t w4a16_skinny_workspace(int M, int N, int K) {
    const int splits = w4a16_skinny_splits(M, N, K);
    return splits > 1 ? static_cast<size_t>(splits) * M * K * sizeof(float) : 0;
}

struct SkMapKey {
    const void *ptr;
    unsigned long long d0, d1;
    unsigned b0, b1;
    int kind;
    bool operator==(const SkMapKey &o) const { return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && b0 == o.b0 && b1 == o.b1 && kind == o.kind; }
};
struct SkMapKeyHash {
    size_t operator()(const SkMapKey &k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        for (unsigned long long v : {k.d0, k.d1, static_cast<unsigned long long>(k.b0), static_cast<unsigned long long>(k.b1), static_cast<unsigned long long>(k.kind)})
            h = h * 1000003u ^ static_cast<size_t>(v);
        return h;
    }
};
// 2-D tensor maps cached per (pointer, shape): kind 0 = 16-bit activations [rows, cols] with the 128-byte swizzle,
// kind 1 = packed weights as bytes [rows, cols/2], 128-byte boxes with the 128-byte swizzle.
static int sk_cached_map(CUtensorMap *out, const void *ptr, int kind, cuuint64_t cols, cuuint64_t rows, cuuint32_t box_cols, cuuint32_t box_rows,
                         CUtensorMapDataType dt, size_t elem) {
    static std::mutex mu;
    static std::unordered_map<SkMapKey, CUtensorMap, SkMapKeyHash> cache;
    SkMapKey key{ptr, cols, rows, box_cols, box_rows, kind * 4 + static_cast<int>(dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT16)};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return TL_OK;
    }
    PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
    if (encode == nullptr) return fail(TL_ECUDA, "quantized_matmul: cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * elem};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap map;
    CUresult r = encode(&map, dt, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TL_ECUDA, "quantized_matmul: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, map);
    *out = map;
    return TL_OK;
}

template <typename T, int NT, bool WIDE1 = false>
static int skinny_launch(const CUtensorMap &ma, const CUtensorMap &mw, const SkArgs &args, int grid, cudaStream_t st) {
    constexpr size_t smem = SkSmem<NT, WIDE1>::BYTES;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(w4a16_skinny_kernel<T, NT, WIDE1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess ||
            cudaFuncSetAttribute(w4a16_skinny_kernel<T, NT, WIDE1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
            return fail(TL_ECUDA, "quantized_matmul: cannot raise shared memory limit");
        configured = true;
    }
    cudaError_t e = launch_chained(w4a16_skinny_kernel<T, NT, WIDE1>, dim3(grid), dim3(SK_THREADS), smem, st, ma, mw, args);
    if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_skinny: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("w4a16_skinny");
    return TL_OK;
}

template <typename T>
static int skinny_t(const void *scales, const void *biases, const void *a, const void *b, void *out, const void *residual, int M, int N, int K,
                    int epilogue, void *ws, size_t ws_bytes, const void *norm_w, float norm_eps, void *normed, bool *norm_done, cudaStream_t st,
                    int *planes_out) {
    if (!aligned16(a) || !aligned16(b)) return fail(TL_EINVAL, "quantized_matmul: a and b must be 16-byte aligned");
    const int NT = M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128));
    const int tiles = (K + SK_FEAT - 1) / SK_FEAT;
    const int num_gb = N / SK_GB;
    SkArgs args{};
    args.scales = scales, args.biases = biases, args.residual = residual, args.out = out;
    args.M = M, args.N = N, args.K = K, args.epilogue = epilogue;
    args.splits = w4a16_skinny_splits(M, N, K);
    args.gb_per_split = (num_gb + args.splits - 1) / args.splits;
    if (args.splits > 1 && (ws == nullptr || ws_bytes < w4a16_skinny_workspace(M, N, K)))
        return fail(TL_EWORKSPACE, "quantized_matmul: workspace too small (%zu < %zu)", ws_bytes, w4a16_skinny_workspace(M, N, K));
    args.partials = static_cast<float *>(ws);
    CUtensorMap ma, mw;
    const CUtensorMapDataType dt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (int e = sk_cached_map(&ma, a, 0, N, M, SK_KB, NT, dt, 2)) return e;
    if (int e = sk_cached_map(&mw, b, 1, static_cast<cuuint64_t>(N) / 2, K, SK_PG * SK_GB / 2, SK_FEAT, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1)) return e;
    const int grid = tiles * args.splits;
    int rc;
    switch (NT) {
        case 16: rc = skinny_launch<T, 16>(ma, mw, args, grid, st); break;
        case 32: rc = skinny_launch<T, 32>(ma, mw, args, grid, st); break;
        case 64: rc = skinny_launch<T, 64>(ma, mw, args, grid, st); break;
        default: rc = skinny_wide1() ? skinny_launch<T, 128, true>(ma, mw, args, grid, st) : skinny_launch<T, 128>(ma, mw, args, grid, st); break;
    }
    if (planes_out != nullptr) {  // the caller adds the planes itself (fused consumer); 1 = `out` holds the finished result
        *planes_out = args.splits;
        if (args.splits > 1) return rc;
    }
    if (rc != TL_OK || args.splits == 1) return rc;
    if (normed != nullptr && epilogue == SK_EPI_RESIDUAL && K <= 256 * SK_NORM_PER) {
        cudaError_t e = launch_chained(w4a16_skinny_reduce_norm_kernel<T>, dim3(M), dim3(256), 0, st, static_cast<const float *>(args.partials),
                                       static_cast<const T *>(residual), static_cast<const T *>(norm_w), static_cast<T *>(out), static_cast<T *>(normed),
                                       M, K, args.splits, norm_eps);
        if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_skinny_reduce_norm: launch failed: %s", cudaGetErrorString(e));
        TL_LAUNCH_CHECK("w4a16_skinny_reduce_norm");
        *norm_done = true;
        return TL_OK;
    }
    const size_t outputs = static_cast<size_t>(M) * (epilogue == SK_EPI_SWIGLU_PAIRS ? K / 2 : K);
    cudaError_t e = launch_chained(w4a16_skinny_reduce_kernel<T>, dim3(static_cast<unsigned>((outputs + 255) / 256)), dim3(256), 0, st,
                                   static_cast<const float *>(args.partials), static_cast<const T *>(residual), static_cast<T *>(out), M, K,
                                   args.splits, epilogue);
    if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_skinny_reduce: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("w4a16_skinny_reduce");
    return TL_OK;
}

int launch_w4a16_skinny(const void *scales, const void *biases, const void *a, const void *b, void *out, const void *residual, int M, int N,
                        int K, int epilogue, int dtype, void *ws, size_t ws_bytes, cudaStream_t st, const void *norm_w, float norm_eps, void *normed,
                        bool *norm_done, int *planes_out) {
    bool unused = false;
    if (norm_done == nullptr) norm_done = &unused;
    *norm_done = false;
    if (M == 0 || K == 0) return TL_OK;
    if (epilogue == SK_EPI_SWIGLU_PAIRS && K % 16 != 0) return fail(TL_EINVAL, "quantized_matmul: interleaved gate|up rows need K %% 16 == 0");
    if (dtype == TL_BF16)
        return skinny_t<__nv_bfloat16>(scales, biases, a, b, out, residual, M, N, K, epilogue, ws, ws_bytes, norm_w, norm_eps, normed, norm_done, st,
                                       planes_out);
    if (dtype == TL_F16)
        return skinny_t<__half>(scales, biases, a, b, out, residual, M, N, K, epilogue, ws, ws_bytes, norm_w, norm_eps, normed, norm_done, st, planes_out);
    return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
}

}  // namespace tl
