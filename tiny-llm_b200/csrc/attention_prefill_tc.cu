// Paged causal FlashAttention prefill on the 5th-generation tensor cores (tcgen05 + TMEM + TMA),
// bf16, head_dim 128, L > 8.
//
// Replaces paged_attention_mma_bf16_d128 (/root/reference/src/extensions_ref/src/
// paged_attention.metal:250-506): same arithmetic - fp32 scores in base 2 (:414 scale_log2),
// bottom-right causal limit key <= row + (context - L) (:411), fp32 running max / sum, the
// probabilities rounded to bf16 before P V (:439-444), output = O / sum (0 when nothing is visible).
//
// One CTA = one (request, KV head, block of RH query positions): its 128 MMA rows are the G = Hq/Hkv
// query heads of the KV head x RH = 128/G consecutive query positions, so every K/V tile the CTA
// fetches is shared by all the query heads that need it, and the causal frontier is almost the same
// for every row (RH <= 128 positions apart).  Per 64-key tile:
//
//   warp 4 (one lane)  TMA producer: Q once (3-D map [D, L, B*Hq], box 64 x RH x G, 128-byte
//                      swizzle = the K-major UMMA layout), then K and V tiles - a (page, kv head)
//                      slab is one contiguous [page x 128] bf16 block, fetched as 4-D boxes
//                      [64 d x 64 keys] keyed by block_table (invalid page ids land outside the
//                      tensor and are zero-filled by the TMA unit) - through a 2-stage ring;
//   warp 5 (one lane)  MMA issuer: S = Q K^T  (tcgen05.mma kind::f16, M128 N64 K16 x 8, K tile =
//                      K-major B operand), then O += P V (M128 N128 K16 x 4, V tile = MN-major B
//                      operand straight from the page layout), accumulators S and O in TMEM;
//   warps 0-3          softmax: thread = MMA row = TMEM lane.  tcgen05.ld the 64 scores of its row,
//                      scale, mask, running max (no shuffles: a thread owns the row), exp2, row sum,
//                      bf16 probabilities -> shared memory in the swizzled K-major layout (A operand
//                      of P V).  O is rescaled IN TMEM (tcgen05.ld / st) only when a row's maximum
//                      grows by more than 2^8 since the last rescale: P and the row sum always use
//                      the same (possibly stale) maximum, so the result is exact.
//                      Epilogue: O / sum -> bf16 -> global.
//
// 256 TMEM columns and 112 KB of shared memory per CTA: two CTAs per SM, so one CTA's softmax
// overlaps the other's MMAs without an intra-CTA ping-pong.
#include <stdlib.h>
#include <math_constants.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"

namespace tl {

typedef __nv_bfloat16 bf16;

constexpr int TC_D = 128;         // head dim
constexpr int TC_BM = 128;        // MMA rows per CTA (G heads x RH query positions)
constexpr int TC_BN = 64;         // keys per tile
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 6 * 32;
constexpr int TC_SOFTMAX_THREADS = 128;
constexpr int TC_Q_BYTES = TC_BM * TC_D * 2;        // 32 KiB: two 64-column halves of [128 rows x 128 B]
constexpr int TC_KV_TILE = TC_BN * TC_D * 2;        // 16 KiB: two 64-column halves of [64 rows x 128 B]
constexpr int TC_P_BYTES = TC_BM * TC_BN * 2;       // 16 KiB: [128 rows x 128 B]
constexpr int TC_Q_OFF = 0;
constexpr int TC_P_OFF = TC_Q_OFF + TC_Q_BYTES;
constexpr int TC_K_OFF = TC_P_OFF + TC_P_BYTES;
constexpr int TC_V_OFF = TC_K_OFF + TC_STAGES * TC_KV_TILE;
constexpr int TC_BAR_OFF = TC_V_OFF + TC_STAGES * TC_KV_TILE;
constexpr int TC_SMEM_BYTES = TC_BAR_OFF + 256;
constexpr int TC_TMEM_COLS = 256;  // S (two buffers): columns [0, 64) and [64, 128), O: columns [128, 256)
constexpr int TC_TMEM_O = 128;
constexpr float TC_LOG2E = 1.44269504089f;
constexpr float TC_RESCALE_THRESHOLD = 8.0f;  // log2: rescale O when a row maximum grew by more than 2^8

// kind::f16 instruction descriptor: bf16 x bf16 -> f32, A K-major; b_mn: B operand MN-major.
__host__ __device__ constexpr uint32_t tc_instr_desc(int n, bool b_mn) {
    return (1u << 4)                                // c_format = f32
           | (1u << 7)                              // a_format = bf16
           | (1u << 10)                             // b_format = bf16
           | (0u << 15)                             // a_major = K
           | ((b_mn ? 1u : 0u) << 16)               // b_major
           | (static_cast<uint32_t>(n >> 3) << 17)  // n_dim
           | (static_cast<uint32_t>(TC_BM >> 4) << 24);
}

// ex2.approx.ftz: one MUFU instruction (exp2f() adds denormal-range scaling, ~3 more instructions per element; ncu of the
// first version: 10.7 instructions per score element, issue slots 54 % busy with the tensor pipe at 40 %).  Relative error
// 2^-22; the reference kernel uses fast::exp2 (paged_attention.metal:428-436).
__device__ __forceinline__ float tc_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct TcArgs {
    const int32_t *block_table, *context_lens;
    bf16 *out;
    int L, Hq, Hkv, G, RH;       // RH = query positions per CTA = 128 / G
    int page_size, max_pages, num_pages;
    int tiles_per_page;          // page_size / 64
    float scale_log2;
    int is_causal;
    // split-KV (decode, L <= 8): CTA (query block, split) covers key tiles [split * tiles_per_split, ...) and
    // leaves an unnormalised partial (O fp32, running max in log2 units, sum) for paged_gqa_merge_kernel
    int splits, tiles_per_split;
    int out_token_major;  // unsplit launches only: out[(b, l, head), :] instead of [(b, head, l), :] (the layout the o-projection takes)
    float *ws_o, *ws_m, *ws_l;
};

// PT: the probabilities of a tile stay in TENSOR memory - P_j (bf16 pairs, 32 columns) overwrites the first half of the
// score buffer S_j it was computed from, and P V reads its A operand from there (the TS form of tcgen05.mma).  The
// shared-memory form kept ONE P tile, so the exponentials of tile j could not start before P V of tile j-1 had finished
// reading it; now the exponentials overlap it (the wait sits just before the barrier arrival, and in the rare rescale
// of O), and the eight 16-byte stores + proxy fence per row become one tcgen05.st.  Buffer reuse needs no barrier: S_{j+2} is issued behind P V_j in the same MMA queue.
template <bool PT>
__global__ void __launch_bounds__(TC_THREADS, 2)
paged_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const TcArgs a) {
    extern __shared__ __align__(1024) unsigned char tsm[];
    griddep_launch();  // programmatic dependent launch (common.cuh): the successor may set itself up under this grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = static_cast<int>(blockIdx.x) % a.splits;
    const int qb = static_cast<int>(gridDim.x) / a.splits - 1 - static_cast<int>(blockIdx.x) / a.splits;  // long (late) query blocks first
    const int kvh = blockIdx.y, b = blockIdx.z;
    const int q0 = qb * a.RH;
    const int ctx = min(a.context_lens[b], a.max_pages * a.page_size);
    // keys any row of this CTA may see: [0, key_end)
    const int last_row = min(q0 + a.RH, a.L) - 1;
    const int key_end = ctx <= 0 ? 0 : (a.is_causal ? min(ctx, max(ctx - a.L + last_row + 1, 0)) : ctx);
    const int all_tiles = (key_end + TC_BN - 1) / TC_BN;
    // Split-KV: the launch fixes the NUMBER of splits from the block table's width (the grid of a captured graph cannot
    // follow the context), the tiles are dealt out here from the request's real length, so a request far below the
    // bound still keeps all its splits busy.
    const int tps = a.splits > 1 ? max((all_tiles + a.splits - 1) / a.splits, 1) : a.tiles_per_split;
    const int t0 = min(split * tps, all_tiles);
    const int n_tiles = min(tps, all_tiles - t0);  // this CTA: global tiles t0 .. t0 + n_tiles - 1

    const uint32_t q_base = g_smem_u32(tsm + TC_Q_OFF), p_base = g_smem_u32(tsm + TC_P_OFF);
    const uint32_t k_base = g_smem_u32(tsm + TC_K_OFF), v_base = g_smem_u32(tsm + TC_V_OFF);
    const uint32_t bar = g_smem_u32(tsm + TC_BAR_OFF);
    const uint32_t q_full = bar, s_full = bar + 8 /* two: one per S buffer */, p_full = bar + 24, pv_done = bar + 32;
    const uint32_t k_full = bar + 40, k_empty = k_full + 8 * TC_STAGES, v_full = k_full + 16 * TC_STAGES, v_empty = k_full + 24 * TC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tsm + TC_BAR_OFF + 40 + 32 * TC_STAGES);

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
    }
    if (warp == 5 && lane == 0) {
        g_mbar_init(q_full, 1);
        g_mbar_init(s_full, 1);
        g_mbar_init(s_full + 8, 1);
        g_mbar_init(p_full, TC_SOFTMAX_THREADS / 32);  // one arrival per softmax warp
        g_mbar_init(pv_done, 1);
        for (int i = 0; i < TC_STAGES; ++i) {
            g_mbar_init(k_full + 8 * i, 1);
            g_mbar_init(k_empty + 8 * i, 1);
            g_mbar_init(v_full + 8 * i, 1);
            g_mbar_init(v_empty + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    g_tc_fence_before();
    __syncthreads();
    g_tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // q and the newest K/V rows are the predecessor's output, and this grid overwrites buffers (output, split partials)
    // the predecessor may still read; block tables and context lengths are step inputs uploaded before the chain.
    griddep_wait();

    if (warp == 4) {
        // ------------------------------------------------------------ TMA producer
        if (n_tiles > 0) {  // whole warp, one elected lane issues (see the MMA warp)
            const int head0 = b * a.Hq + kvh * a.G;
            if (g_elect_one()) {
                g_mbar_expect_tx(q_full, TC_Q_BYTES);
                g_tma_load_3d(q_base, &tmap_q, 0, q0, head0, q_full);
                g_tma_load_3d(q_base + TC_Q_BYTES / 2, &tmap_q, 64, q0, head0, q_full);
            }
            __syncwarp();
            const int32_t *table = a.block_table + static_cast<size_t>(b) * a.max_pages;
            int s = 0;
            uint32_t ph = 1;  // producer side: first pass through the ring does not wait
            for (int j = 0; j < n_tiles; ++j) {
                const int lp = (t0 + j) / a.tiles_per_page;
                int pid = table[lp];
                if (pid < 0 || pid >= a.num_pages) pid = a.num_pages;  // outside the tensor: the TMA unit writes zeros
                const int slot0 = (t0 + j - lp * a.tiles_per_page) * TC_BN;
                g_mbar_wait(k_empty + 8 * s, ph);
                if (g_elect_one()) {
                    g_mbar_expect_tx(k_full + 8 * s, TC_KV_TILE);
                    g_tma_load_4d(k_base + s * TC_KV_TILE, &tmap_k, 0, slot0, kvh, pid, k_full + 8 * s);
                    g_tma_load_4d(k_base + s * TC_KV_TILE + TC_KV_TILE / 2, &tmap_k, 64, slot0, kvh, pid, k_full + 8 * s);
                }
                __syncwarp();
                g_mbar_wait(v_empty + 8 * s, ph);
                if (g_elect_one()) {
                    g_mbar_expect_tx(v_full + 8 * s, TC_KV_TILE);
                    g_tma_load_4d(v_base + s * TC_KV_TILE, &tmap_v, 0, slot0, kvh, pid, v_full + 8 * s);
                    g_tma_load_4d(v_base + s * TC_KV_TILE + TC_KV_TILE / 2, &tmap_v, 64, slot0, kvh, pid, v_full + 8 * s);
                }
                __syncwarp();
                if (++s == TC_STAGES) s = 0, ph ^= 1u;
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------ MMA issuer
        // Order: S_0, then per tile j: S_{j+1} (second score buffer) | wait P_j | O += P_j V_j.  The tensor core works
        // on the next tile's scores while the softmax warps are busy with this tile's.
        // The whole warp walks the loop (uniform control flow, running stage / phase counters, descriptors derived by
        // adding to three base descriptors), one elected lane issues (tc05.cuh: g_elect_one - behind `if (lane == 0)`
        // every tcgen05 instruction sat in an ELECT loop with its operands moved through R2UR, and the issue latency
        // of a tile was of the order of its 512 cycles of tensor work).
        if (n_tiles > 0) {
            constexpr uint32_t idesc_s = tc_instr_desc(TC_BN, false);
            constexpr uint32_t idesc_o = tc_instr_desc(TC_D, true);
            const uint64_t qdesc0 = g_smem_desc_sw128(q_base, 0, 1024), pdesc0 = g_smem_desc_sw128(p_base, 0, 1024);
            const uint64_t kdesc0 = g_smem_desc_sw128(k_base, 0, 1024);
            // V: 16 keys = 2 groups of 8 rows (SBO 1024 B); the two 64-wide d blocks are TC_KV_TILE/2 apart (LBO)
            const uint64_t vdesc0 = g_smem_desc_sw128(v_base, TC_KV_TILE / 2, 1024);
            g_mbar_wait(q_full, 0);
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0;
            auto issue_scores = [&](int j) {  // S_j = Q K_j^T: both operands K-major, two 64-wide halves of the head dimension
                g_mbar_wait(k_full + 8 * ks, kph);
                g_tc_fence_after();
                if (g_elect_one()) {
                    const uint64_t kd = kdesc0 + static_cast<uint64_t>(ks * (TC_KV_TILE >> 4));
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t half = (k >> 2), kk = (k & 3);
                        g_tc_mma(tmem + (j & 1) * TC_BN, qdesc0 + half * (TC_Q_BYTES / 2 >> 4) + 2 * kk, kd + half * (TC_KV_TILE / 2 >> 4) + 2 * kk,
                                 idesc_s, k > 0 ? 1u : 0u);
                    }
                    g_tc_commit(k_empty + 8 * ks);       // K stage reusable once these MMAs have read it
                    g_tc_commit(s_full + 8 * (j & 1));   // ... and the scores of tile j are complete
                }
                __syncwarp();
                if (++ks == TC_STAGES) ks = 0, kph ^= 1u;
            };
            issue_scores(0);
            for (int j = 0; j < n_tiles; ++j) {
                // score buffer (j+1)&1 was last read for tile j-1, whose P has been waited for below
                if (j + 1 < n_tiles) issue_scores(j + 1);
                // ---- O += P V: P K-major [128 x 64 keys], V MN-major [64 keys x 128 d] as loaded from the page
                g_mbar_wait(p_full, j & 1);
                g_mbar_wait(v_full + 8 * vs, vph);
                g_tc_fence_after();
                if (g_elect_one()) {
                    const uint64_t vd = vdesc0 + static_cast<uint64_t>(vs * (TC_KV_TILE >> 4));
#pragma unroll
                    for (int k = 0; k < TC_BN / 16; ++k) {
                        if constexpr (PT)  // 16 keys = 8 columns of bf16 pairs per K step
                            g_tc_mma_ts(tmem + TC_TMEM_O, tmem + (j & 1) * TC_BN + 8 * k, vd + k * (2048 >> 4), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                        else
                            g_tc_mma(tmem + TC_TMEM_O, pdesc0 + 2 * k, vd + k * (2048 >> 4), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    g_tc_commit(v_empty + 8 * vs);
                    g_tc_commit(pv_done);  // O holds tiles 0..j and the P buffer is free again
                }
                __syncwarp();
                if (++vs == TC_STAGES) vs = 0, vph ^= 1u;
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps (thread = row = TMEM lane)
        const int r = threadIdx.x;                 // 0..127
        const int g = r / a.RH, lq = r - g * a.RH;  // query head of the KV group, position inside the block
        const int l = q0 + lq;
        const bool row_valid = l < a.L && ctx > 0;
        // last key this row may see (bottom-right causal alignment, paged_attention.metal:411)
        const int limit = !row_valid ? -1 : (a.is_causal ? min(ctx - 1, l + (ctx - a.L)) : ctx - 1);
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        float m_used = -CUDART_INF_F, l_sum = 0.f;
        const int32_t *table = a.block_table + static_cast<size_t>(b) * a.max_pages;
        for (int j = 0; j < n_tiles; ++j) {
            const int lp = (t0 + j) / a.tiles_per_page;
            const int pid = table[lp];
            const bool page_ok = pid >= 0 && pid < a.num_pages;
            g_mbar_wait(s_full + 8 * (j & 1), (j >> 1) & 1);
            g_tc_fence_after();
            uint32_t sv[2][32];
            g_tmem_ld32_nowait(tmem + lane_base + (j & 1) * TC_BN, sv[0]);
            g_tmem_ld32_nowait(tmem + lane_base + (j & 1) * TC_BN + 32, sv[1]);
            g_tmem_ld_wait();
            const int key0 = (t0 + j) * TC_BN;
            const int visible = page_ok ? min(limit - key0 + 1, TC_BN) : 0;  // keys [0, visible) of the tile
            // raw maximum (the scale is positive, so it commutes with max); masking only on boundary tiles
            float raw_max = -CUDART_INF_F;
            if (visible >= TC_BN) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < 32; ++c) raw_max = fmaxf(raw_max, __uint_as_float(sv[h][c]));
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if ((h * 32 + c) >= visible) sv[h][c] = 0xff800000u;  // -inf: exp2 gives exactly 0
                        raw_max = fmaxf(raw_max, __uint_as_float(sv[h][c]));
                    }
            }
            const float tile_max = raw_max * a.scale_log2;  // -inf stays -inf
            if constexpr (!PT)
                if (j > 0) g_mbar_wait(pv_done, (j - 1) & 1);  // P V of tile j-1 complete: O is stable, the P buffer is free
            // lazy rescale: keep the stale maximum unless this tile exceeds it by more than 2^8
            const bool grow = tile_max > m_used + TC_RESCALE_THRESHOLD || (m_used == -CUDART_INF_F && tile_max != -CUDART_INF_F);
            if (__any_sync(0xffffffffu, grow) && j > 0) {
                if constexpr (PT) g_mbar_wait(pv_done, (j - 1) & 1);  // O must be stable while it is rescaled
                g_tc_fence_after();
                const float m_new = grow ? tile_max : m_used;
                const float alpha = (grow && m_used != -CUDART_INF_F) ? exp2f(m_used - m_new) : (grow ? 0.f : 1.f);
                l_sum *= alpha;
#pragma unroll
                for (int cb = 0; cb < TC_D / 32; ++cb) {
                    uint32_t ov[32];
                    g_tmem_ld32(tmem + lane_base + TC_TMEM_O + cb * 32, ov);
#pragma unroll
                    for (int c = 0; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * alpha);
                    g_tmem_st32(tmem + lane_base + TC_TMEM_O + cb * 32, ov);
                }
                g_tmem_st_wait();
                m_used = m_new;
            } else if (grow) {
                m_used = tile_max;  // first tile: nothing accumulated yet
            }
            const float m_eff = m_used == -CUDART_INF_F ? 0.f : m_used;
            float tile_sum = 0.f;
            uint32_t pk[32];  // 64 bf16 probabilities
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float p0 = tc_ex2(fmaf(__uint_as_float(sv[h][c]), a.scale_log2, -m_eff));
                    const float p1 = tc_ex2(fmaf(__uint_as_float(sv[h][c + 1]), a.scale_log2, -m_eff));
                    tile_sum += p0 + p1;
                    pk[h * 16 + c / 2] = pack2<bf16>(p0, p1);
                }
            l_sum += tile_sum;
            if constexpr (PT) {
                // lane = row, column c = keys (2c, 2c + 1) of the tile: over the first 32 columns of the score buffer just read
                g_tmem_st32(tmem + lane_base + (j & 1) * TC_BN, pk);
                g_tmem_st_wait();
                // Do not ARRIVE for tile j before P V of tile j-1 has completed (its exponentials above did overlap it): p_full
                // counts four arrivals per phase, and a warp that ran a whole tile ahead would arrive twice in one phase and
                // complete it without the slowest warp - P V_j then read rows that were not written yet (1907 of 16.7 M
                // outputs off by up to 3 % in the constant-V test of the first version).
                if (j > 0) g_mbar_wait(pv_done, (j - 1) & 1);
            } else {
                // P row r: 128 bytes = 8 chunks of 16 B, chunk c at (c ^ (r & 7)) - the 128-byte swizzle of a K-major tile
                unsigned char *prow = tsm + TC_P_OFF + r * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4 *>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                g_fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            }
            g_tc_fence_before();    // orders this thread's tcgen05.ld / st before the MMAs released by the arrive
            __syncwarp();           // one arrival per warp: 128 arrivals on one mbarrier are 128 serialised shared-memory atomics
            if (lane == 0) g_mbar_arrive(p_full);
        }
        // ---- epilogue: O / sum -> bf16 -> out[(b*Hq + head) * L + l, :]
        if (n_tiles > 0) {
            g_mbar_wait(pv_done, (n_tiles - 1) & 1);
            g_tc_fence_after();
        }
        const size_t out_row = static_cast<size_t>(b * a.Hq + kvh * a.G + g) * a.L + l;
        if (a.splits > 1) {  // unnormalised partial for the merge kernel (attention_decode.cu: paged_gqa_merge_kernel)
            const size_t slot = out_row * a.splits + split;
            float *po = a.ws_o + slot * TC_D;
#pragma unroll
            for (int cb = 0; cb < TC_D / 32; ++cb) {
                uint32_t ov[32];
                if (n_tiles > 0) {  // warp-uniform: tcgen05.ld is warp-collective, rows beyond L take part too
                    g_tmem_ld32(tmem + lane_base + TC_TMEM_O + cb * 32, ov);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) ov[c] = 0u;
                }
                if (l < a.L) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<uint4 *>(po + cb * 32 + c) = make_uint4(ov[c], ov[c + 1], ov[c + 2], ov[c + 3]);
                }
            }
            if (l < a.L) {
                a.ws_m[slot] = (m_used == -CUDART_INF_F || !row_valid) ? -1e30f : m_used;
                a.ws_l[slot] = row_valid ? l_sum : 0.f;
            }
        } else {
        const float inv = (l_sum == 0.f || !row_valid) ? 0.f : 1.0f / l_sum;
        bf16 *dst = a.out + (a.out_token_major ? (static_cast<size_t>(b) * a.L + l) * a.Hq + (kvh * a.G + g) : out_row) * TC_D;
#pragma unroll
        for (int cb = 0; cb < TC_D / 32; ++cb) {
            uint32_t ov[32];
            if (n_tiles > 0) {
                g_tmem_ld32(tmem + lane_base + TC_TMEM_O + cb * 32, ov);
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) ov[c] = 0u;
            }
            if (l < a.L) {
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                    uint4 o;
                    o.x = pack2<bf16>(__uint_as_float(ov[c]) * inv, __uint_as_float(ov[c + 1]) * inv);
                    o.y = pack2<bf16>(__uint_as_float(ov[c + 2]) * inv, __uint_as_float(ov[c + 3]) * inv);
                    o.z = pack2<bf16>(__uint_as_float(ov[c + 4]) * inv, __uint_as_float(ov[c + 5]) * inv);
                    o.w = pack2<bf16>(__uint_as_float(ov[c + 6]) * inv, __uint_as_float(ov[c + 7]) * inv);
                    *reinterpret_cast<uint4 *>(dst + cb * 32 + c) = o;
                }
            }
        }
        }
    }
    g_tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        g_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side --
// Tensor maps are cached per (base pointer, shape): a prefill of 36 layers x N chunks re-uses a
// handful of distinct operands, and cuTensorMapEncodeTiled costs microseconds per call.
struct MapKey {
    const void *ptr;
    unsigned long long d0, d1, d2, d3;
    unsigned b1, b2;
    bool operator==(const MapKey &o) const { return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && b1 == o.b1 && b2 == o.b2; }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        for (unsigned long long v : {k.d0, k.d1, k.d2, k.d3, static_cast<unsigned long long>(k.b1), static_cast<unsigned long long>(k.b2)})
            h = h * 1000003u ^ static_cast<size_t>(v);
        return h;
    }
};
static int cached_map(CUtensorMap *out, const void *ptr, int rank, const cuuint64_t *dims, const cuuint32_t *box) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, dims[0], dims[1], dims[2], rank > 3 ? dims[3] : 0, box[1], box[2]};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return TL_OK;
    }
    PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
    if (encode == nullptr) return fail(TL_ECUDA, "paged_attention: cuTensorMapEncodeTiled is unavailable");
    cuuint64_t strides[3];
    cuuint64_t acc = 2;
    for (int i = 0; i + 1 < rank; ++i) {
        acc *= dims[i];
        strides[i] = acc;
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap map;
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void *>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TL_ECUDA, "paged_attention: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, map);
    *out = map;
    return TL_OK;
}

bool paged_prefill_tc_supported(int L, int num_pages, int page_size, int num_kv_heads, int num_heads) {
    static const bool off = [] { const char *e = getenv("TL_PREFILL_TC"); return e != nullptr && e[0] == '0'; }();
    if (off || num_kv_heads < 1 || num_heads % num_kv_heads != 0) return false;
    const int G = num_heads / num_kv_heads;
    if (G < 1 || G > TC_BM || (TC_BM % G) != 0) return false;
    return L > 0 && num_pages > 0 && page_size >= TC_BN && page_size % TC_BN == 0;
}

// allow_split (decode, L <= 8): the key range of every (request, KV head) is cut into up to GQA_MAX_SPLITS (32, the
// size paged_decode_workspace() budgets for) pieces of >= 4 tiles so that a small batch still fills the GPU; the
// partials are combined by paged_gqa_merge_kernel.
int launch_paged_prefill_tc(const void *q, const void *kp, const void *vp, const int32_t *bt, const int32_t *cl, void *out, int rows,
                            int L, int num_pages, int page_size, int max_pages, float scale, int is_causal, int num_kv_heads,
                            int num_heads, bool allow_split, void *ws, size_t ws_bytes, cudaStream_t st, bool out_token_major) {
    const int G = num_heads / num_kv_heads;
    const int B = rows / num_heads;
    TcArgs a{};
    a.block_table = bt, a.context_lens = cl, a.out = static_cast<bf16 *>(out);
    a.L = L, a.Hq = num_heads, a.Hkv = num_kv_heads, a.G = G, a.RH = TC_BM / G;
    a.page_size = page_size, a.max_pages = max_pages, a.num_pages = num_pages;
    a.tiles_per_page = page_size / TC_BN;
    a.scale_log2 = scale * TC_LOG2E;
    a.is_causal = is_causal;
    CUtensorMap mq, mk, mv;
    {
        const cuuint64_t dims[3] = {TC_D, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(rows)};
        const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(a.RH), static_cast<cuuint32_t>(G)};
        if (int e = cached_map(&mq, q, 3, dims, box)) return e;
    }
    {
        const cuuint64_t dims[4] = {TC_D, static_cast<cuuint64_t>(page_size), static_cast<cuuint64_t>(num_kv_heads), static_cast<cuuint64_t>(num_pages)};
        const cuuint32_t box[4] = {64, TC_BN, 1, 1};
        if (int e = cached_map(&mk, kp, 4, dims, box)) return e;
        if (int e = cached_map(&mv, vp, 4, dims, box)) return e;
    }
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(paged_prefill_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(paged_prefill_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
            cudaFuncSetAttribute(paged_prefill_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(paged_prefill_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
            return fail(TL_ECUDA, "paged_attention: cannot raise shared memory limit");
        configured = true;
    }
    const int q_blocks = (L + a.RH - 1) / a.RH;
    const long long max_tiles = (static_cast<long long>(max_pages) * page_size + TC_BN - 1) / TC_BN;
    a.out_token_major = out_token_major ? 1 : 0;
    a.splits = 1;
    a.tiles_per_split = static_cast<int>(max_tiles < 1 ? 1 : max_tiles);
    if (allow_split && ws != nullptr && !out_token_major) {
        // Split count: the one that minimises waves x (tiles per CTA + fixed cost) + merge launch, in units of one 64-key
        // tile (~1.45 us of HBM time at a 1/296 share), with 296 CTA slots, ~3 tiles of fixed cost per CTA (set-up, Q load,
        // epilogue) and 2 + splits / 4 for the merge launch (it reads one partial per split and row).  (The first policy aimed at 4 x #SMs CTAs: at 64 requests x 1024 tokens
        // that is 1024 CTAs in 3.5 -> 4 waves plus a merge, 60 us, where 512 unsplit CTAs need 2 waves.)
        const long long base = static_cast<long long>(q_blocks) * num_kv_heads * B;
        const long long slots = 2LL * sm_count();
        const size_t rows_total = static_cast<size_t>(rows) * L;
        long long best = 1, best_cost = -1;
        for (long long sp = 1; sp <= 32 && sp <= (max_tiles > 1 ? max_tiles : 1); ++sp) {
            const long long tps = (max_tiles + sp - 1) / sp;
            const long long real = (max_tiles + tps - 1) / tps;
            if (real != sp) continue;
            if (sp > 1 && ws_bytes < rows_total * sp * (TC_D + 2) * sizeof(float)) break;
            const long long waves = (base * sp + slots - 1) / slots;
            const long long cost = 4 * waves * (tps + 3) + (sp > 1 ? 8 + sp : 0);  // x 4: quarter-tile units
            if (best_cost < 0 || cost < best_cost) best_cost = cost, best = sp;
        }
        if (best > 1) {
            a.splits = static_cast<int>(best), a.tiles_per_split = static_cast<int>((max_tiles + best - 1) / best);
            a.ws_o = static_cast<float *>(ws);
            a.ws_m = a.ws_o + rows_total * best * TC_D;
            a.ws_l = a.ws_m + rows_total * best;
        }
    }
    dim3 grid(q_blocks * a.splits, num_kv_heads, B);
    if (grid.y > 65535 || grid.z > 65535) return fail(TL_EINVAL, "paged_attention: too many heads / requests for one launch");
    static const bool p_tmem = [] { const char *e = getenv("TL_FA_P_TMEM"); return e == nullptr || e[0] != '0'; }();  // 0: P through shared memory (control)
    cudaError_t le = p_tmem ? launch_chained(paged_prefill_tc_kernel<true>, grid, dim3(TC_THREADS), TC_SMEM_BYTES, st, mq, mk, mv, a)
                            : launch_chained(paged_prefill_tc_kernel<false>, grid, dim3(TC_THREADS), TC_SMEM_BYTES, st, mq, mk, mv, a);
    if (le != cudaSuccess) return fail(TL_ECUDA, "paged_attention: launch failed: %s", cudaGetErrorString(le));
    TL_LAUNCH_CHECK("paged_prefill_tc");
    if (a.splits > 1) return launch_paged_gqa_merge(a.ws_o, a.ws_m, a.ws_l, out, rows * L, a.splits, st);
    return TL_OK;
}

}  // namespace tl
