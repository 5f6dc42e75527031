// W4A16 prefill GEMM on CTA PAIRS (tcgen05 cta_group::2): the large-M form of w4a16_gemm.cu.
//
//   out[m, n] = sum_k a[m, k] * T(code[n, k] * scale[n, k/128] + bias[n, k/128])       (m = token, n = feature)
//
// Why: the one-CTA kernel is bound by L2 -> SM bandwidth, not by the tensor core.  Its per-block cycle trace
// (tools/gemm_blocks.py, profiles/r02_gemm_blocks.md) shows the activation tiles arriving just in time and the MMA warp
// paced at ~1500 cycles per 64-wide block with four token tiles (1024 cycles of tensor work): 68 KB per block and SM is
// 45 B/cycle, the chip's ~42 B/cycle/SM L2 ceiling (B300_MICROARCH.md: LTS cap ~6300 B/cycle).  Every CTA needs
// tokens x 64 x 2 bytes of activations per 128 features it computes; the only way to need less is to compute more
// features per activation byte - and the dequantisers cannot feed a second weight tile per CTA.  A CTA pair can: with
// cta_group::2 one MMA spans 256 token rows (128 per CTA) x 256 features, each CTA dequantises HALF of the weight tile
// (its 128 features) into its own shared memory and the tensor cores of both SMs read both halves.  Per CTA and block:
// 32 KB of activations (two 256-token MMA tiles) + one dequantised tile for 1024 cycles of tensor work: 35 B/cycle.
//
// Roles per CTA (384 threads): warp 0 TMA producer (its 128 token rows of each MMA tile), warp 1 MMA issuer in the
// leader CTA (rank 0) / readiness forwarder in the peer, warp 2 TMEM allocation (cta_group::2, all 512 columns: two
// accumulators of 256 features), warps 4-11 dequantisers then epilogue (lane = token row).
// Barriers live at the same offsets in both CTAs: full (own TMA + own dequantisers; in the leader also the peer's
// "my tiles of this stage are complete"), empty and tmem_full (arrived in BOTH CTAs by the leader's multicast commit).
// Replaces quantized_matmul_simdgroup_w4a16_g128 (/root/reference/src/extensions_ref/src/quantized_matmul.metal:96-249)
// for M > 256; same rounding points as w4a16_gemm.cu (weights rounded to the activation dtype, fp32 accumulation).
#include <cuda.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"
#include "trace.cuh"

namespace tl {

constexpr int P_TOK = 128;        // token rows per CTA and MMA tile (the pair's MMA spans 256)
constexpr int P_FEAT = 128;       // features dequantised per CTA (the pair's MMA spans 256)
constexpr int P_KB = 64;          // reduction elements per stage (one 128-byte swizzle atom)
constexpr int P_MT = 2;           // 256-token MMA tiles per pair: 2 x 256 accumulator columns = all of TMEM
constexpr int P_STAGES = 4;
constexpr int P_TILE = P_TOK * P_KB * 2;          // 16 KiB, activation and weight tiles alike
constexpr int P_A_STAGE = P_MT * P_TILE;
constexpr int P_STAGE = P_A_STAGE + P_TILE;       // 48 KiB
constexpr int P_BAR_OFF = P_STAGES * P_STAGE;     // 192 KiB
constexpr int P_SMEM = P_BAR_OFF + 256;
constexpr int P_THREADS = 384;
constexpr int P_DEQ_WARPS = 8;
constexpr int P_TMEM_COLS = 512;

// ---- cluster / pair primitives ------------------------------------------------
__device__ __forceinline__ uint32_t p_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void p_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void p_mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar),
        "r"(rank)
        : "memory");
}
// wait with cluster-scope acquire (the arrival came from the other CTA)
__device__ __forceinline__ void p_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void p_tc_mma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void p_tc_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(static_cast<unsigned short>(3))
                 : "memory");
}

template <typename T>
__host__ __device__ constexpr uint32_t p_instr_desc() {
    const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;  // 0 = f16, 1 = bf16
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (0u << 15) | (0u << 16) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
}

template <typename T>
struct PDeq;
template <>
struct PDeq<__nv_bfloat16> {
    using V2 = __nv_bfloat162;
    static constexpr uint32_t MAGIC = 0x43004300u;  // (128, 128)
};
template <>
struct PDeq<__half> {
    using V2 = __half2;
    static constexpr uint32_t MAGIC = 0x64006400u;  // (1024, 1024)
};

template <typename T>
__global__ void __launch_bounds__(P_THREADS, 1) w4a16_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const T *__restrict__ scales,
                                                                   const T *__restrict__ biases, const uint32_t *__restrict__ b, T *__restrict__ out,
                                                                   int M, int N, int K, int vec_store) {
    extern __shared__ __align__(1024) unsigned char psm[];
#if TL_TRACE
    // per-block cycle stamps of the leader CTA (0,0) (tools/gemm_blocks.py): role 0 = MMA warp (k: stage complete in both
    // CTAs, -, MMAs issued), role 1 = first dequantiser warp (k: loop top, math done, stage free, handed over)
    __shared__ unsigned long long trc[2][40][4];
#define P_TRC(role, i, k)                                                                                                    \
    do {                                                                                                                     \
        if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (i) < 40) trc[role][i][k] = clock64();         \
    } while (0)
#else
#define P_TRC(role, i, k) do { } while (0)
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = p_cta_rank();                    // 0 = leader (issues the MMAs)
    const int n_pair = blockIdx.x >> 1, m_pair = blockIdx.y;
    const int num_kb = N / P_KB;
    const int G = N / 128;

    const uint32_t smem0 = g_smem_u32(psm);
    const uint32_t bar = smem0 + P_BAR_OFF;
    // ONE "full" barrier per stage and CTA: own TMA (expect_tx arrival + bytes) + own eight dequantiser warps; in the
    // leader one more arrival, sent by the peer when ITS barrier of the stage has completed.  The MMA warp pays a single
    // try_wait per block: what it executes between two blocks' MMAs is time the tensor cores idle.
    const uint32_t full = bar, empty = bar + 8 * P_STAGES;
    const uint32_t tmem_full = bar + 16 * P_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(psm + P_BAR_OFF + 16 * P_STAGES + 8);

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < P_STAGES; ++i) {
            g_mbar_init(full + 8 * i, 1 + P_DEQ_WARPS + (rank == 0 ? 1 : 0));
            g_mbar_init(empty + 8 * i, 1);
        }
        g_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // both CTAs of the pair issue the paired allocation (same shared-memory slot offset)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "n"(P_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    g_tc_fence_before();
    p_cluster_sync();  // barriers of both CTAs are initialised before anyone arrives remotely; also a CTA barrier
    g_tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: this CTA's 128 token rows of each MMA tile
        int s = 0;
        uint32_t ph = 1;
        for (int kb = 0; kb < num_kb; ++kb) {
            g_mbar_wait(empty + 8 * s, ph);
            if (g_elect_one()) {
                g_mbar_expect_tx(full + 8 * s, P_A_STAGE);
#pragma unroll
                for (int i = 0; i < P_MT; ++i)  // token rows beyond M are zero-filled by the TMA unit
                    g_tma_load_2d(smem0 + s * P_STAGE + i * P_TILE, &tmap_a, kb * P_KB, (m_pair * P_MT + i) * 256 + static_cast<int>(rank) * P_TOK,
                                  full + 8 * s);
            }
            __syncwarp();
            if (++s == P_STAGES) s = 0, ph ^= 1u;
        }
    } else if (warp == 1) {
        int s = 0;
        uint32_t ph = 0;
        if (rank == 0) {
            // ------------------------------------------------ MMA issuer (leader): 8 MMAs of 256 x 256 x 16 per stage
            constexpr uint32_t idesc = p_instr_desc<T>();
            const uint64_t adesc0 = g_smem_desc_sw128(smem0, 0, 1024), bdesc0 = g_smem_desc_sw128(smem0 + P_A_STAGE, 0, 1024);
            for (int kb = 0; kb < num_kb; ++kb) {
                p_mbar_wait_cluster(full + 8 * s, ph);  // cluster-scope acquire: one of the arrivals is the peer's
                P_TRC(0, kb, 0);
                P_TRC(0, kb, 1);
                g_tc_fence_after();
                if (g_elect_one()) {
                    const uint64_t stage = static_cast<uint64_t>(s * (P_STAGE >> 4));
#pragma unroll
                    for (int i = 0; i < P_MT; ++i) {
#pragma unroll
                        for (int k = 0; k < P_KB / 16; ++k)  // +32 bytes along K per step: +2 in the (addr >> 4) field
                            p_tc_mma2(tmem_d + i * 256, adesc0 + stage + i * (P_TILE >> 4) + 2 * k, bdesc0 + stage + 2 * k, idesc,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    p_tc_commit2(empty + 8 * s);  // both CTAs may refill the stage once these MMAs have read it
                }
                __syncwarp();
                P_TRC(0, kb, 2);
                if (++s == P_STAGES) s = 0, ph ^= 1u;
            }
            if (g_elect_one()) p_tc_commit2(tmem_full);
            __syncwarp();
        } else {
            // ------------------------------------------------ peer: tell the leader when this CTA's tiles of a stage are complete
            for (int kb = 0; kb < num_kb; ++kb) {
                g_mbar_wait(full + 8 * s, ph);
                if (lane == 0) p_mbar_arrive_remote(full + 8 * s, 0);
                __syncwarp();
                if (++s == P_STAGES) s = 0, ph ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ dequantisers: this CTA's 128 features, thread = (row, 32-element half)
        const int dt = threadIdx.x - 128;
        const int row = dt >> 1, half = dt & 1;  // adjacent lanes = the two 16-byte halves of one weight row
        const int n = min(n_pair * 256 + static_cast<int>(rank) * P_FEAT + row, K - 1);
        const uint32_t *wrow = b + static_cast<size_t>(n) * (N / 8);
        const T *srow = scales + static_cast<size_t>(n) * G;
        const T *crow = biases + static_cast<size_t>(n) * G;
        using V2 = typename PDeq<T>::V2;
        const uint32_t magic = PDeq<T>::MAGIC;
        const V2 offset2 = *reinterpret_cast<const V2 *>(&magic);
        uint4 packed = *reinterpret_cast<const uint4 *>(wrow + half * 4);
        T sc_next = srow[0], bi_next = crow[0];
        int s = 0;
        uint32_t ph = 1;
        for (int kb = 0; kb < num_kb; ++kb) {
            if (warp == 4) P_TRC(1, kb, 0);
            const uint4 cur = packed;
            const T sc = sc_next, bi = bi_next;
            if (kb + 1 < num_kb) {  // next block's words and (every other block) the next group's scale pair: one round trip ahead
                packed = *reinterpret_cast<const uint4 *>(wrow + (kb + 1) * 8 + half * 4);
                sc_next = srow[(kb + 1) >> 1], bi_next = crow[(kb + 1) >> 1];
            }
            V2 s2, b2;
            s2.x = sc, s2.y = sc, b2.x = bi, b2.y = bi;
            uint32_t outw[16];
            const uint32_t wv[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t p[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t bits;  // (128 + q_i, 128 + q_{i+4}) in one LOP3
                    asm("lop3.b32 %0, %1, 0x000F000F, %2, 0xEA;" : "=r"(bits) : "r"(wv[j] >> (4 * i)), "r"(magic));
                    V2 q = __hsub2(*reinterpret_cast<V2 *>(&bits), offset2);     // exact codes
                    V2 v = __hfma2(q, s2, b2);                                    // q*scale+bias, one rounding
                    p[i] = *reinterpret_cast<uint32_t *>(&v);
                }
                outw[4 * j + 0] = __byte_perm(p[0], p[1], 0x5410);  // (e0, e1)
                outw[4 * j + 1] = __byte_perm(p[2], p[3], 0x5410);  // (e2, e3)
                outw[4 * j + 2] = __byte_perm(p[0], p[1], 0x7632);  // (e4, e5)
                outw[4 * j + 3] = __byte_perm(p[2], p[3], 0x7632);  // (e6, e7)
            }
            if (warp == 4) P_TRC(1, kb, 1);
            g_mbar_wait(empty + 8 * s, ph);
            if (warp == 4) P_TRC(1, kb, 2);
            unsigned char *tile = psm + s * P_STAGE + P_A_STAGE + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = half * 4 + j;  // 16-byte chunk (8 elements) along K
                *reinterpret_cast<uint4 *>(tile + ((chunk ^ (row & 7)) << 4)) = make_uint4(outw[4 * j], outw[4 * j + 1], outw[4 * j + 2], outw[4 * j + 3]);
            }
            g_fence_proxy_async();  // generic-proxy stores -> visible to the tensor cores' async proxy (of both SMs)
            __syncwarp();
            if (lane == 0) g_mbar_arrive(full + 8 * s);
            if (warp == 4) P_TRC(1, kb, 3);
            if (++s == P_STAGES) s = 0, ph ^= 1u;
        }
        // ---- epilogue: TMEM lane = token row of THIS CTA; warps 4-7 take features 0..127 of the pair, warps 8-11 features 128..255
        g_mbar_wait(tmem_full, 0);
        g_tc_fence_after();
        const int q = warp & 3;
        const int col_half = (warp - 4) >> 2;
#pragma unroll 1
        for (int i = 0; i < P_MT; ++i) {
            const int m = (m_pair * P_MT + i) * 256 + static_cast<int>(rank) * P_TOK + q * 32 + lane;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                const int col0 = col_half * 128 + cb * 32;
                uint32_t v[32];
                g_tmem_ld32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + i * 256 + col0, v);
                const int n0 = n_pair * 256 + col0;
                if (m < M && n0 < K) {
                    T *dst = out + static_cast<size_t>(m) * K + n0;
                    const int valid = min(32, K - n0);
                    if (vec_store && valid == 32) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 o;
                            o.x = pack2<T>(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                            o.y = pack2<T>(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                            o.z = pack2<T>(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                            o.w = pack2<T>(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                            *reinterpret_cast<uint4 *>(dst + 8 * j) = o;
                        }
                    } else {
                        for (int j = 0; j < valid; ++j) dst[j] = from_f<T>(__uint_as_float(v[j]));
                    }
                }
            }
        }
    }
    g_tc_fence_before();
    p_cluster_sync();  // neither CTA frees tensor memory (or exits) while the other may still depend on it
    if (warp == 2) {
        g_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(P_TMEM_COLS) : "memory");
    }
#if TL_TRACE
    if (blockIdx.x == 0 && blockIdx.y == 0 && g_trace_buf != nullptr) {  // tag = 20000 + role * 1000 + block * 4 + k (as w4a16_gemm.cu)
        for (int e = threadIdx.x; e < 2 * 40 * 4; e += P_THREADS) {
            const int role = e / 160, rest = e - role * 160;
            if (rest / 4 < num_kb && !(role == 0 && (rest & 3) == 3)) {
                const unsigned at = atomicAdd(g_trace_n, 1u);
                if (at < g_trace_cap) g_trace_buf[2 * at] = 20000 + role * 1000 + rest, g_trace_buf[2 * at + 1] = trc[role][rest / 4][rest & 3];
            }
        }
    }
#endif
}

#if TL_TRACE
void trace_bind_gemm2(unsigned long long *buf, unsigned int *n, unsigned int cap) { trace_bind(buf, n, cap); }
#endif

// ---------------------------------------------------------------- host side --
// Mode: 0 = never, 1 = where the pair grid fills the SMs (default), 2 = every M > 256 (tests, A/B runs).  TL_GEMM2 sets
// the initial value, tl_set_gemm_pairs() changes it.
static int pairs_default() {
    const char *e = getenv("TL_GEMM2");
    return e == nullptr ? 1 : atoi(e);
}
static int g_pairs_mode = pairs_default();
void set_gemm_pairs(int mode) { g_pairs_mode = mode; }

// Measured at M = 4096 (profiles/r02_gemm_bench.json): the pair kernel and the one-CTA kernel are within 2 % of each
// other per CTA-second; what decides is the tail - a pair tile is 512 tokens x 256 features, so K = 2560 gives 160 CTAs
// (1.08 waves of 148: 683 / 743 TF/s on o / down against 766 / 869), while q|k|v (384 CTAs) and gate|up (1216) fill
// their last wave to > 85 % (896 / 965 TF/s against 891 / 950).
bool w4a16_gemm2_supported(int M, int N, int K, int dtype) {
    if (g_pairs_mode <= 0 || !(dtype == TL_BF16 || dtype == TL_F16) || M <= 256 || K <= 0 || N % 128 != 0) return false;
    if (g_pairs_mode >= 2) return true;
    const long long ctas = 2LL * ceil_div(K, 256) * ceil_div(M, P_MT * 256);
    const long long waves = (ctas + sm_count() - 1) / sm_count();
    return M >= 1024 && ctas * 100 >= waves * sm_count() * 85;
}

template <typename T>
static int gemm2_t(const void *scales, const void *biases, const void *a, const void *b, void *out, int M, int N, int K, cudaStream_t st) {
    PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
    if (encode == nullptr) return fail(TL_ECUDA, "quantized_matmul: cuTensorMapEncodeTiled is unavailable");
    if (!aligned16(a) || !aligned16(b)) return fail(TL_EINVAL, "quantized_matmul: a and b must be 16-byte aligned");
    // tensor maps cached per (pointer, shape): encoding costs ~2 us of host time per launch
    struct Key {
        const void *p;
        int M, N;
        bool operator==(const Key &o) const { return p == o.p && M == o.M && N == o.N; }
    };
    struct Hash {
        size_t operator()(const Key &k) const { return reinterpret_cast<size_t>(k.p) * 1000003u ^ (static_cast<size_t>(k.M) << 20) ^ static_cast<size_t>(k.N); }
    };
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    CUtensorMap map;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(Key{a, M, N});
        if (it != cache.end()) {
            map = it->second;
        } else {
            const cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
            const cuuint64_t strides[1] = {static_cast<cuuint64_t>(N) * 2};
            const cuuint32_t box[2] = {P_KB, P_TOK};
            const cuuint32_t estr[2] = {1, 1};
            const CUtensorMapDataType dt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
            CUresult r = encode(&map, dt, 2, const_cast<void *>(a), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(TL_ECUDA, "quantized_matmul: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
            if (cache.size() > 4096) cache.clear();
            cache.emplace(Key{a, M, N}, map);
        }
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(w4a16_gemm2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
        if (e != cudaSuccess) return fail(TL_ECUDA, "quantized_matmul: cannot raise shared memory limit: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int vec_store = (K % 8 == 0 && aligned16(out)) ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * ceil_div(K, 256), ceil_div(M, P_MT * 256));
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = P_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, w4a16_gemm2_kernel<T>, map, static_cast<const T *>(scales), static_cast<const T *>(biases),
                                       static_cast<const uint32_t *>(b), static_cast<T *>(out), M, N, K, vec_store);
    if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_gemm2: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("w4a16_gemm2");
    return TL_OK;
}

int launch_w4a16_gemm2(const void *scales, const void *biases, const void *a, const void *b, void *out, int M, int N, int K, int dtype,
                       cudaStream_t st) {
    if (M == 0 || K == 0) return TL_OK;
    if (dtype == TL_BF16) return gemm2_t<__nv_bfloat16>(scales, biases, a, b, out, M, N, K, st);
    if (dtype == TL_F16) return gemm2_t<__half>(scales, biases, a, b, out, M, N, K, st);
    return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
}

}  // namespace tl
