// W4A16 (4-bit codes, group 128, affine) weight-streaming kernels.
//
//   out[m, k] = sum_n a[m, n] * (code[k, n] * scale[k, n/128] + bias[k, n/128])
//
// a [M, N] (bf16|f16), b [K, N/8] packed u32 (code i of a word = (w >> 4i) & 15,
// /root/reference/src/tiny_llm_ref/quantize.py:113-115), scales/biases [K, N/128].
//
// w4a16_stream_kernel is the decode kernel (reference: quantized_matvec_x4_fast,
// /root/reference/src/extensions_ref/src/quantized_matmul.metal:441-538).  It is
// HBM-bound: every packed weight byte is read exactly once with 128-bit
// coalesced loads that bypass L1; activations (tiny, shared by every CTA) are
// staged once per CTA in shared memory.  The per-weight ALU work would exceed
// the HBM time on CUDA cores (~3 ops/weight), so the 16x128 code tile of each
// warp is fed to the tensor cores as an exact small-integer bf16 operand:
//
//   * codes are turned into bf16 (128 + q) two at a time with ONE lop3
//     ((w >> 4i) & 0x000F000F | 0x43004300): ~1 ALU op per weight;
//   * mma.sync m16n8k16 (fp32 accumulate) forms D = sum_n (128 + q) * a exactly
//     (products of an 8-bit and a 4-bit significand are exact in fp32);
//   * once per 128-wide group: acc += scale * D + (bias - 128 * scale) * sum_n a,
//     which equals sum_n (q*scale + bias) * a - the same "scale * qdot +
//     bias * asum" factorisation the Metal kernel uses (:515-521), in fp32.
//
// The k-slot <-> n mapping of the MMA is a free permutation (the reduction is
// commutative), chosen so that a lane's A fragment comes from ITS OWN 128-bit
// weight load and the matching B fragment is one 128-bit shared-memory load.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace tl {

template <typename T>
struct Mma;
template <>
struct Mma<__nv_bfloat16> {
    static constexpr uint32_t MAGIC = 0x43004300u;  // bf16 128.0 in both halves
    static constexpr float OFFSET = 128.f;
    static __device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};
template <>
struct Mma<__half> {
    static constexpr uint32_t MAGIC = 0x64006400u;  // fp16 1024.0 in both halves
    static constexpr float OFFSET = 1024.f;
    static __device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

constexpr int STREAM_TEAM_WARPS = 4;  // warps that split the reduction of one 16-row tile
constexpr int STREAM_TEAM_THREADS = STREAM_TEAM_WARPS * 32;
constexpr int STREAM_PREFETCH = 2;  // weight groups in flight per warp beyond the current one

__device__ __forceinline__ void team_barrier(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(STREAM_TEAM_THREADS) : "memory");
}

struct GroupLoad {
    uint4 w0, w1;       // 32 codes of row g and of row g+8 for this lane
    float s0, c0, s1, c1;  // scale and (bias - OFFSET*scale) of the two rows
};

// MT = number of 8-column activation tiles (rows of `a` handled per pass = 8*MT).
template <typename T, int MT>
__global__ void __launch_bounds__(512) w4a16_stream_kernel(const T *__restrict__ scales, const T *__restrict__ biases,
                                                          const T *__restrict__ a_all, const uint32_t *__restrict__ b,
                                                          T *__restrict__ out_all, int M, int N, int K,
                                                          int rows_per_pass) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pass = blockIdx.y;
    const int Mp = min(rows_per_pass, M - pass * rows_per_pass);
    const T *a = a_all + static_cast<size_t>(pass) * rows_per_pass * N;
    T *out = out_all + static_cast<size_t>(pass) * rows_per_pass * K;

    const int words = N / 8;   // packed words per weight row == 8-element activation chunks
    const int G = N / 128;     // quantisation groups per row
    const int teams = blockDim.x / STREAM_TEAM_THREADS;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int team = warp / STREAM_TEAM_WARPS;
    const int wit = warp % STREAM_TEAM_WARPS;  // warp in team
    const int g = lane >> 2;                   // MMA group id: weight row / activation column
    const int t = lane & 3;                    // MMA thread-in-group: k-slot owner

    // shared: [words][Mp] 16-byte permuted activation chunks | asum [G][Mp] | red [teams][4][16][8*MT]
    uint4 *act = reinterpret_cast<uint4 *>(smem_raw);
    float *asum = reinterpret_cast<float *>(smem_raw + static_cast<size_t>(words) * Mp * 16);
    float *red = asum + ((G * Mp + 3) & ~3);

    // ---- stage activations once: chunk order [0,4,1,5,2,6,3,7] so that the B
    // fragment registers {(e0,e4),(e1,e5),(e2,e6),(e3,e7)} are one LDS.128.
    {
        const int total = Mp * words;
        for (int base = (threadIdx.x & ~31); base < total; base += blockDim.x) {
            const int idx = base + lane;
            float part = 0.f;
            int m = 0, c = 0;
            if (idx < total) {
                m = idx / words;
                c = idx - m * words;
                const uint4 raw = *reinterpret_cast<const uint4 *>(a + static_cast<size_t>(m) * N + c * 8);
                uint4 p;
                p.x = __byte_perm(raw.x, raw.z, 0x5410);
                p.y = __byte_perm(raw.x, raw.z, 0x7632);
                p.z = __byte_perm(raw.y, raw.w, 0x5410);
                p.w = __byte_perm(raw.y, raw.w, 0x7632);
                // 4x4 transpose of the chunk order inside a group: the four lanes of an MMA
                // group then read 64 contiguous bytes per sub-step (no bank conflicts).
                const int pos = (c & ~15) | ((c & 3) << 2) | ((c >> 2) & 3);
                act[static_cast<size_t>(pos) * Mp + m] = p;
                const float2 f0 = unpack2<T>(raw.x), f1 = unpack2<T>(raw.y), f2 = unpack2<T>(raw.z), f3 = unpack2<T>(raw.w);
                part = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
            }
            // 16 consecutive chunks (one group) live in 16 consecutive lanes.
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (idx < total && (c & 15) == 0) asum[(c >> 4) * Mp + m] = part;
        }
    }
    __syncthreads();

    const int tiles = (K + 15) / 16;
    float *my_red = red + static_cast<size_t>(team) * STREAM_TEAM_WARPS * 16 * 8 * MT;

    for (int tile = blockIdx.x * teams + team; tile < tiles; tile += gridDim.x * teams) {
        const int row0 = min(tile * 16 + g, K - 1);
        const int row1 = min(tile * 16 + g + 8, K - 1);
        const uint32_t *b0p = b + static_cast<size_t>(row0) * words + 4 * t;
        const uint32_t *b1p = b + static_cast<size_t>(row1) * words + 4 * t;
        const T *s0p = scales + static_cast<size_t>(row0) * G;
        const T *s1p = scales + static_cast<size_t>(row1) * G;
        const T *c0p = biases + static_cast<size_t>(row0) * G;
        const T *c1p = biases + static_cast<size_t>(row1) * G;

        auto load_group = [&](int u) -> GroupLoad {
            GroupLoad r;
            r.w0 = ldg_stream(b0p + 16 * u);
            r.w1 = ldg_stream(b1p + 16 * u);
            r.s0 = to_f(s0p[u]);
            r.s1 = to_f(s1p[u]);
            r.c0 = to_f(c0p[u]) - Mma<T>::OFFSET * r.s0;
            r.c1 = to_f(c1p[u]) - Mma<T>::OFFSET * r.s1;
            return r;
        };

        float acc[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;

        GroupLoad buf[STREAM_PREFETCH + 1];
#pragma unroll
        for (int p = 0; p < STREAM_PREFETCH; ++p) {
            const int u = wit + p * STREAM_TEAM_WARPS;
            if (u < G) buf[p] = load_group(u);
        }

        for (int u = wit; u < G; u += STREAM_TEAM_WARPS) {
            {
                const int un = u + STREAM_PREFETCH * STREAM_TEAM_WARPS;
                if (un < G) buf[STREAM_PREFETCH] = load_group(un);
            }
            const GroupLoad cur = buf[0];
            float d[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f;

            const uint32_t x0[4] = {cur.w0.x, cur.w0.y, cur.w0.z, cur.w0.w};
            const uint32_t x1[4] = {cur.w1.x, cur.w1.y, cur.w1.z, cur.w1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                constexpr uint32_t MASK = 0x000F000Fu;
                const uint32_t p0 = (x0[j] & MASK) | Mma<T>::MAGIC;
                const uint32_t p1 = ((x0[j] >> 4) & MASK) | Mma<T>::MAGIC;
                const uint32_t p2 = ((x0[j] >> 8) & MASK) | Mma<T>::MAGIC;
                const uint32_t p3 = ((x0[j] >> 12) & MASK) | Mma<T>::MAGIC;
                const uint32_t q0 = (x1[j] & MASK) | Mma<T>::MAGIC;
                const uint32_t q1 = ((x1[j] >> 4) & MASK) | Mma<T>::MAGIC;
                const uint32_t q2 = ((x1[j] >> 8) & MASK) | Mma<T>::MAGIC;
                const uint32_t q3 = ((x1[j] >> 12) & MASK) | Mma<T>::MAGIC;
                const int chunk = 16 * u + 4 * j + t;  // staged position of chunk 16u + 4t + j
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int col = mt * 8 + g;
                    uint4 bf = make_uint4(0u, 0u, 0u, 0u);
                    if (col < Mp) bf = act[static_cast<size_t>(chunk) * Mp + col];
                    Mma<T>::mma(d[mt], p0, q0, p1, q1, bf.x, bf.y);
                    Mma<T>::mma(d[mt], p2, q2, p3, q3, bf.z, bf.w);
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int m0 = mt * 8 + 2 * t;
                const float as0 = m0 < Mp ? asum[u * Mp + m0] : 0.f;
                const float as1 = m0 + 1 < Mp ? asum[u * Mp + m0 + 1] : 0.f;
                acc[mt][0] += cur.s0 * d[mt][0] + cur.c0 * as0;
                acc[mt][1] += cur.s0 * d[mt][1] + cur.c0 * as1;
                acc[mt][2] += cur.s1 * d[mt][2] + cur.c1 * as0;
                acc[mt][3] += cur.s1 * d[mt][3] + cur.c1 * as1;
            }
#pragma unroll
            for (int p = 0; p < STREAM_PREFETCH; ++p) buf[p] = buf[p + 1];
        }

        // ---- reduce the 4 warps of the team, then 128 threads store 16 x (8*MT) outputs
        float *wred = my_red + static_cast<size_t>(wit) * 16 * 8 * MT;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            wred[g * 8 * MT + mt * 8 + 2 * t] = acc[mt][0];
            wred[g * 8 * MT + mt * 8 + 2 * t + 1] = acc[mt][1];
            wred[(g + 8) * 8 * MT + mt * 8 + 2 * t] = acc[mt][2];
            wred[(g + 8) * 8 * MT + mt * 8 + 2 * t + 1] = acc[mt][3];
        }
        team_barrier(team);
        const int tt = threadIdx.x % STREAM_TEAM_THREADS;
#pragma unroll
        for (int rep = 0; rep < MT; ++rep) {
            const int o = rep * STREAM_TEAM_THREADS + tt;  // o = m * 16 + r
            const int m = o >> 4;
            const int r = o & 15;
            const int k = tile * 16 + r;
            if (m < Mp && k < K) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < STREAM_TEAM_WARPS; ++w) v += my_red[(w * 16 + r) * 8 * MT + m];
                out[static_cast<size_t>(m) * K + k] = from_f<T>(v);
            }
        }
        team_barrier(team);
    }
}

static size_t stream_smem_bytes(int N, int Mp, int MT, int teams) {
    const size_t act = static_cast<size_t>(N / 8) * Mp * 16;
    const size_t asum = static_cast<size_t>(((N / 128) * Mp + 3) & ~3) * 4;
    const size_t red = static_cast<size_t>(teams) * STREAM_TEAM_WARPS * 16 * 8 * MT * 4;
    return act + asum + red;
}

template <typename T, int MT>
static int stream_launch(const void *scales, const void *biases, const void *a, const void *b, void *out, int M, int N,
                         int K, int rows_per_pass, int teams, cudaStream_t st) {
    static size_t configured = 0;
    const size_t smem = stream_smem_bytes(N, rows_per_pass < M ? rows_per_pass : M, MT, teams);
    if (smem > configured && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(w4a16_stream_kernel<T, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024);
        if (e != cudaSuccess) return fail(TL_ECUDA, "quantized_matmul: cannot raise shared memory limit: %s", cudaGetErrorString(e));
        configured = 200 * 1024;
    }
    const int tiles = ceil_div(K, 16);
    const int want = ceil_div(tiles, teams);
    const int cap = sm_count() * 4;
    dim3 grid(want < cap ? want : cap, ceil_div(M, rows_per_pass));
    w4a16_stream_kernel<T, MT><<<grid, teams * STREAM_TEAM_THREADS, smem, st>>>(
        static_cast<const T *>(scales), static_cast<const T *>(biases), static_cast<const T *>(a),
        static_cast<const uint32_t *>(b), static_cast<T *>(out), M, N, K, rows_per_pass);
    TL_LAUNCH_CHECK("w4a16_stream");
    return TL_OK;
}

template <typename T>
static int stream_t(const void *scales, const void *biases, const void *a, const void *b, void *out, int M, int N,
                    int K, cudaStream_t st) {
    if (!aligned16(a) || !aligned16(b)) return fail(TL_EINVAL, "quantized_matmul: a and b must be 16-byte aligned");
    // Rows per pass: as many as fit in shared memory next to the reduction scratch.
    const size_t budget = 160 * 1024;
    int rpp = 8;
    if (M > 16 && static_cast<size_t>(N) * 32 * 2 <= budget)
        rpp = 32;
    else if (M > 8 && static_cast<size_t>(N) * 16 * 2 <= budget)
        rpp = 16;
    const int tiles = ceil_div(K, 16);
    const int sms = sm_count();
    int teams = 1;
    if (tiles >= 8 * sms)
        teams = 4;
    else if (tiles >= 4 * sms)
        teams = 2;
    if (rpp == 32) return stream_launch<T, 4>(scales, biases, a, b, out, M, N, K, rpp, teams, st);
    if (rpp == 16) return stream_launch<T, 2>(scales, biases, a, b, out, M, N, K, rpp, teams, st);
    return stream_launch<T, 1>(scales, biases, a, b, out, M, N, K, rpp, teams, st);
}

int launch_w4a16_fused(const void *scales, const void *biases, const void *b, void *out, const void *p0, const void *p1,
                       const void *residual, int M, int N, int K, int lda, int prologue, int epilogue, float eps, int dtype,
                       cudaStream_t st);

static bool stream_v1_forced() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("TL_STREAM_V1");
        cached = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return cached == 1;
}

int launch_w4a16_stream(const void *scales, const void *biases, const void *a, const void *b, void *out, int M, int N,
                        int K, int dtype, cudaStream_t st) {
    if (M == 0 || K == 0) return TL_OK;
    if (!stream_v1_forced())
        return launch_w4a16_fused(scales, biases, b, out, a, nullptr, nullptr, M, N, K, N, 0, 0, 0.f, dtype, st);
    switch (dtype) {
        case TL_F16: return stream_t<__half>(scales, biases, a, b, out, M, N, K, st);
        case TL_BF16: return stream_t<__nv_bfloat16>(scales, biases, a, b, out, M, N, K, st);
    }
    return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
}

// ---------------------------------------------------------------------------
// v2: TMA-bulk fed streaming kernel.
//
// v1 keeps only ~2 KB of weight bytes in flight per warp (register prefetch),
// far short of the ~40 KB per SM that 6.5 TB/s at ~800 ns latency needs.  Here a
// dedicated producer warp streams the packed weights of consecutive 16-row
// tiles into a shared-memory ring with cp.async.bulk (one 256 B row segment per
// lane, completion counted on an mbarrier), so a CTA has STAGES x 4 KB in flight
// at no register cost; eight consumer warps run the same mma-fed arithmetic as
// v1 out of shared memory.  Weights never depend on the previous kernel, so the
// producer starts before griddepcontrol.wait (programmatic dependent launch):
// the ring fills while the previous kernel drains.
//
// Optional fusions that keep every rounding point of the unfused call sequence
// (so results are bit-identical to rms_norm -> matvec, swiglu -> matvec,
// matvec -> add):
//   prologue RMSNORM : a = T(x * rsqrt(mean(x^2)+eps) * w)   (week2_kernels.metal:41-47)
//   prologue SWIGLU  : a = T(g / (1 + exp(-g)) * u)          (week2_kernels.metal:115-116)
//   epilogue RESIDUAL: out = T(float(res) + float(T(acc)))   (qwen3_week3.py:204-206)
constexpr int SK_CONSUMERS = 8;                       // consumer warps
constexpr int SK_THREADS = (SK_CONSUMERS + 1) * 32;   // + 1 producer warp
constexpr int SK_CG = 4;                              // 128-column groups per ring stage
constexpr int SK_ROW_BYTES = SK_CG * 64;              // packed bytes of one row in a stage
constexpr int SK_ROW_STRIDE = SK_ROW_BYTES + 64;      // +64 B: rows g / g+1 hit different bank halves
constexpr int SK_STAGE_BYTES = 16 * SK_ROW_STRIDE;

enum { PRO_NONE = 0, PRO_RMSNORM = 1, PRO_SWIGLU = 2 };
enum { EPI_NONE = 0, EPI_RESIDUAL = 1 };

struct StreamArgs {
    const void *scales, *biases;
    const uint32_t *b;
    void *out;
    const void *p0, *p1, *residual;
    int M, N, K, lda;
    int prologue, epilogue;
    float eps;
    int rows_per_pass, stages;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();  // a lost arrival must not hang the GPU
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() {
    asm volatile("bar.sync 1, %0;" ::"r"(SK_CONSUMERS * 32) : "memory");
}

template <typename T, int MT>
__global__ void __launch_bounds__(SK_THREADS) w4a16_stream2_kernel(const StreamArgs args) {
    extern __shared__ __align__(128) unsigned char smem2_raw[];
    unsigned char *smem_raw = smem2_raw;
    const int N = args.N, K = args.K;
    const int pass = blockIdx.y;
    const int Mp = min(args.rows_per_pass, args.M - pass * args.rows_per_pass);
    const int words = N / 8;
    const int G = N / 128;
    const int S = args.stages;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // shared layout: ring | barriers | act | asum | row stats | red[2]
    unsigned char *ring = smem_raw;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + static_cast<size_t>(S) * SK_STAGE_BYTES);
    uint4 *act = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(bars) + ((2 * S * 8 + 15) & ~15));
    float *asum = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(act) + static_cast<size_t>(words) * Mp * 16);
    float *rowstat = asum + ((G * Mp + 3) & ~3);
    float *red = rowstat + 32;
    const uint32_t ring_s = smem_u32(ring);
    const uint32_t full_s = smem_u32(bars);
    const uint32_t empty_s = full_s + 8 * S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(full_s + 8 * i, 1);
            mbar_init(empty_s + 8 * i, SK_CONSUMERS / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    griddep_launch();

    const int tiles = (K + 15) / 16;
    const int tile_begin = static_cast<int>(static_cast<long long>(tiles) * blockIdx.x / gridDim.x);
    const int tile_end = static_cast<int>(static_cast<long long>(tiles) * (blockIdx.x + 1) / gridDim.x);
    const int cpt = (G + SK_CG - 1) / SK_CG;  // ring stages per tile

    if (warp == SK_CONSUMERS) {
        // ------------------------------ producer: weights only, no dependency on the previous kernel
        const unsigned char *bbytes = reinterpret_cast<const unsigned char *>(args.b);
        int it = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int row = min(tile * 16 + (lane & 15), K - 1);
            const unsigned char *src_row = bbytes + static_cast<size_t>(row) * (N / 2);
            for (int c = 0; c < cpt; ++c, ++it) {
                const int s = it % S;
                const uint32_t phase = (it / S) & 1;
                const int groups = min(SK_CG, G - c * SK_CG);
                mbar_wait(empty_s + 8 * s, phase ^ 1);
                if (lane == 0) mbar_expect_tx(full_s + 8 * s, 16u * groups * 64u);
                __syncwarp();
                if (lane < 16)
                    bulk_g2s(ring_s + s * SK_STAGE_BYTES + lane * SK_ROW_STRIDE, src_row + static_cast<size_t>(c) * SK_ROW_BYTES,
                             groups * 64u, full_s + 8 * s);
            }
        }
        return;
    }

    // ---------------------------------- consumers ----------------------------------
    griddep_wait();  // activations (and the residual) come from the previous kernel
    const int ctid = threadIdx.x;  // 0 .. 255
    const T *p0 = static_cast<const T *>(args.p0) + static_cast<size_t>(pass) * args.rows_per_pass * args.lda;
    const T *p1 = args.prologue == PRO_SWIGLU
                      ? static_cast<const T *>(args.p1) + static_cast<size_t>(pass) * args.rows_per_pass * args.lda
                      : static_cast<const T *>(args.p1);
    if (args.prologue == PRO_RMSNORM) {
        if (ctid < 32) rowstat[ctid] = 0.f;
        consumer_barrier();
        const int total = Mp * words;
        for (int base = (ctid & ~31); base < total; base += SK_CONSUMERS * 32) {
            const int idx = base + lane;
            float part = 0.f;
            int m = 0;
            if (idx < total) {
                m = idx / words;
                const int c = idx - m * words;
                const uint4 raw = *reinterpret_cast<const uint4 *>(p0 + static_cast<size_t>(m) * args.lda + c * 8);
                const float2 f0 = unpack2<T>(raw.x), f1 = unpack2<T>(raw.y), f2 = unpack2<T>(raw.z), f3 = unpack2<T>(raw.w);
                part = f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y + f2.x * f2.x + f2.y * f2.y + f3.x * f3.x + f3.y * f3.y;
            }
            // a warp's 32 chunks belong to at most two rows (words % 16 == 0): reduce per half-warp
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (idx < total && (lane & 15) == 0) atomicAdd(&rowstat[m], part);
        }
        consumer_barrier();
    }
    {
        const int total = Mp * words;
        for (int base = (ctid & ~31); base < total; base += SK_CONSUMERS * 32) {
            const int idx = base + lane;
            float part = 0.f;
            int m = 0, c = 0;
            if (idx < total) {
                m = idx / words;
                c = idx - m * words;
                uint4 raw = *reinterpret_cast<const uint4 *>(p0 + static_cast<size_t>(m) * args.lda + c * 8);
                if (args.prologue != PRO_NONE) {
                    const uint4 aux = *reinterpret_cast<const uint4 *>(
                        args.prologue == PRO_SWIGLU ? p1 + static_cast<size_t>(m) * args.lda + c * 8 : p1 + c * 8);
                    const uint32_t xin[4] = {raw.x, raw.y, raw.z, raw.w};
                    const uint32_t yin[4] = {aux.x, aux.y, aux.z, aux.w};
                    uint32_t o[4];
                    const float inv = args.prologue == PRO_RMSNORM ? rsqrtf(rowstat[m] / static_cast<float>(N) + args.eps) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 xv = unpack2<T>(xin[i]), yv = unpack2<T>(yin[i]);
                        float r0, r1;
                        if (args.prologue == PRO_RMSNORM) {
                            r0 = xv.x * inv * yv.x;
                            r1 = xv.y * inv * yv.y;
                        } else {
                            r0 = (xv.x / (1.0f + expf(-xv.x))) * yv.x;
                            r1 = (xv.y / (1.0f + expf(-xv.y))) * yv.y;
                        }
                        o[i] = pack2<T>(r0, r1);
                    }
                    raw = make_uint4(o[0], o[1], o[2], o[3]);
                }
                uint4 p;
                p.x = __byte_perm(raw.x, raw.z, 0x5410);
                p.y = __byte_perm(raw.x, raw.z, 0x7632);
                p.z = __byte_perm(raw.y, raw.w, 0x5410);
                p.w = __byte_perm(raw.y, raw.w, 0x7632);
                const int pos = (c & ~15) | ((c & 3) << 2) | ((c >> 2) & 3);
                act[static_cast<size_t>(pos) * Mp + m] = p;
                const float2 f0 = unpack2<T>(raw.x), f1 = unpack2<T>(raw.y), f2 = unpack2<T>(raw.z), f3 = unpack2<T>(raw.w);
                part = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
            }
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (idx < total && (c & 15) == 0) asum[(c >> 4) * Mp + m] = part;
        }
    }
    consumer_barrier();

    const int g = lane >> 2, t = lane & 3;
    const int gi = warp & (SK_CG - 1);  // group of the stage this warp owns
    const int par = warp >> 2;          // stages with (it & 1) == par are this warp's
    const T *scales = static_cast<const T *>(args.scales);
    const T *biases = static_cast<const T *>(args.biases);
    T *out = static_cast<T *>(args.out) + static_cast<size_t>(pass) * args.rows_per_pass * K;
    const T *res = args.epilogue == EPI_RESIDUAL ? static_cast<const T *>(args.residual) + static_cast<size_t>(pass) * args.rows_per_pass * K
                                                 : nullptr;

    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int row0 = min(tile * 16 + g, K - 1);
        const int row1 = min(tile * 16 + g + 8, K - 1);
        float acc[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
        for (int c = 0; c < cpt; ++c, ++it) {
            if ((it & 1) != par) continue;
            const int s = it % S;
            const uint32_t phase = (it / S) & 1;
            const int u = c * SK_CG + gi;
            float s0 = 0.f, s1 = 0.f, c0 = 0.f, c1 = 0.f;
            if (u < G) {
                s0 = to_f(scales[static_cast<size_t>(row0) * G + u]);
                s1 = to_f(scales[static_cast<size_t>(row1) * G + u]);
                c0 = to_f(biases[static_cast<size_t>(row0) * G + u]) - Mma<T>::OFFSET * s0;
                c1 = to_f(biases[static_cast<size_t>(row1) * G + u]) - Mma<T>::OFFSET * s1;
            }
            mbar_wait(full_s + 8 * s, phase);
            if (u < G) {
                const unsigned char *stage = ring + s * SK_STAGE_BYTES + gi * 64 + t * 16;
                const uint4 w0 = *reinterpret_cast<const uint4 *>(stage + g * SK_ROW_STRIDE);
                const uint4 w1 = *reinterpret_cast<const uint4 *>(stage + (g + 8) * SK_ROW_STRIDE);
                float d[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f;
                const uint32_t x0[4] = {w0.x, w0.y, w0.z, w0.w};
                const uint32_t x1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    constexpr uint32_t MASK = 0x000F000Fu;
                    const uint32_t a0 = (x0[j] & MASK) | Mma<T>::MAGIC;
                    const uint32_t a1 = ((x0[j] >> 4) & MASK) | Mma<T>::MAGIC;
                    const uint32_t a2 = ((x0[j] >> 8) & MASK) | Mma<T>::MAGIC;
                    const uint32_t a3 = ((x0[j] >> 12) & MASK) | Mma<T>::MAGIC;
                    const uint32_t b0 = (x1[j] & MASK) | Mma<T>::MAGIC;
                    const uint32_t b1 = ((x1[j] >> 4) & MASK) | Mma<T>::MAGIC;
                    const uint32_t b2 = ((x1[j] >> 8) & MASK) | Mma<T>::MAGIC;
                    const uint32_t b3 = ((x1[j] >> 12) & MASK) | Mma<T>::MAGIC;
                    const int chunk = 16 * u + 4 * j + t;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int col = mt * 8 + g;
                        uint4 bf = make_uint4(0u, 0u, 0u, 0u);
                        if (col < Mp) bf = act[static_cast<size_t>(chunk) * Mp + col];
                        Mma<T>::mma(d[mt], a0, b0, a1, b1, bf.x, bf.y);
                        Mma<T>::mma(d[mt], a2, b2, a3, b3, bf.z, bf.w);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int m0 = mt * 8 + 2 * t;
                    const float as0 = m0 < Mp ? asum[u * Mp + m0] : 0.f;
                    const float as1 = m0 + 1 < Mp ? asum[u * Mp + m0 + 1] : 0.f;
                    acc[mt][0] += s0 * d[mt][0] + c0 * as0;
                    acc[mt][1] += s0 * d[mt][1] + c0 * as1;
                    acc[mt][2] += s1 * d[mt][2] + c1 * as0;
                    acc[mt][3] += s1 * d[mt][3] + c1 * as1;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_s + 8 * s);
        }
        // ---- cross-warp reduction of the tile (double-buffered scratch, one barrier per tile)
        float *buf = red + static_cast<size_t>(tile & 1) * SK_CONSUMERS * 16 * 8 * MT;
        float *wred = buf + static_cast<size_t>(warp) * 16 * 8 * MT;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            wred[g * 8 * MT + mt * 8 + 2 * t] = acc[mt][0];
            wred[g * 8 * MT + mt * 8 + 2 * t + 1] = acc[mt][1];
            wred[(g + 8) * 8 * MT + mt * 8 + 2 * t] = acc[mt][2];
            wred[(g + 8) * 8 * MT + mt * 8 + 2 * t + 1] = acc[mt][3];
        }
        consumer_barrier();
        for (int o = ctid; o < 16 * 8 * MT; o += SK_CONSUMERS * 32) {
            const int m = o >> 4, r = o & 15;  // consecutive threads -> consecutive output features
            const int k = tile * 16 + r;
            if (m < Mp && k < K) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < SK_CONSUMERS; ++w) v += buf[(w * 16 + r) * 8 * MT + m];
                T vb = from_f<T>(v);
                if (res != nullptr) vb = from_f<T>(to_f(res[static_cast<size_t>(m) * K + k]) + to_f(vb));
                out[static_cast<size_t>(m) * K + k] = vb;
            }
        }
    }
}

static bool g_use_pdl = false;
void set_use_pdl(bool on) { g_use_pdl = on; }
bool use_pdl() { return g_use_pdl; }

static size_t stream2_smem_bytes(int N, int Mp, int MT, int stages) {
    size_t bytes = static_cast<size_t>(stages) * SK_STAGE_BYTES;
    bytes += (2 * stages * 8 + 15) & ~15;
    bytes += static_cast<size_t>(N / 8) * Mp * 16;
    bytes += static_cast<size_t>(((N / 128) * Mp + 3) & ~3) * 4;
    bytes += 32 * 4;
    bytes += 2ull * SK_CONSUMERS * 16 * 8 * MT * 4;
    return bytes;
}

template <typename T, int MT>
static int stream2_launch(StreamArgs args, cudaStream_t st) {
    const int Mp = args.rows_per_pass < args.M ? args.rows_per_pass : args.M;
    const size_t fixed = stream2_smem_bytes(args.N, Mp, MT, 0);
    // ring depth: as deep as fits next to a second resident CTA, within [4, 16] stages
    const size_t per_cta_budget = fixed > 100 * 1024 ? 220 * 1024 : 110 * 1024;
    int stages = static_cast<int>((per_cta_budget - fixed) / SK_STAGE_BYTES);
    stages = stages > 16 ? 16 : stages;
    if (stages < 2) return fail(TL_EINVAL, "quantized_matmul: activations do not fit in shared memory (N=%d, rows=%d)", args.N, Mp);
    args.stages = stages;
    const size_t smem = stream2_smem_bytes(args.N, Mp, MT, stages);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(w4a16_stream2_kernel<T, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return fail(TL_ECUDA, "quantized_matmul: cannot raise shared memory limit: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int tiles = ceil_div(args.K, 16);
    const int ctas_per_sm = smem > 110 * 1024 ? 1 : 2;
    const int cap = sm_count() * ctas_per_sm;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles < cap ? tiles : cap, ceil_div(args.M, args.rows_per_pass));
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, w4a16_stream2_kernel<T, MT>, args);
    if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_stream2: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("w4a16_stream2");
    return TL_OK;
}

template <typename T>
static int stream2_t(StreamArgs args, cudaStream_t st) {
    if (!aligned16(args.p0) || !aligned16(args.b) || (args.p1 && !aligned16(args.p1)) || (args.lda % 8) != 0)
        return fail(TL_EINVAL, "quantized_matmul: operands must be 16-byte aligned");
    const size_t budget = 150 * 1024;
    int rpp = 8;
    if (args.M > 16 && static_cast<size_t>(args.N) * 32 * 2 <= budget)
        rpp = 32;
    else if (args.M > 8 && static_cast<size_t>(args.N) * 16 * 2 <= budget)
        rpp = 16;
    args.rows_per_pass = rpp;
    if (rpp == 32) return stream2_launch<T, 4>(args, st);
    if (rpp == 16) return stream2_launch<T, 2>(args, st);
    return stream2_launch<T, 1>(args, st);
}

int launch_w4a16_fused(const void *scales, const void *biases, const void *b, void *out, const void *p0, const void *p1,
                       const void *residual, int M, int N, int K, int lda, int prologue, int epilogue, float eps, int dtype,
                       cudaStream_t st) {
    if (M == 0 || K == 0) return TL_OK;
    StreamArgs args{};
    args.scales = scales, args.biases = biases, args.b = static_cast<const uint32_t *>(b), args.out = out;
    args.p0 = p0, args.p1 = p1, args.residual = residual;
    args.M = M, args.N = N, args.K = K, args.lda = lda;
    args.prologue = prologue, args.epilogue = epilogue, args.eps = eps;
    switch (dtype) {
        case TL_F16: return stream2_t<__half>(args, st);
        case TL_BF16: return stream2_t<__nv_bfloat16>(args, st);
    }
    return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
}

// ---------------------------------------------------------------------------
// Scalar control kernel: one thread per output, the literal arithmetic of
// quantized_matmul_vanilla_w4a16_g128 (quantized_matmul.metal:8-56):
//   sum += (float(code) * scale + bias) * float(a), codes in nibble order.
template <typename T>
__global__ void w4a16_vanilla_kernel(const T *__restrict__ scales, const T *__restrict__ biases,
                                     const T *__restrict__ a, const uint32_t *__restrict__ b, T *__restrict__ out,
                                     int M, int N, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (k >= K || i >= M) return;
    const int words = N / 8;
    const int G = N / 128;
    const uint32_t *brow = b + static_cast<size_t>(k) * words;
    const T *arow = a + static_cast<size_t>(i) * N;
    float sum = 0.f;
    for (int grp = 0; grp < G; ++grp) {
        const float s = to_f(scales[static_cast<size_t>(k) * G + grp]);
        const float c = to_f(biases[static_cast<size_t>(k) * G + grp]);
        for (int w = 0; w < 16; ++w) {
            const uint32_t packed = brow[grp * 16 + w];
            const T *av = arow + grp * 128 + w * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sum += (static_cast<float>((packed >> (4 * j)) & 0xFu) * s + c) * to_f(av[j]);
        }
    }
    out[static_cast<size_t>(i) * K + k] = from_f<T>(sum);
}

int launch_w4a16_vanilla(const void *scales, const void *biases, const void *a, const void *b, void *out, int M,
                         int N, int K, int dtype, cudaStream_t st) {
    if (M == 0 || K == 0) return TL_OK;
    dim3 grid(ceil_div(K, 128), M);
    if (dtype == TL_BF16)
        w4a16_vanilla_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(
            static_cast<const __nv_bfloat16 *>(scales), static_cast<const __nv_bfloat16 *>(biases),
            static_cast<const __nv_bfloat16 *>(a), static_cast<const uint32_t *>(b), static_cast<__nv_bfloat16 *>(out),
            M, N, K);
    else if (dtype == TL_F16)
        w4a16_vanilla_kernel<__half><<<grid, 128, 0, st>>>(static_cast<const __half *>(scales),
                                                            static_cast<const __half *>(biases),
                                                            static_cast<const __half *>(a),
                                                            static_cast<const uint32_t *>(b),
                                                            static_cast<__half *>(out), M, N, K);
    else
        return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
    TL_LAUNCH_CHECK("w4a16_vanilla");
    return TL_OK;
}

}  // namespace tl
