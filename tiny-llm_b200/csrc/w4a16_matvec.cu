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
// v3: per-warp cp.async rings.
//
// Measured on B200 (profiles/r01_kbench_*): v1 (register prefetch, 2 KB in flight
// per warp) reached 31 % of HBM peak on the tied head; v2 (one producer warp issuing
// 256-byte cp.async.bulk rows into a CTA ring) was slower still - small bulk copies
// are issue-limited.  v3 gives every warp a PRIVATE ring of RING 1-KiB slots (one
// 16-row x 128-column weight group each) filled with 16-byte cp.async: 8 KiB in
// flight per warp, 128+ KiB per SM, no mbarriers and no cross-warp traffic on the
// load path.  A tile's reduction dimension is split over a team of TW warps
// (TW = 1..8, chosen so that small projections still put >= 16 warps on every SM);
// only that team synchronises (named barrier) to combine its partial sums.
//
// Weights never depend on the previous kernel: each warp fills its ring BEFORE
// griddepcontrol.wait, so with programmatic dependent launch the HBM stream of
// kernel n+1 starts while kernel n drains.
//
// Optional fusions that keep every rounding point of the unfused call sequence
// (so results match rms_norm -> matvec, swiglu -> matvec, matvec -> add):
//   prologue RMSNORM : a = T(x * rsqrt(mean(x^2)+eps) * w)   (week2_kernels.metal:41-47)
//   prologue SWIGLU  : a = T(g / (1 + exp(-g)) * u)          (week2_kernels.metal:115-116)
//   epilogue RESIDUAL: out = T(float(res) + float(T(acc)))   (qwen3_week3.py:204-206)
constexpr int S3_WARPS = 8;
constexpr int S3_THREADS = S3_WARPS * 32;
constexpr int S3_RING = 8;              // slots per warp
constexpr int S3_CODE_BYTES = 1024;     // 16 rows x 64 B of packed codes
constexpr int S3_SLOT_BYTES = 1024 + 128;  // + 16 scale words + 16 bias words (4 B each, see issue())

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
    int rows_per_pass, team_warps;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <typename T, int MT>
__global__ void __launch_bounds__(S3_THREADS) w4a16_stream3_kernel(const StreamArgs args) {
    extern __shared__ __align__(128) unsigned char smem3_raw[];
    const int N = args.N, K = args.K;
    const int pass = blockIdx.y;
    const int Mp = min(args.rows_per_pass, args.M - pass * args.rows_per_pass);
    const int words = N / 8;
    const int G = N / 128;
    const int TW = args.team_warps;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int team = warp / TW;
    const int wit = warp - team * TW;
    const int teams = S3_WARPS / TW;
    const int g = lane >> 2, t = lane & 3;

    // shared layout: rings [8 warps][RING][1 KiB] | act | asum | row stats | red
    unsigned char *ring = smem3_raw + static_cast<size_t>(warp) * S3_RING * S3_SLOT_BYTES;
    uint4 *act = reinterpret_cast<uint4 *>(smem3_raw + static_cast<size_t>(S3_WARPS) * S3_RING * S3_SLOT_BYTES);
    float *asum = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(act) + static_cast<size_t>(words) * Mp * 16);
    float *rowstat = asum + ((G * Mp + 3) & ~3);
    float *red = rowstat + 32;
    const uint32_t ring_s = smem_u32(ring);

    griddep_launch();

    // ---- this warp's work list: tiles of its team, groups u = wit, wit+TW, ... of each tile
    const int tiles = (K + 15) / 16;
    const int team_id = blockIdx.x * teams + team;
    const int team_count = gridDim.x * teams;
    const int gpw = (G - wit + TW - 1) / TW;  // groups of one tile that belong to this warp
    const int my_tiles = team_id < tiles ? (tiles - team_id + team_count - 1) / team_count : 0;
    const int total_items = my_tiles * gpw;
    const unsigned char *bbytes = reinterpret_cast<const unsigned char *>(args.b);
    const int crow = lane >> 1;         // row of the tile this lane copies
    const int chalf = (lane & 1) * 32;  // which 32 bytes of the row's 64

    // Scales/biases ride in the same ring: a 2-byte value cannot be cp.async'ed on its own, so the
    // aligned 4-byte word that contains it is copied and the consumer picks the half by parity.
    const unsigned char *sbytes = reinterpret_cast<const unsigned char *>(args.scales);
    const unsigned char *cbytes = reinterpret_cast<const unsigned char *>(args.biases);
    const long long sb_elems = static_cast<long long>(K) * G;
    auto issue = [&](int item) {  // item -> (tile, group); one commit group per item, always
        if (item < total_items) {
            const int tile = team_id + (item / gpw) * team_count;
            const int u = wit + (item % gpw) * TW;
            const int row = min(tile * 16 + crow, K - 1);
            const unsigned char *src = bbytes + static_cast<size_t>(row) * (N / 2) + u * 64 + chalf;
            const uint32_t slot = ring_s + (item % S3_RING) * S3_SLOT_BYTES;
            const uint32_t dst = slot + crow * 64 + chalf;
            cp_async16(dst, src);
            cp_async16(dst + 16, src + 16);
            const int prow = min(tile * 16 + (lane & 15), K - 1);
            const long long e = static_cast<long long>(prow) * G + u;
            const unsigned char *table = lane < 16 ? sbytes : cbytes;
            const uint32_t pdst = slot + S3_CODE_BYTES + lane * 4;
            if ((e | 1) < sb_elems) {
                cp_async4(pdst, table + ((e >> 1) << 2));
            } else {  // last element of an odd-sized table: its pair would cross the end, copy synchronously
                const uint16_t v = *reinterpret_cast<const uint16_t *>(table + e * 2);
                *reinterpret_cast<uint32_t *>(ring + (item % S3_RING) * S3_SLOT_BYTES + S3_CODE_BYTES + lane * 4) =
                    (e & 1) ? (static_cast<uint32_t>(v) << 16) : static_cast<uint32_t>(v);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < S3_RING - 1; ++i) issue(i);  // fill the ring before touching activations

    griddep_wait();  // activations (and the residual) come from the previous kernel
    const T *p0 = static_cast<const T *>(args.p0) + static_cast<size_t>(pass) * args.rows_per_pass * args.lda;
    const T *p1 = args.prologue == PRO_SWIGLU
                      ? static_cast<const T *>(args.p1) + static_cast<size_t>(pass) * args.rows_per_pass * args.lda
                      : static_cast<const T *>(args.p1);
    // ---- stage activations: each 8-element chunk is loaded ONCE (kept in registers across the
    // row-statistics barrier when the tile is small enough), transformed by the prologue, rounded
    // to T, permuted into MMA-fragment order and written to shared memory with its group sum.
    {
        constexpr int CACHE = 4;
        const int total = Mp * words;
        const bool cached = total <= CACHE * S3_THREADS;
        const bool rms = args.prologue == PRO_RMSNORM;
        uint4 held[CACHE];
        auto chunk_src = [&](int idx, int &m, int &c) -> const T * {
            m = idx / words;
            c = idx - m * words;
            return p0 + static_cast<size_t>(m) * args.lda + c * 8;
        };
        auto half_warp_sum = [](float v) {
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            return v;
        };
        auto square_sum = [](const uint4 &raw) {
            const float2 f0 = unpack2<T>(raw.x), f1 = unpack2<T>(raw.y), f2 = unpack2<T>(raw.z), f3 = unpack2<T>(raw.w);
            return f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y + f2.x * f2.x + f2.y * f2.y + f3.x * f3.x + f3.y * f3.y;
        };
        auto emit = [&](int idx, uint4 raw) {  // idx may be >= total (lane padding): contributes nothing
            float part = 0.f;
            int m = 0, c = 0;
            if (idx < total) {
                const T *src = chunk_src(idx, m, c);
                if (args.prologue != PRO_NONE) {
                    const uint4 aux = *reinterpret_cast<const uint4 *>(
                        args.prologue == PRO_SWIGLU ? p1 + (src - p0) : p1 + c * 8);
                    const uint32_t xin[4] = {raw.x, raw.y, raw.z, raw.w};
                    const uint32_t yin[4] = {aux.x, aux.y, aux.z, aux.w};
                    uint32_t o[4];
                    const float inv = rms ? rsqrtf(rowstat[m] / static_cast<float>(N) + args.eps) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 xv = unpack2<T>(xin[i]), yv = unpack2<T>(yin[i]);
                        float r0, r1;
                        if (rms) {
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
                // 4x4 transpose of the chunk order inside a group: the four lanes of an MMA group
                // then read 64 contiguous bytes per sub-step (no bank conflicts).
                const int pos = (c & ~15) | ((c & 3) << 2) | ((c >> 2) & 3);
                act[static_cast<size_t>(pos) * Mp + m] = p;
                const float2 f0 = unpack2<T>(raw.x), f1 = unpack2<T>(raw.y), f2 = unpack2<T>(raw.z), f3 = unpack2<T>(raw.w);
                part = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
            }
            part = half_warp_sum(part);  // 16 consecutive chunks (one group) live in 16 consecutive lanes
            if (idx < total && (c & 15) == 0) asum[(c >> 4) * Mp + m] = part;
        };
        const int base0 = threadIdx.x & ~31;
        if (cached) {
#pragma unroll
            for (int j = 0; j < CACHE; ++j) {
                const int idx = base0 + j * S3_THREADS + lane;
                held[j] = make_uint4(0u, 0u, 0u, 0u);
                if (idx < total) {
                    int m, c;
                    held[j] = *reinterpret_cast<const uint4 *>(chunk_src(idx, m, c));
                }
            }
        }
        if (rms) {
            if (threadIdx.x < 32) rowstat[threadIdx.x] = 0.f;
            __syncthreads();
            // a warp's 32 chunks belong to at most two rows (words % 16 == 0): reduce per half-warp
            if (cached) {
#pragma unroll
                for (int j = 0; j < CACHE; ++j) {
                    const int idx = base0 + j * S3_THREADS + lane;
                    if (base0 + j * S3_THREADS < total) {
                        const float part = half_warp_sum(idx < total ? square_sum(held[j]) : 0.f);
                        if (idx < total && (lane & 15) == 0) atomicAdd(&rowstat[idx / words], part);
                    }
                }
            } else {
                for (int base = base0; base < total; base += S3_THREADS) {
                    const int idx = base + lane;
                    float part = 0.f;
                    if (idx < total) {
                        int m, c;
                        part = square_sum(*reinterpret_cast<const uint4 *>(chunk_src(idx, m, c)));
                    }
                    part = half_warp_sum(part);
                    if (idx < total && (lane & 15) == 0) atomicAdd(&rowstat[idx / words], part);
                }
            }
            __syncthreads();
        }
        if (cached) {
#pragma unroll
            for (int j = 0; j < CACHE; ++j)
                if (base0 + j * S3_THREADS < total) emit(base0 + j * S3_THREADS + lane, held[j]);
        } else {
            for (int base = base0; base < total; base += S3_THREADS) {
                const int idx = base + lane;
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (idx < total) {
                    int m, c;
                    raw = *reinterpret_cast<const uint4 *>(chunk_src(idx, m, c));
                }
                emit(idx, raw);
            }
        }
    }
    __syncthreads();

    T *out = static_cast<T *>(args.out) + static_cast<size_t>(pass) * args.rows_per_pass * K;
    const T *res = args.epilogue == EPI_RESIDUAL ? static_cast<const T *>(args.residual) + static_cast<size_t>(pass) * args.rows_per_pass * K
                                                 : nullptr;
    float *team_red = red + static_cast<size_t>(team) * TW * 16 * 8 * MT;
    const int team_threads = TW * 32;
    const int ttid = threadIdx.x - team * team_threads;

    int item = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int tile = team_id + ti * team_count;
        const int row0 = min(tile * 16 + g, K - 1);
        const int row1 = min(tile * 16 + g + 8, K - 1);
        float acc[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
        for (int u = wit; u < G; u += TW, ++item) {
            cp_async_wait<S3_RING - 2>();  // the oldest outstanding group (this item) has landed
            __syncwarp();
            const unsigned char *slot0 = ring + (item % S3_RING) * S3_SLOT_BYTES;
            const unsigned char *slot = slot0 + t * 16;
            const uint4 w0 = *reinterpret_cast<const uint4 *>(slot + g * 64);
            const uint4 w1 = *reinterpret_cast<const uint4 *>(slot + (g + 8) * 64);
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(slot0 + S3_CODE_BYTES);
            const int par0 = static_cast<int>((static_cast<long long>(row0) * G + u) & 1);
            const int par1 = static_cast<int>((static_cast<long long>(row1) * G + u) & 1);
            const float2 sp0 = unpack2<T>(sw[g]), sp1 = unpack2<T>(sw[g + 8]);
            const float2 cp0 = unpack2<T>(sw[16 + g]), cp1 = unpack2<T>(sw[16 + g + 8]);
            const float s0 = par0 ? sp0.y : sp0.x, s1 = par1 ? sp1.y : sp1.x;
            const float c0 = (par0 ? cp0.y : cp0.x) - Mma<T>::OFFSET * s0;
            const float c1 = (par1 ? cp1.y : cp1.x) - Mma<T>::OFFSET * s1;
            __syncwarp();                   // every lane has read the slot the next issue may overwrite ...
            issue(item + S3_RING - 1);      // ... which is slot (item - 1) % RING, consumed one iteration ago
            // The MMAs of one group are independent instructions (two accumulator sets per column
            // tile, summed afterwards): legacy mma.sync has a long latency on sm_100 and a dependent
            // chain of eight is ~4x slower than back-to-back issue.
            constexpr int CH = MT == 1 ? 8 : (MT == 2 ? 4 : 2);  // independent accumulator sets per column tile
            float dd[MT][CH][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int c = 0; c < CH; ++c) dd[mt][c][0] = dd[mt][c][1] = dd[mt][c][2] = dd[mt][c][3] = 0.f;
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
                    Mma<T>::mma(dd[mt][(2 * j) % CH], a0, b0, a1, b1, bf.x, bf.y);
                    Mma<T>::mma(dd[mt][(2 * j + 1) % CH], a2, b2, a3, b3, bf.z, bf.w);
                }
            }
            float d[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float sum = 0.f;
#pragma unroll
                    for (int k = 0; k < CH; ++k) sum += dd[mt][k][c];
                    d[mt][c] = sum;
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
        // ---- combine the team's partial sums and store 16 x (8*MT) outputs
        float *wred = team_red + static_cast<size_t>(wit) * 16 * 8 * MT;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            wred[g * 8 * MT + mt * 8 + 2 * t] = acc[mt][0];
            wred[g * 8 * MT + mt * 8 + 2 * t + 1] = acc[mt][1];
            wred[(g + 8) * 8 * MT + mt * 8 + 2 * t] = acc[mt][2];
            wred[(g + 8) * 8 * MT + mt * 8 + 2 * t + 1] = acc[mt][3];
        }
        if (TW == 1)
            __syncwarp();
        else
            named_barrier(1 + team, team_threads);
        for (int o = ttid; o < 16 * 8 * MT; o += team_threads) {
            const int m = o >> 4, r = o & 15;  // consecutive threads -> consecutive output features
            const int k = tile * 16 + r;
            if (m < Mp && k < K) {
                float v = 0.f;
                for (int w = 0; w < TW; ++w) v += team_red[(w * 16 + r) * 8 * MT + m];
                T vb = from_f<T>(v);
                if (res != nullptr) vb = from_f<T>(to_f(res[static_cast<size_t>(m) * K + k]) + to_f(vb));
                out[static_cast<size_t>(m) * K + k] = vb;
            }
        }
        if (TW == 1)
            __syncwarp();
        else
            named_barrier(1 + team, team_threads);
    }
    cp_async_wait<0>();
}

static bool g_use_pdl = false;
void set_use_pdl(bool on) { g_use_pdl = on; }
bool use_pdl() { return g_use_pdl; }

static size_t stream3_smem_bytes(int N, int Mp, int MT) {
    size_t bytes = static_cast<size_t>(S3_WARPS) * S3_RING * S3_SLOT_BYTES;
    bytes += static_cast<size_t>(N / 8) * Mp * 16;
    bytes += static_cast<size_t>(((N / 128) * Mp + 3) & ~3) * 4;
    bytes += 32 * 4;
    bytes += static_cast<size_t>(S3_WARPS) * 16 * 8 * MT * 4;
    return bytes;
}

template <typename T, int MT>
static int stream3_launch(StreamArgs args, cudaStream_t st) {
    const int Mp = args.rows_per_pass < args.M ? args.rows_per_pass : args.M;
    const size_t smem = stream3_smem_bytes(args.N, Mp, MT);
    if (smem > 224 * 1024)
        return fail(TL_EINVAL, "quantized_matmul: activations do not fit in shared memory (N=%d, rows=%d)", args.N, Mp);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(w4a16_stream3_kernel<T, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return fail(TL_ECUDA, "quantized_matmul: cannot raise shared memory limit: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int tiles = ceil_div(args.K, 16);
    const int G = args.N / 128;
    const int sms = sm_count();
    // team width: split a tile's reduction over TW warps until the launch offers ~16 warps per SM
    int tw = 1;
    while (tw < S3_WARPS && tiles * tw < sms * 16 && tw * 2 <= G) tw *= 2;
    args.team_warps = tw;
    const int teams_per_cta = S3_WARPS / tw;
    int ctas_per_sm = static_cast<int>((224 * 1024) / (smem + 1024));
    ctas_per_sm = ctas_per_sm > 3 ? 3 : (ctas_per_sm < 1 ? 1 : ctas_per_sm);
    const int want = ceil_div(tiles, teams_per_cta);
    const int cap = sms * ctas_per_sm;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(want < cap ? want : cap, ceil_div(args.M, args.rows_per_pass));
    cfg.blockDim = dim3(S3_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, w4a16_stream3_kernel<T, MT>, args);
    if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_stream3: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("w4a16_stream3");
    return TL_OK;
}

template <typename T>
static int stream3_t(StreamArgs args, cudaStream_t st) {
    if (!aligned16(args.p0) || !aligned16(args.b) || (args.p1 && !aligned16(args.p1)) || (args.lda % 8) != 0)
        return fail(TL_EINVAL, "quantized_matmul: operands must be 16-byte aligned");
    // rows of `a` handled per pass: as many as fit in shared memory next to the weight rings
    const size_t budget = 224 * 1024 - stream3_smem_bytes(args.N, 0, 4) - 2048;
    const size_t per_row = static_cast<size_t>(args.N) * 2 + static_cast<size_t>(args.N / 128) * 4;
    const int fit = static_cast<int>(budget / per_row);
    if (fit < 1) return fail(TL_EINVAL, "quantized_matmul: reduction length %d does not fit in shared memory", args.N);
    int rpp = fit < 8 ? fit : 8;
    if (args.M > 16 && fit >= 32)
        rpp = 32;
    else if (args.M > 8 && fit >= 16)
        rpp = 16;
    args.rows_per_pass = rpp;
    if (rpp == 32) return stream3_launch<T, 4>(args, st);
    if (rpp == 16) return stream3_launch<T, 2>(args, st);
    return stream3_launch<T, 1>(args, st);
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
        case TL_F16: return stream3_t<__half>(args, st);
        case TL_BF16: return stream3_t<__nv_bfloat16>(args, st);
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
