// W4A16 weight-streaming matvec, generation 6: weights staged by TMA into per-warp shared-memory rings, so
// that the kernel fits in 64 registers per thread and TWO 16-warp CTAs - one from each of two consecutive
// launches of a programmatic-dependent-launch chain - share an SM.
//
// Why: generation 5 keeps its weight pipeline in registers (128 regs x 512 threads = one
// CTA per SM), so the CTAs of launch n+1 start only when those of launch n exit, and their first weight bytes,
// their activations (queued behind 128 KiB of weight loads per SM) and their launch latency are all exposed:
// ~2.5 us of an average 7.7 us launch.  The round-2 attempt to co-schedule two 8-warp CTAs halved the warps of
// the (issue-bound) consume phase and lost.  Here both CTAs keep 16 warps: launch n+1 sits beside launch n with
// its 64 KiB ring already filled (the ring is requested BEFORE griddepcontrol.wait), waits, stages its
// activations from an idle load queue and consumes from shared memory at once.
//
// Layout of one unit = 16 rows x 128 bytes (two 128-column groups) = one TMA box with the 128-byte swizzle:
// chunk c (16 B) of row r sits at r * 128 + ((c ^ (r & 7)) << 4).  Lane (g, t) of the consuming warp reads chunks
// 4*it + t of rows g and g + 8 - exactly the register image generation 5 loads from global memory - and feeds the
// same exact-integer tensor-core arithmetic (w4a16_item.cuh).  Scales and biases of the CTA's rows are copied
// to shared memory once (rows x N/128 pairs).  Everything else (row split over CTAs, unit split over warps,
// deterministic warp-order reduction, fused prologue / epilogue) is generation 5's.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"
#include "trace.cuh"
#include "w4a16_item.cuh"

namespace tl {

constexpr int S6_WARPS = 16;
constexpr int S6_THREADS = S6_WARPS * 32;
constexpr int S6_SLOTS = 2;            // units per warp in flight: 16 warps x 2 x 2 KiB = 64 KiB per CTA
constexpr int S6_UNIT_BYTES = 2048;
constexpr int S6_U = 2;                // 128-column groups per unit
enum { S6_PRO_NONE = W4_PRO_NONE, S6_PRO_RMSNORM = W4_PRO_RMSNORM, S6_PRO_SWIGLU = W4_PRO_SWIGLU };
enum { S6_EPI_NONE = 0, S6_EPI_RESIDUAL = 1, S6_EPI_SWIGLU_PAIRS = 2 };

struct Stream6Args {
    const void *scales, *biases;
    void *out;
    const void *p0, *p1, *residual;
    int M, N, K, lda;
    int prologue, epilogue;
    float eps;
    int sb_rows;  // rows of scale/bias staged per CTA (>= rows of any CTA)
};

template <typename T, int MP>
__global__ void __launch_bounds__(S6_THREADS, 2) w4a16_stream6_kernel(const __grid_constant__ CUtensorMap tmap_w, const Stream6Args args) {
    constexpr int U = S6_U;
    constexpr int NW = S6_WARPS, NT = S6_THREADS;
    constexpr int MT = 1;
    constexpr int MPA = w4_mpa(MP);
    constexpr int ENTRY = 16 * 8 * MT;
    static_assert(MP <= 8, "generation 6 serves the small-batch decode path");
    extern __shared__ __align__(1024) unsigned char s6_raw[];
    __shared__ int warp_begin[NW + 1];
    const int N = args.N, K = args.K;
    const int Mp = args.M;
    const int words = N / 8, G = N / 128, P = G / U;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    // shared layout: ring [NW][SLOTS][2 KiB] | barriers [NW][SLOTS] | act | asum | scale pairs [rows][G] | bias pairs | entries
    unsigned char *ring = s6_raw + static_cast<size_t>(warp) * S6_SLOTS * S6_UNIT_BYTES;
    const uint32_t bar0 = g_smem_u32(s6_raw + NW * S6_SLOTS * S6_UNIT_BYTES) + warp * S6_SLOTS * 8;
    unsigned char *after = s6_raw + NW * S6_SLOTS * S6_UNIT_BYTES + NW * S6_SLOTS * 8;
    uint4 *act = reinterpret_cast<uint4 *>(after);
    float *asum = reinterpret_cast<float *>(after + static_cast<size_t>(words) * MP * 16);
    __nv_bfloat16 *sc_s = reinterpret_cast<__nv_bfloat16 *>(asum + G * MPA);
    __nv_bfloat16 *bi_s = sc_s + static_cast<size_t>(args.sb_rows) * G;
    float *entries = reinterpret_cast<float *>(bi_s + static_cast<size_t>(args.sb_rows) * G);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    TL_TRACE_STAMP(10);

    // ---- rows of this CTA (whole 16-row chunks) and units of this warp (as in generation 5)
    const unsigned all = static_cast<unsigned>(K + 15) >> 4;
    const int c0 = static_cast<int>(all * blockIdx.x / gridDim.x), c1 = static_cast<int>(all * (blockIdx.x + 1) / gridDim.x);
    const int chunks = c1 - c0;
    const int r0 = c0 * 16, r1 = min(K, c1 * 16);
    const unsigned units = static_cast<unsigned>(chunks) * P;
    const int begin = static_cast<int>(units * warp / NW), end = static_cast<int>(units * (warp + 1) / NW);
    if (threadIdx.x <= NW) warp_begin[threadIdx.x] = static_cast<int>(units * threadIdx.x / NW);

    // ---- this warp's ring: lane 0 owns the barriers and issues the TMA boxes
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S6_SLOTS; ++s) g_mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (warp == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    }
    __syncwarp();
    auto request = [&](int i) {  // unit i of this warp -> slot i % SLOTS   (lane 0 only)
        const int ch = i / P, u = i - ch * P;
        const int s = (i - begin) % S6_SLOTS;  // the consumer's slot rule
        g_mbar_expect_tx(bar0 + 8 * s, S6_UNIT_BYTES);  // rows past K are zero-filled by the TMA unit and still counted
        g_tma_load_2d(g_smem_u32(ring + s * S6_UNIT_BYTES), &tmap_w, u * (U * 64), r0 + ch * 16, bar0 + 8 * s);
    };
    if (lane == 0)
        for (int i = begin; i < end && i < begin + S6_SLOTS; ++i) request(i);  // in flight before the activations exist
    // scales / biases of the CTA's rows (weights: independent of the previous kernel)
    {
        const T *sc = static_cast<const T *>(args.scales), *bi = static_cast<const T *>(args.biases);
        const int rows = chunks * 16;
        for (int i = threadIdx.x; i < rows * G; i += NT) {
            const int r = i / G, gi = i - r * G;
            const size_t at = static_cast<size_t>(min(r0 + r, K - 1)) * G + gi;
            reinterpret_cast<T *>(sc_s)[i] = sc[at];
            reinterpret_cast<T *>(bi_s)[i] = bi[at];
        }
    }
    TL_TRACE_STAMP(11);

    asm volatile("griddepcontrol.wait;" ::: "memory");  // activations (and the residual) come from the previous kernel
    TL_TRACE_STAMP(12);
    const T *p0 = static_cast<const T *>(args.p0);
    const T *p1 = static_cast<const T *>(args.p1);
    T *out = static_cast<T *>(args.out);
    const T *res = args.epilogue == S6_EPI_RESIDUAL ? static_cast<const T *>(args.residual) : nullptr;
    const int outs = chunks * 16 * Mp;
    float res_first = 0.f;
    if (res != nullptr && static_cast<int>(threadIdx.x) < outs) {
        const int m = threadIdx.x / (chunks * 16), rr = threadIdx.x - m * (chunks * 16);
        if (r0 + rr < r1) res_first = to_f(res[static_cast<size_t>(m) * K + r0 + rr]);
    }
    w4_stage<T, MP, NT>(p0, args.lda, p1, args.prologue, N, Mp, args.eps, act, asum, entries);
    TL_TRACE_STAMP(13);

    const uint4 *act0 = w4_act_lane<MP>(act, g, t);
    const float *asum0 = w4_asum_lane<MP>(asum, t);
    int chunk = begin / P;
    int u = begin - chunk * P;
    const uint4 *actp = act0 + u * U * w4_act_group_stride<MP>();
    const float *asump = asum0 + u * U * w4_asum_group_stride<MP>();
    float acc[MT][4] = {{0.f, 0.f, 0.f, 0.f}};
    auto flush = [&]() {
        float *e = entries + static_cast<size_t>(chunk + warp) * ENTRY;
        e[g * 8 + 2 * t] = acc[0][0];
        e[g * 8 + 2 * t + 1] = acc[0][1];
        e[(g + 8) * 8 + 2 * t] = acc[0][2];
        e[(g + 8) * 8 + 2 * t + 1] = acc[0][3];
        acc[0][0] = acc[0][1] = acc[0][2] = acc[0][3] = 0.f;
    };
    const int sw = (g & 7) << 4;  // swizzle term of rows g and g + 8
    for (int i = begin; i < end; ++i) {
        const int s = (i - begin) % S6_SLOTS;
        g_mbar_wait(bar0 + 8 * s, ((i - begin) / S6_SLOTS) & 1);
        W4Unit<U> un;
        const unsigned char *base = ring + s * S6_UNIT_BYTES;
#pragma unroll
        for (int it = 0; it < U; ++it) {
            const int off = ((it * 4 + t) << 4) ^ sw;
            un.w[2 * it] = *reinterpret_cast<const uint4 *>(base + g * 128 + off);
            un.w[2 * it + 1] = *reinterpret_cast<const uint4 *>(base + (g + 8) * 128 + off);
        }
        {  // lanes 0-15: scales of row (lane & 15), lanes 16-31: biases; both groups of the unit in one 32-bit word
            const __nv_bfloat16 *tab = lane < 16 ? sc_s : bi_s;
            un.sb = *reinterpret_cast<const uint32_t *>(tab + static_cast<size_t>(chunk * 16 + (lane & 15)) * G + u * U);
        }
        w4_consume<T, MP, U>(un, actp, asump, g, acc);
        __syncwarp();  // every lane's reads of the slot have completed (their values were just used)
        if (lane == 0 && i + S6_SLOTS < end) request(i + S6_SLOTS);
        actp += U * w4_act_group_stride<MP>();
        asump += U * w4_asum_group_stride<MP>();
        if (++u == P) {
            flush();
            u = 0;
            chunk += 1;
            actp = act0;
            asump = asum0;
        }
    }
    if (u != 0) flush();
    TL_TRACE_STAMP(14);
    __syncthreads();
    TL_TRACE_STAMP(16);

    auto chunk_sum = [&](int ch, int row, int m) {
        float v = 0.f;
        const int lo = ch * P, hi = lo + P;
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (warp_begin[w] < hi && warp_begin[w + 1] > lo && warp_begin[w] < warp_begin[w + 1])
                v += entries[static_cast<size_t>(ch + w) * ENTRY + row * 8 + m];
        return v;
    };
    if (args.epilogue == S6_EPI_SWIGLU_PAIRS) {
        for (int o = threadIdx.x; o < chunks * 8 * Mp; o += NT) {
            const int m = o / (chunks * 8);
            const int rr = o - m * (chunks * 8);
            const int ch = rr >> 3, row = rr & 7;
            const float gate = to_f(from_f<T>(chunk_sum(ch, row, m))), up = to_f(from_f<T>(chunk_sum(ch, row + 8, m)));
            out[static_cast<size_t>(m) * (K / 2) + c0 * 8 + rr] = from_f<T>((gate / (1.0f + expf(-gate))) * up);
        }
        TL_TRACE_STAMP(15);
        return;
    }
    for (int o = threadIdx.x; o < outs; o += NT) {
        const int m = o / (chunks * 16);
        const int rr = o - m * (chunks * 16);
        const int ch = rr >> 4, row = rr & 15;
        const int k = r0 + rr;
        if (k < r1) {
            T vb = from_f<T>(chunk_sum(ch, row, m));
            if (res != nullptr) {
                const float rv = o == static_cast<int>(threadIdx.x) ? res_first : to_f(res[static_cast<size_t>(m) * K + k]);
                vb = from_f<T>(rv + to_f(vb));
            }
            out[static_cast<size_t>(m) * K + k] = vb;
        }
    }
    TL_TRACE_STAMP(15);
}

#if TL_TRACE
void trace_bind_stream6(unsigned long long *buf, unsigned int *n, unsigned int cap) { trace_bind(buf, n, cap); }
#endif

// ---------------------------------------------------------------- host side --
static size_t stream6_smem_bytes(int N, int K, int MP, int grid, int *sb_rows) {
    const int MPA = MP < 8 ? 8 : MP;
    const int all = (K + 15) / 16;
    const int chunks = (all + grid - 1) / grid;
    const int G = N / 128;
    *sb_rows = chunks * 16;
    size_t bytes = static_cast<size_t>(S6_WARPS) * S6_SLOTS * S6_UNIT_BYTES + S6_WARPS * S6_SLOTS * 8;
    bytes += static_cast<size_t>(N / 8) * MP * 16 + static_cast<size_t>(G) * MPA * 4;
    bytes += 2 * static_cast<size_t>(*sb_rows) * G * 2;
    const size_t entries = static_cast<size_t>(chunks + S6_WARPS) * 16 * 8 * 4, sq = static_cast<size_t>(G) * MP * 4;
    bytes += entries > sq ? entries : sq;
    return (bytes + 15) & ~static_cast<size_t>(15);
}
constexpr size_t S6_SMEM_MAX = 112 * 1024;  // two CTAs per SM

bool w4a16_stream6_supported(int M, int N, int K, int dtype) {
    static const bool on = [] { const char *e = getenv("TL_STREAM6"); return e != nullptr && e[0] == '1'; }();
    if (!on || dtype != TL_BF16 || M < 1 || M > 8 || N % 256 != 0 || K < 16) return false;
    int sb_rows;
    const int grid = sm_count();
    return stream6_smem_bytes(N, K, w4_pad_cols(M), grid < (K + 15) / 16 ? grid : (K + 15) / 16, &sb_rows) <= S6_SMEM_MAX;
}

static int s6_map(CUtensorMap *out, const void *ptr, int N, int K) {
    struct Key {
        const void *p;
        int n, k;
        bool operator==(const Key &o) const { return p == o.p && n == o.n && k == o.k; }
    };
    struct Hash {
        size_t operator()(const Key &k) const { return reinterpret_cast<size_t>(k.p) * 1000003u ^ (static_cast<size_t>(k.n) << 20) ^ static_cast<size_t>(k.k); }
    };
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(Key{ptr, N, K});
    if (it != cache.end()) {
        *out = it->second;
        return TL_OK;
    }
    PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
    if (encode == nullptr) return fail(TL_ECUDA, "quantized_matmul: cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(N) / 2, static_cast<cuuint64_t>(K)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(N) / 2};
    const cuuint32_t box[2] = {128, 16};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap map;
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TL_ECUDA, "quantized_matmul: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    if (cache.size() > 8192) cache.clear();
    cache.emplace(Key{ptr, N, K}, map);
    *out = map;
    return TL_OK;
}

template <int MP>
static int stream6_launch(const CUtensorMap &map, Stream6Args args, int grid, size_t smem, cudaStream_t st) {
    using T = __nv_bfloat16;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(w4a16_stream6_kernel<T, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S6_SMEM_MAX)) != cudaSuccess)
            return fail(TL_ECUDA, "quantized_matmul: cannot raise shared memory limit");
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(S6_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, w4a16_stream6_kernel<T, MP>, map, args);
    if (e != cudaSuccess) return fail(TL_ECUDA, "w4a16_stream6: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("w4a16_stream6");
    return TL_OK;
}

int launch_w4a16_stream6(const void *scales, const void *biases, const void *b, void *out, const void *p0, const void *p1,
                         const void *residual, int M, int N, int K, int lda, int prologue, int epilogue, float eps, cudaStream_t st) {
    if (!aligned16(p0) || !aligned16(b) || (p1 && !aligned16(p1)) || (lda % 8) != 0) return fail(TL_EINVAL, "quantized_matmul: operands must be 16-byte aligned");
    const int all = (K + 15) / 16;
    const int grid = all < sm_count() ? all : sm_count();
    const int MP = w4_pad_cols(M);
    Stream6Args args{};
    args.scales = scales, args.biases = biases, args.out = out, args.p0 = p0, args.p1 = p1, args.residual = residual;
    args.M = M, args.N = N, args.K = K, args.lda = lda, args.prologue = prologue, args.epilogue = epilogue, args.eps = eps;
    const size_t smem = stream6_smem_bytes(N, K, MP, grid, &args.sb_rows);
    CUtensorMap map;
    if (int e = s6_map(&map, b, N, K)) return e;
    switch (MP) {
        case 1: return stream6_launch<1>(map, args, grid, smem, st);
        case 2: return stream6_launch<2>(map, args, grid, smem, st);
        case 4: return stream6_launch<4>(map, args, grid, smem, st);
        default: return stream6_launch<8>(map, args, grid, smem, st);
    }
}

}  // namespace tl
