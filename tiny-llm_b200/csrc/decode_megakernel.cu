// Whole-step persistent decode kernel for the Week-3 Qwen3 model (bf16, W4A16,
// paged KV), batch M <= 8, one query token per request.
//
// Why: at B200 speeds the projections of one decode step last 0.2-2 us each at
// HBM rate, but a kernel boundary costs ~5 us (launch + two dependent memory
// round trips before the first byte is used).  Measured (profiles/r01_*): 221
// fused kernels per token => 2.5 ms/token although the HBM traffic is worth
// 0.33 ms.  This kernel runs the ENTIRE step as one cooperative launch:
//
//   embed | per layer: [rmsnorm+qkv] [qk-norm+rope+append+attention] ([merge])
//           [o_proj+residual] [rmsnorm+gate|up] [swiglu+down+residual]
//         | [rmsnorm+head (+argmax partials)] [argmax + advance]
//
// with a grid-wide barrier between phases (every phase consumes a full
// activation vector produced by all CTAs of the previous one).  The point of
// the design is what happens ACROSS the barriers: every warp keeps MK_DEPTH
// weight units (2 KiB each) in flight in registers, and the cursor that issues
// those loads walks the concatenated work list of ALL streaming phases,
// independent of the consumer.  While a warp waits at a barrier or re-stages
// activations, the first units of the next projection are already on their way
// (16 warps x 8 KiB x 148 SMs = 19 MB in flight: a third of a layer).
//
// Work split of a projection [K rows x N]: CTA c owns rows [K*c/C, K*(c+1)/C)
// (all CTAs stream the same number of bytes), in 16-row chunks x N/(128 U) units;
// the (chunk, unit) pairs are dealt to the CTA's warps as contiguous ranges.
// A warp accumulates a chunk in registers (w4a16_item.cuh: mma.sync on exact
// (128+q) codes) and parks its partial sum in shared-memory entry (chunk + warp);
// entries of a chunk are summed in warp order: deterministic.
//
// Arithmetic and rounding points are those of the operator sequence
// (rms_norm -> quantized_matmul -> ... , see engine.py::_forward_unfused).
#include <algorithm>
#include <stdlib.h>
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"
#include "w4a16_item.cuh"
#include "trace.cuh"

namespace tl {

constexpr int MK_WARPS = 16;
constexpr int MK_THREADS = MK_WARPS * 32;
constexpr int MK_MAXB = 8;
constexpr float MK_LOG2E = 1.44269504089f;
constexpr float MK_NEG = -1e30f;

typedef __nv_bfloat16 bf16;

// The argument block is the C-ABI struct of include/tiny_llm_b200.h (plain pointers and sizes).
typedef tl_decode_layer MkLayer;
typedef tl_decode_args MkArgs;

// ------------------------------------------------------------- primitives --
__device__ __forceinline__ uint4 mk_ldcg16(const void *p) { return ld_cg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ float mk_bf(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float mk_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// Optional in-kernel timeline (CTA 0, thread 0): (tag, clock64) pairs.
struct Prof {
    long long *buf;
    int n, cap;
    __device__ __forceinline__ void stamp(int tag) {
        if (buf != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && n < cap) {
            buf[2 * n] = tag;
            buf[2 * n + 1] = clock64();
            n += 1;
        }
    }
};

// Grid-wide barrier: monotonically increasing arrival counter, self-cleaned at kernel exit.
__device__ __forceinline__ void mk_grid_sync(unsigned *counter, unsigned &epoch) {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
        // release-increment / acquire-poll: orders this CTA's phase output before the arrival and
        // the other CTAs' output after the observation (the barrier above made them thread 0's)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        const unsigned target = epoch * gridDim.x;
        unsigned spins = 0;
        while (true) {
            unsigned seen;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen >= target) break;
            if (++spins > (1u << 22)) __trap();  // ~seconds: a lost CTA must not hang the GPU
        }
    }
    __syncthreads();
}

// ------------------------------------------------------- streaming phases --
struct SPhase {
    const uint32_t *w;
    const bf16 *s, *b;
    int N, K;
};

__device__ __forceinline__ int mk_num_sphases(const MkArgs &a) { return 4 * a.n_layers + 1; }

__device__ __forceinline__ SPhase mk_sphase(const MkArgs &a, int sp) {
    SPhase p;
    if (sp >= 4 * a.n_layers) {
        p.w = static_cast<const uint32_t *>(a.w_head), p.s = static_cast<const bf16 *>(a.s_head), p.b = static_cast<const bf16 *>(a.b_head), p.N = a.H, p.K = a.V;
        return p;
    }
    const MkLayer &l = a.layers[sp >> 2];
    switch (sp & 3) {
        case 0: p.w = static_cast<const uint32_t *>(l.w_qkv), p.s = static_cast<const bf16 *>(l.s_qkv), p.b = static_cast<const bf16 *>(l.b_qkv), p.N = a.H, p.K = (a.Hq + 2 * a.Hkv) * a.D; break;
        case 1: p.w = static_cast<const uint32_t *>(l.w_o), p.s = static_cast<const bf16 *>(l.s_o), p.b = static_cast<const bf16 *>(l.b_o), p.N = a.Hq * a.D, p.K = a.H; break;
        case 2: p.w = static_cast<const uint32_t *>(l.w_gu), p.s = static_cast<const bf16 *>(l.s_gu), p.b = static_cast<const bf16 *>(l.b_gu), p.N = a.H, p.K = 2 * a.I; break;
        default: p.w = static_cast<const uint32_t *>(l.w_down), p.s = static_cast<const bf16 *>(l.s_down), p.b = static_cast<const bf16 *>(l.b_down), p.N = a.I, p.K = a.H; break;
    }
    return p;
}

// Rows of this CTA and the unit range of this warp for one streaming phase.  A unit is U
// consecutive 128-column groups of one 16-row chunk (w4a16_item.cuh).
struct SRange {
    int r0, r1, P, chunks, begin, end;  // units [begin, end) in chunk-major order, P units per chunk
};
template <int U>
__device__ __forceinline__ SRange mk_range(const SPhase &p, int warp) {
    SRange r;
    // Whole 16-row chunks per CTA: a chunk shared by two CTAs would cost both of them a full tile of
    // tensor-core work for a few rows each (K = 2560 over 148 CTAs did exactly that).
    const unsigned all = static_cast<unsigned>(p.K + 15) >> 4;  // all * gridDim.x < 2^31: 32-bit divisions only
    const int c0 = static_cast<int>(all * blockIdx.x / gridDim.x), c1 = static_cast<int>(all * (blockIdx.x + 1) / gridDim.x);
    r.r0 = c0 * 16;
    r.r1 = min(p.K, c1 * 16);
    r.P = p.N / (128 * U);
    r.chunks = c1 - c0;
    const unsigned units = r.chunks * r.P;
    r.begin = static_cast<int>(units * warp / MK_WARPS);
    r.end = static_cast<int>(units * (warp + 1) / MK_WARPS);
    return r;
}

// Per-warp register pipeline: DEPTH weight units in flight, filled by a cursor that walks the
// (streaming phase, unit) pairs of the WHOLE step, so the first units of the next projection are
// already on their way while the warp sits in a grid barrier or re-stages activations.
#ifndef MK_DEPTH_N
#define MK_DEPTH_N 4
#endif
constexpr int MK_DEPTH = MK_DEPTH_N;
template <int U>
struct Pipe {
    W4Unit<U> buf[MK_DEPTH];
    int slot;                            // buffer the consumer takes next
    int sp, i, end, hold;                // load cursor: phase, next unit, end of this warp's range; parked at an o_proj phase
    int u, P, row_base, r1, G;           // unit inside its chunk, units per chunk, chunk rows, CTA row end
    const unsigned char *w, *tab;        // phase weights; this lane's scale (lanes 0-15) or bias table
    const unsigned char *p0, *p8, *sb;   // this lane's addresses for the next unit
};

template <int U>
__device__ __forceinline__ void mk_chunk_ptrs(Pipe<U> &c, int lane) {  // rows past the CTA's range are clamped
    const int g = lane >> 2, t = lane & 3;
    const int last = c.r1 - 1;
    const size_t row_bytes = static_cast<size_t>(c.G) * 64;
    c.p0 = c.w + min(c.row_base + g, last) * row_bytes + c.u * (U * 64) + t * 16;
    c.p8 = c.w + min(c.row_base + g + 8, last) * row_bytes + c.u * (U * 64) + t * 16;
    c.sb = c.tab + (static_cast<size_t>(min(c.row_base + (lane & 15), last)) * c.G + c.u * U) * 2;
}

// Position the load cursor on the first unit of phase c.sp (or a later one) that belongs to this
// warp.  The cursor never walks into an o_proj phase on its own (resume == false): the attention
// phase in front of it wants the registers, so the pipeline is drained there and refilled after.
template <int U>
__device__ __forceinline__ void mk_cursor_phase(const MkArgs &a, Pipe<U> &c, int warp, int lane, bool resume) {
    while (c.sp < mk_num_sphases(a)) {
        if (!resume && (c.sp & 3) == 1 && c.sp < 4 * a.n_layers) {
            c.hold = 1;
            return;
        }
        resume = false;
        const SPhase p = mk_sphase(a, c.sp);
        const SRange r = mk_range<U>(p, warp);
        if (r.begin < r.end) {
            c.i = r.begin, c.end = r.end, c.P = r.P, c.r1 = r.r1, c.G = p.N / 128;
            const int chunk = r.begin / r.P;
            c.u = r.begin - chunk * r.P;
            c.row_base = r.r0 + chunk * 16;
            c.w = reinterpret_cast<const unsigned char *>(p.w);
            c.tab = reinterpret_cast<const unsigned char *>(lane < 16 ? p.s : p.b);
            mk_chunk_ptrs<U>(c, lane);
            return;
        }
        c.sp += 1;
    }
}

// FORCE: always define `un` (a dummy load when the cursor is parked or exhausted), so that the
// pipeline registers are provably dead before a refill.
template <int U, bool FORCE>
__device__ __forceinline__ void mk_load_next(const MkArgs &a, Pipe<U> &c, W4Unit<U> &un, int warp, int lane) {
    if (c.sp < mk_num_sphases(a) && !c.hold) {
        w4_load<U>(un, c.p0, c.p8, c.sb);
        c.i += 1;
        if (c.i == c.end) {
            c.sp += 1;
            mk_cursor_phase<U>(a, c, warp, lane, false);
        } else if (++c.u == c.P) {
            c.u = 0;
            c.row_base += 16;
            mk_chunk_ptrs<U>(c, lane);
        } else {
            c.p0 += U * 64;
            c.p8 += U * 64;
            c.sb += U * 2;
        }
    } else if (FORCE) {
        const unsigned char *dummy = static_cast<const unsigned char *>(a.w_head);
        w4_load<U>(un, dummy, dummy, static_cast<const unsigned char *>(a.s_head));
    }
}

// (Re)start the pipeline: every buffer is rewritten, the consumer restarts at buffer 0.
template <int U>
__device__ __forceinline__ void mk_refill(const MkArgs &a, Pipe<U> &c, int warp, int lane, bool resume) {
    c.slot = 0;
    if (resume && c.hold) {  // only the o_proj phase itself releases a parked cursor
        c.hold = 0;
        mk_cursor_phase<U>(a, c, warp, lane, true);
    }
#pragma unroll
    for (int k = 0; k < MK_DEPTH; ++k) mk_load_next<U, true>(a, c, c.buf[k], warp, lane);
}

// One projection: out[m, k] (= res[m, k] +) sum_n act[m, n] * dequant(w[k, n]) for this CTA's rows.
// prologue 0: plain; 1: rms_norm(x, norm_w, eps); 2: swiglu(gate, up) with up = in + up_off.
template <int MP, int U>
__device__ void mk_stream(const MkArgs &a, int sp, Pipe<U> &pipe, const bf16 *in, int ld, int up_off, int prologue, const bf16 *norm_w,
                          bf16 *out, const bf16 *res, bool swiglu_pairs, unsigned char *dyn, Prof &prof) {
    constexpr int MPA = w4_mpa(MP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const SPhase p = mk_sphase(a, sp);
    const SRange r = mk_range<U>(p, warp);
    const int B = a.B;
    const int words = p.N / 8, G = p.N / 128;
    uint4 *act = reinterpret_cast<uint4 *>(dyn);
    float *asum = reinterpret_cast<float *>(dyn + static_cast<size_t>(words) * MP * 16);
    float *rowstat = asum + G * MPA;
    float *entries = rowstat + 32;  // [(chunks + MK_WARPS) entries][16 rows][8 cols]
    w4_stage<bf16, MP, MK_THREADS>(in, ld, prologue == 2 ? in + up_off : norm_w, prologue, p.N, B, a.eps, act, asum, rowstat, []() {});
    prof.stamp(100 + sp * 10 + 1);

    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    int chunk = r.begin / r.P;
    int u = r.begin - chunk * r.P;
    const uint4 *act0 = w4_act_lane<MP>(act, g, t);
    const float *asum0 = w4_asum_lane<MP>(asum, t);
    const uint4 *actp = act0 + u * U * w4_act_group_stride<MP>();
    const float *asump = asum0 + u * U * w4_asum_group_stride<MP>();
    auto flush = [&]() {  // park this warp's partial sums of `chunk` in entry (chunk + warp)
        float *e = entries + static_cast<size_t>(chunk + warp) * 128;
        e[g * 8 + 2 * t] = acc[0][0];
        e[g * 8 + 2 * t + 1] = acc[0][1];
        e[(g + 8) * 8 + 2 * t] = acc[0][2];
        e[(g + 8) * 8 + 2 * t + 1] = acc[0][3];
        acc[0][0] = acc[0][1] = acc[0][2] = acc[0][3] = 0.f;
    };
    int i = r.begin;
    while (i < r.end) {
#pragma unroll
        for (int k = 0; k < MK_DEPTH; ++k) {
            if (pipe.slot == k && i < r.end) {  // warp-uniform; k is a compile-time register index
                w4_consume<bf16, MP, U>(pipe.buf[k], actp, asump, g, acc);
                mk_load_next<U, false>(a, pipe, pipe.buf[k], warp, lane);
                pipe.slot = (k + 1) % MK_DEPTH;
                i += 1;
                actp += U * w4_act_group_stride<MP>();
                asump += U * w4_asum_group_stride<MP>();
                if (++u == r.P) {
                    flush();
                    u = 0;
                    chunk += 1;
                    actp = act0;
                    asump = asum0;
                }
            }
        }
    }
    if (u != 0) flush();
    prof.stamp(100 + sp * 10 + 2);
    __syncthreads();
    prof.stamp(100 + sp * 10 + 3);

    // ---- deterministic reduction of the entries of each chunk (in warp order) + epilogue
    const unsigned units = r.chunks * r.P;
    if (swiglu_pairs) {  // gate rows 0-7 / up rows 8-15 of every chunk (engine.py::_interleave_gate_up) -> 8 activations
        auto chunk_sum = [&](int ch, int row, int m) {
            float v = 0.f;
            for (int w = 0; w < MK_WARPS; ++w) {
                const int wb = static_cast<int>(units * w / MK_WARPS);
                const int we = static_cast<int>(units * (w + 1) / MK_WARPS);
                if (wb < we && wb < (ch + 1) * r.P && we > ch * r.P) v += entries[static_cast<size_t>(ch + w) * 128 + row * 8 + m];
            }
            return v;
        };
        for (int o = threadIdx.x; o < r.chunks * 8 * B; o += MK_THREADS) {
            const int m = o / (r.chunks * 8);
            const int rr = o - m * (r.chunks * 8);
            const int ch = rr >> 3, row = rr & 7;
            const float gate = mk_round(chunk_sum(ch, row, m)), up = mk_round(chunk_sum(ch, row + 8, m));
            out[static_cast<size_t>(m) * (p.K / 2) + (r.r0 >> 1) + rr] = __float2bfloat16_rn((gate / (1.0f + expf(-gate))) * up);
        }
        __syncthreads();
        return;
    }
    for (int o = threadIdx.x; o < r.chunks * 16 * B; o += MK_THREADS) {
        const int m = o / (r.chunks * 16);  // request-major so that consecutive threads store consecutive features
        const int rr = o - m * (r.chunks * 16);
        const int ch = rr >> 4, row = rr & 15;
        const int k = r.r0 + rr;
        if (k < r.r1) {
            float v = 0.f;
            for (int w = 0; w < MK_WARPS; ++w) {
                const int wb = static_cast<int>(units * w / MK_WARPS);
                const int we = static_cast<int>(units * (w + 1) / MK_WARPS);
                if (wb < we && wb < (ch + 1) * r.P && we > ch * r.P) v += entries[static_cast<size_t>(ch + w) * 128 + row * 8 + m];
            }
            bf16 vb = __float2bfloat16_rn(v);
            if (res != nullptr) vb = __float2bfloat16_rn(mk_bf(ld_cg(res + static_cast<size_t>(m) * p.K + k)) + mk_bf(vb));  // L2: written by other phases
            out[static_cast<size_t>(m) * p.K + k] = vb;
        }
    }
    __syncthreads();
}

struct BestPair {
    float v;
    int i;
};
__device__ __forceinline__ BestPair mk_better(BestPair a, BestPair b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

// Arg-max of the logits rows [r0, r1) this CTA has just written, one result per request.
__device__ void mk_cta_argmax(const MkArgs &a, int r0, int r1, float *scratch_v, int *scratch_i) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int m = 0; m < a.B; ++m) {
        BestPair mine{-CUDART_INF_F, 0x7fffffff};
        const bf16 *row = static_cast<const bf16 *>(a.logits) + static_cast<size_t>(m) * a.V;
        for (int k = r0 + threadIdx.x; k < r1; k += MK_THREADS) mine = mk_better(mine, BestPair{mk_bf(row[k]), k});
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            BestPair other{__shfl_xor_sync(0xffffffffu, mine.v, o), __shfl_xor_sync(0xffffffffu, mine.i, o)};
            mine = mk_better(mine, other);
        }
        if (lane == 0) scratch_v[warp] = mine.v, scratch_i[warp] = mine.i;
        __syncthreads();
        if (warp == 0) {
            BestPair r{lane < MK_WARPS ? scratch_v[lane] : -CUDART_INF_F, lane < MK_WARPS ? scratch_i[lane] : 0x7fffffff};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                BestPair other{__shfl_xor_sync(0xffffffffu, r.v, o), __shfl_xor_sync(0xffffffffu, r.i, o)};
                r = mk_better(r, other);
            }
            if (lane == 0) a.amax_val[blockIdx.x * a.B + m] = r.v, a.amax_idx[blockIdx.x * a.B + m] = r.i;
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------- attention --
// One CTA per (request, kv head, split), L == 1.  The phase is pure latency (8 CTAs x 64 KB of
// K/V at context 128), so it is organised around dependent round trips and wide parallelism, not
// bandwidth.  Per round of up to MK_ATT_TOK tokens:
//   A  page ids of the round -> shared; warps 0..G+1 already hold their q/k/v head rows, norm
//      weights and rope frequencies in registers (loaded before anything else)
//   B  cp.async ALL K/V rows of the round into shared memory (row stride 528 B: bank spread)
//   C  while those are in flight: q/k RMSNorm + RoPE (rounded like rms_norm -> rope), V copy
//   S  scores = Q K^T on the tensor cores: mma.sync m16n8k16, A = the G query heads (rows >= G
//      zero), B = K rows straight from shared memory; 8 tokens per MMA tile, 16 warps
//   M  one warp per head: max / exp2 / sum over the round's tokens (fp32)
//   V  out = P V on CUDA cores with fp32 probabilities: thread = (head, 8 dims, token subset)
//   then the running (max, sum, out) of thread (head, dim) absorbs the round.
constexpr int MK_ATT_TOK = 256;     // K/V rows staged per round
constexpr int MK_KV_STRIDE = 528;   // bytes per staged token: K row | V row | 16 B pad
constexpr size_t MK_ATT_BYTES = 4 * 128 * 2 + 2 * 128 * 2 + 4 * MK_ATT_TOK * 4 + 256 * 4 + (MK_ATT_TOK + 4) * 4 +
                                static_cast<size_t>(MK_ATT_TOK) * MK_KV_STRIDE;

__device__ __forceinline__ void mk_cp16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}

// DEPWAIT (stand-alone kernel under programmatic dependent launch): everything that does not read
// this step's qkv row - page ids, the K/V rows of older tokens - is requested BEFORE
// griddepcontrol.wait, i.e. while the qkv projection is still draining.
template <bool DEPWAIT>
__device__ void mk_attention(const MkArgs &a, const MkLayer &l, unsigned char *dyn, Prof &prof) {
    const int D = a.D, G = a.Hq / a.Hkv;
    const int items = a.B * a.Hkv * a.nsplit;
    if (static_cast<int>(blockIdx.x) >= items) return;
    const int split = blockIdx.x % a.nsplit;
    const int kvh = (blockIdx.x / a.nsplit) % a.Hkv;
    const int b = blockIdx.x / (a.nsplit * a.Hkv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qkv_w = (a.Hq + 2 * a.Hkv) * D;

    bf16 *q_s = reinterpret_cast<bf16 *>(dyn);                      // [4][128] rope output (unscaled)
    bf16 *k_cur = q_s + 4 * 128;                                    // [128] newest token
    bf16 *v_cur = k_cur + 128;                                      // [128]
    float *s_s = reinterpret_cast<float *>(v_cur + 128);            // [4][MK_ATT_TOK] scores, then probabilities
    float *st_s = s_s + 4 * MK_ATT_TOK;                             // per (head, tile): max [4][32], then sum [4][32]
    int *pg_s = reinterpret_cast<int *>(st_s + 256);                // page ids of the round
    unsigned char *kv_s = reinterpret_cast<unsigned char *>(pg_s + MK_ATT_TOK + 4);  // [MK_ATT_TOK][528]

    // ---- head rows for the q path (registers; used after the K/V copies are in flight)
    const bool qpath = warp < G + 2;
    const bool is_q = warp < G, is_k = warp == G;
    float re[2] = {0.f, 0.f}, im[2] = {0.f, 0.f}, wre[2] = {0.f, 0.f}, wim[2] = {0.f, 0.f};
    double freq[2] = {0.0, 0.0};
    int position = 0;
    auto load_q = [&]() {
        if (qpath) {
            const int head_off = is_q ? (kvh * G + warp) * D : (is_k ? (a.Hq + kvh) * D : (a.Hq + a.Hkv + kvh) * D);
            const bf16 *src = static_cast<const bf16 *>(a.qkv) + static_cast<size_t>(b) * qkv_w + head_off;
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // lane owns pairs (i, i+64) for i = lane, lane+32   (D == 128)
                re[h] = mk_bf(ld_cg(src + lane + 32 * h));
                im[h] = mk_bf(ld_cg(src + lane + 32 * h + 64));
            }
        }
    };
    if (qpath) {
        const bf16 *w = static_cast<const bf16 *>(is_q ? l.q_norm : l.k_norm);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (is_q || is_k) {
                wre[h] = mk_bf(w[lane + 32 * h]);
                wim[h] = mk_bf(w[lane + 32 * h + 64]);
                freq[h] = a.rope_inv_freq[lane + 32 * h];
            }
        }
        position = a.offsets[b];
    }
    if (!DEPWAIT) load_q();
    // page ids of the first round do not need the context length: request them together with it
    const int begin = split * a.tokens_per_split;
    {
        const int lp_first = begin / a.page_size + static_cast<int>(threadIdx.x);
        if (static_cast<int>(threadIdx.x) <= MK_ATT_TOK && lp_first < a.max_pages)
            pg_s[threadIdx.x] = l.table[static_cast<size_t>(b) * a.max_pages + lp_first];
    }
    const int ctx = min(a.context_lens[b], a.max_pages * a.page_size);
    const int end = min(ctx, begin + a.tokens_per_split);
    const int cur_tok = ctx - 1;
    const int gidx = warp * 2 + (lane >> 4), c8 = lane & 15;   // copy mapping: lane group of 16 per token row
    const int g = lane >> 2, t = lane & 3;                      // MMA fragment coordinates
    const int oh = threadIdx.x >> 7, od = threadIdx.x & 127;    // owner of out[head oh][dim od]
    const float scale2 = a.attn_scale * MK_LOG2E;
    float m_run = MK_NEG, l_run = 0.f, o_run[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // kept by the (threadIdx & 7) == 0 lanes
    prof.stamp(50001);

    for (int rb = begin; rb < end || rb == begin; rb += MK_ATT_TOK) {  // one pass even for an empty split (q path, barriers)
        const int rend = min(end, rb + MK_ATT_TOK);
        const int cnt = max(rend - rb, 0);
        const int lp0 = rb / a.page_size;
        const int npg = cnt > 0 ? (rend - 1) / a.page_size - lp0 + 1 : 0;
        if (rb != begin && static_cast<int>(threadIdx.x) < npg) pg_s[threadIdx.x] = l.table[static_cast<size_t>(b) * a.max_pages + lp0 + threadIdx.x];
        __syncthreads();
        prof.stamp(50002);
        // ---- B: all K/V rows of the round in flight
        {
            int lpi = (rb + gidx) / a.page_size;        // logical page and row inside it of this lane group's next token,
            int row = rb + gidx - lpi * a.page_size;    // advanced by 32 tokens per step without further divisions
            const size_t head_rows = static_cast<size_t>(a.Hkv) * a.page_size;
            unsigned char *dst = kv_s + gidx * MK_KV_STRIDE + c8 * 16;
#pragma unroll
            for (int j = 0; j < MK_ATT_TOK / 32; ++j) {
                const int tok = rb + j * 32 + gidx;
                if (tok < rend && tok != cur_tok) {
                    const int pid = pg_s[lpi - lp0];
                    if (pid >= 0 && pid < a.num_pages) {
                        const size_t off = ((pid * head_rows + static_cast<size_t>(kvh) * a.page_size + row) << 7) + c8 * 8;  // D == 128
                        mk_cp16(dst, static_cast<const bf16 *>(l.k_pages) + off);
                        mk_cp16(dst + 256, static_cast<const bf16 *>(l.v_pages) + off);
                    }
                }
                dst += 32 * MK_KV_STRIDE;
                row += 32;
                while (row >= a.page_size) row -= a.page_size, lpi += 1;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        prof.stamp(50003);
        if (DEPWAIT && rb == begin) {
            TL_TRACE_STAMP(21);
            asm volatile("griddepcontrol.wait;" ::: "memory");  // the qkv row of this step exists now
            TL_TRACE_STAMP(22);
            load_q();
        }
        // ---- C: q path (first round only)
        if (rb == begin && qpath) {
            if (is_q || is_k) {
                float ss = re[0] * re[0] + im[0] * im[0] + re[1] * re[1] + im[1] * im[1];
                ss = warp_sum(ss);
                const float inv = rsqrtf(ss / static_cast<float>(D) + a.eps);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
                    const float nre = mk_round(re[h] * inv * wre[h]);
                    const float nim = mk_round(im[h] * inv * wim[h]);
                    const float angle = static_cast<float>(static_cast<double>(position) * freq[h]);
                    float sn, cs;
                    sincosf(angle, &sn, &cs);
                    const bf16 ore = __float2bfloat16_rn(nre * cs - nim * sn), oim = __float2bfloat16_rn(nim * cs + nre * sn);
                    bf16 *dst = is_q ? q_s + warp * 128 : k_cur;
                    dst[i] = ore, dst[i + 64] = oim;
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) v_cur[lane + 32 * h] = __float2bfloat16_rn(re[h]), v_cur[lane + 32 * h + 64] = __float2bfloat16_rn(im[h]);
            }
        }
        prof.stamp(50004);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // q_s / k_cur / v_cur and every lane's K/V pieces are visible
        // the newest token: its K/V rows join the staged rows, and the split that owns it appends it
        // to the cache (paged_cache_update semantics)
        if (ctx > 0 && cur_tok >= rb && cur_tok < rend && threadIdx.x < 2 * (D / 8)) {
            const bool kk = threadIdx.x < D / 8;
            const int ch = kk ? threadIdx.x : threadIdx.x - D / 8;
            const uint4 row = *reinterpret_cast<const uint4 *>((kk ? k_cur : v_cur) + ch * 8);
            *reinterpret_cast<uint4 *>(kv_s + (cur_tok - rb) * MK_KV_STRIDE + (kk ? 0 : 256) + ch * 16) = row;
            const int lp = cur_tok / a.page_size;
            const int pid = pg_s[lp - lp0];
            if (pid >= 0 && pid < a.num_pages) {
                bf16 *dst = static_cast<bf16 *>(kk ? l.k_pages : l.v_pages) + ((static_cast<size_t>(pid) * a.Hkv + kvh) * a.page_size + (cur_tok - lp * a.page_size)) * D + ch * 8;
                *reinterpret_cast<uint4 *>(dst) = row;
            }
        }
        __syncthreads();
        prof.stamp(50005);
        if (DEPWAIT) TL_TRACE_STAMP(23);
        // ---- S: scores of 8 tokens x G heads per MMA tile, with the tile's softmax statistics
        for (int tile = warp; tile * 8 < cnt; tile += MK_WARPS) {
            float d[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};  // two independent MMA chains
            const unsigned char *krow = kv_s + (tile * 8 + g) * MK_KV_STRIDE + t * 4;
            const bf16 *qrow = q_s + (g < G ? g : 0) * 128 + 2 * t;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t a0 = *reinterpret_cast<const uint32_t *>(qrow + ks * 16);
                uint32_t a2 = *reinterpret_cast<const uint32_t *>(qrow + ks * 16 + 8);
                if (g >= G) a0 = a2 = 0u;
                const uint32_t b0 = *reinterpret_cast<const uint32_t *>(krow + ks * 32);
                const uint32_t b1 = *reinterpret_cast<const uint32_t *>(krow + ks * 32 + 16);
                if (ks & 1)
                    W4Num<bf16>::mma(d2, a0, 0u, a2, 0u, b0, b1);
                else
                    W4Num<bf16>::mma(d, a0, 0u, a2, 0u, b0, b1);
            }
            // lane (g, t) holds head g, tokens 2t and 2t+1 of the tile
            float sc[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int slot = tile * 8 + 2 * t + e;
                bool ok = slot < cnt;
                if (ok) {
                    const int pid = pg_s[(rb + slot) / a.page_size - lp0];
                    ok = pid >= 0 && pid < a.num_pages;
                }
                sc[e] = ok ? (d[e] + d2[e]) * scale2 : -CUDART_INF_F;
            }
            float mx = fmaxf(sc[0], sc[1]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float ref = mx == -CUDART_INF_F ? 0.f : mx;  // a fully masked tile: probabilities 0, not NaN
            const float p0 = exp2f(sc[0] - ref), p1 = exp2f(sc[1] - ref);
            float ls = p0 + p1;
            ls += __shfl_xor_sync(0xffffffffu, ls, 1);
            ls += __shfl_xor_sync(0xffffffffu, ls, 2);
            if (g < G) {
                *reinterpret_cast<float2 *>(s_s + g * MK_ATT_TOK + tile * 8 + 2 * t) = make_float2(p0, p1);
                if (t == 0) st_s[g * 32 + tile] = mx == -CUDART_INF_F ? MK_NEG : mx, st_s[128 + g * 32 + tile] = ls;
            }
        }
        __syncthreads();
        prof.stamp(50008);
        // ---- V: thread = (token of the tile, 8 dims, head); one staged row per tile, then a shuffle
        // reduction over the 8 tokens of a tile position; the sub == 0 lane keeps the running state
        {
            const int sub = threadIdx.x & 7, d8 = (threadIdx.x >> 3) & 15;
            const int ntile = (cnt + 7) >> 3;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float m_r = MK_NEG, l_r = 0.f;
            if (oh < G) {
                for (int tile = 0; tile < ntile; ++tile) m_r = fmaxf(m_r, st_s[oh * 32 + tile]);
                for (int tile = 0; tile < ntile; ++tile) {
                    const float f = exp2f(st_s[oh * 32 + tile] - m_r);
                    l_r += st_s[128 + oh * 32 + tile] * f;
                    const int slot = tile * 8 + sub;
                    const float pr = s_s[oh * MK_ATT_TOK + slot] * f;
                    if (pr != 0.f) {  // masked / padded slots hold no valid V row
                        const uint4 vr = *reinterpret_cast<const uint4 *>(kv_s + slot * MK_KV_STRIDE + 256 + d8 * 16);
                        const float2 f0 = unpack2<bf16>(vr.x), f1 = unpack2<bf16>(vr.y), f2 = unpack2<bf16>(vr.z), f3 = unpack2<bf16>(vr.w);
                        acc[0] += pr * f0.x, acc[1] += pr * f0.y, acc[2] += pr * f1.x, acc[3] += pr * f1.y;
                        acc[4] += pr * f2.x, acc[5] += pr * f2.y, acc[6] += pr * f3.x, acc[7] += pr * f3.y;
                    }
                }
            }
#pragma unroll
            for (int off = 1; off < 8; off <<= 1)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
            if (oh < G && sub == 0) {
                const float nm = fmaxf(m_run, m_r);
                const float fr = exp2f(m_run - nm), fn = exp2f(m_r - nm);
#pragma unroll
                for (int i = 0; i < 8; ++i) o_run[i] = o_run[i] * fr + acc[i] * fn;
                l_run = l_run * fr + l_r * fn;
                m_run = nm;
            }
        }
        prof.stamp(50006);
        __syncthreads();  // the round's page ids, rows and scores are dead: the next round may overwrite them
    }
    prof.stamp(50007);
    if (oh < G && (threadIdx.x & 7) == 0) {  // this lane owns out[head oh][8 dims]
        const int head = kvh * G + oh, d0 = ((threadIdx.x >> 3) & 15) * 8;
        if (a.nsplit == 1) {
            const float inv = l_run == 0.f ? 0.f : 1.0f / l_run;
            uint4 o;
            o.x = pack2<bf16>(o_run[0] * inv, o_run[1] * inv), o.y = pack2<bf16>(o_run[2] * inv, o_run[3] * inv);
            o.z = pack2<bf16>(o_run[4] * inv, o_run[5] * inv), o.w = pack2<bf16>(o_run[6] * inv, o_run[7] * inv);
            *reinterpret_cast<uint4 *>(static_cast<bf16 *>(a.y) + (static_cast<size_t>(b) * a.Hq + head) * D + d0) = o;
        } else {
            const size_t row = (static_cast<size_t>(b) * a.Hq + head) * a.nsplit + split;
#pragma unroll
            for (int i = 0; i < 8; ++i) a.attn_ws[row * (D + 2) + d0 + i] = o_run[i];
            if (d0 == 0) a.attn_ws[row * (D + 2) + D] = m_run, a.attn_ws[row * (D + 2) + D + 1] = l_run;
        }
    }
}

__device__ void mk_attention_merge(const MkArgs &a) {
    const int D = a.D;
    const int heads = a.B * a.Hq;
    const int warp_global = blockIdx.x * MK_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (int h = warp_global; h < heads; h += gridDim.x * MK_WARPS) {
        const float *base = a.attn_ws + static_cast<size_t>(h) * a.nsplit * (D + 2);
        float gm = MK_NEG;
        for (int s = 0; s < a.nsplit; ++s) gm = fmaxf(gm, ld_cg(base + s * (D + 2) + D));
        float gl = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < a.nsplit; ++s) {
            const float f = exp2f(ld_cg(base + s * (D + 2) + D) - gm);
            gl += ld_cg(base + s * (D + 2) + D + 1) * f;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] += ld_cg(base + s * (D + 2) + lane + 32 * i) * f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) static_cast<bf16 *>(a.y)[static_cast<size_t>(h) * D + lane + 32 * i] = __float2bfloat16_rn(gl == 0.f ? 0.f : o[i] / gl);
    }
}

// ------------------------------------------------------------------ kernel --
template <int MP, int U>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_megakernel(const MkArgs a) {
    extern __shared__ __align__(128) unsigned char mk_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *dyn = mk_smem_raw;
    __shared__ float amax_v[MK_WARPS];
    __shared__ int amax_i[MK_WARPS];
    unsigned epoch = 0;
    Prof prof{a.prof, 0, a.prof_capacity};
    prof.stamp(1);

    Pipe<U> pipe;
    pipe.sp = 0, pipe.hold = 0;
    mk_cursor_phase<U>(a, pipe, warp, lane, false);
    mk_refill<U>(a, pipe, warp, lane, false);  // weights first: they depend on nothing

    // ---- phase 0: embedding rows -> xa (quantized_matmul.metal:58-89); feature words dealt across the grid
    {
        const int words = a.H / 8;
        for (int idx = blockIdx.x * MK_THREADS + threadIdx.x; idx < a.B * words; idx += gridDim.x * MK_THREADS) {
            const int m = idx / words, wc = idx - m * words;
            const int row = a.tokens[m];
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (row >= 0 && row < a.V) {
                const uint32_t packed = static_cast<const uint32_t *>(a.w_emb)[static_cast<size_t>(row) * words + wc];
                const size_t gidx = static_cast<size_t>(row) * (a.H / 128) + wc / 16;
                const float s = mk_bf(static_cast<const bf16 *>(a.s_emb)[gidx]), bb = mk_bf(static_cast<const bf16 *>(a.b_emb)[gidx]);
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = static_cast<float>((packed >> (4 * j)) & 0xFu) * s + bb;
                o.x = pack2<bf16>(v[0], v[1]), o.y = pack2<bf16>(v[2], v[3]), o.z = pack2<bf16>(v[4], v[5]), o.w = pack2<bf16>(v[6], v[7]);
            }
            *reinterpret_cast<uint4 *>(static_cast<bf16 *>(a.xa) + static_cast<size_t>(m) * a.H + wc * 8) = o;
        }
    }
    prof.stamp(2);
    mk_grid_sync(a.sync_counter, epoch);
    prof.stamp(3);

    bf16 *x = static_cast<bf16 *>(a.xa), *x_alt = static_cast<bf16 *>(a.xb);
    bf16 *qkv = static_cast<bf16 *>(a.qkv), *yb = static_cast<bf16 *>(a.y), *gu = static_cast<bf16 *>(a.gu);
    // One mk_stream call site for all 4L+1 projections (the streaming loop must stay in the
    // instruction cache; five inlined copies did not).
    const int head = 4 * a.n_layers;
    for (int sp = 0; sp <= head; ++sp) {
        const int kind = sp < head ? (sp & 3) : 4;
        const MkLayer &l = a.layers[sp < head ? (sp >> 2) : a.n_layers - 1];
        const bf16 *in = x, *norm_w = nullptr, *res = nullptr;
        bf16 *out = qkv;
        int ld = a.H, up_off = 0, prologue = 1;
        switch (kind) {
            case 0: norm_w = static_cast<const bf16 *>(l.ln1); break;
            case 1: in = yb, ld = a.Hq * a.D, prologue = 0, out = x_alt, res = x; break;
            case 2: in = x_alt, norm_w = static_cast<const bf16 *>(l.ln2), out = gu; break;  // emits swiglu(gate, up) [B, I]
            case 3: in = gu, ld = a.I, prologue = 0, out = x, res = x_alt; break;
            default: norm_w = static_cast<const bf16 *>(a.final_norm), out = static_cast<bf16 *>(a.logits); break;
        }
        if (kind == 1) {  // attention sits in front of o_proj; the weight pipeline is empty here (see mk_cursor_phase)
            mk_attention<false>(a, l, dyn, prof);
            prof.stamp(100 + (sp - 1) * 10 + 6);
            // restart the weight stream BEFORE the barrier: the 140 CTAs without an attention item get
            // here at once, and everyone's first o_proj units fly while the grid synchronises
            mk_refill<U>(a, pipe, warp, lane, true);
            mk_grid_sync(a.sync_counter, epoch);
            prof.stamp(100 + (sp - 1) * 10 + 7);
            if (a.nsplit > 1) {
                mk_attention_merge(a);
                mk_grid_sync(a.sync_counter, epoch);
            }
        }
        mk_stream<MP, U>(a, sp, pipe, in, ld, up_off, prologue, norm_w, out, res, kind == 2, dyn, prof);
        if (kind == 4) {  // greedy arg-max partials of this CTA's logits rows
            const SRange r = mk_range<U>(mk_sphase(a, sp), warp);
            mk_cta_argmax(a, r.r0, r.r1, amax_v, amax_i);
        }
        prof.stamp(100 + sp * 10 + 4);
        mk_grid_sync(a.sync_counter, epoch);
        prof.stamp(100 + sp * 10 + 5);
    }
    if (blockIdx.x == 0 && warp == 0) {
        const int step = *a.step_counter;
        for (int m = 0; m < a.B; ++m) {
            BestPair r{-CUDART_INF_F, 0x7fffffff};
            for (int c = lane; c < static_cast<int>(gridDim.x); c += 32) r = mk_better(r, BestPair{ld_cg(a.amax_val + c * a.B + m), ld_cg(a.amax_idx + c * a.B + m)});
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                BestPair other{__shfl_xor_sync(0xffffffffu, r.v, o), __shfl_xor_sync(0xffffffffu, r.i, o)};
                r = mk_better(r, other);
            }
            if (lane == 0) {
                const int tok = r.i == 0x7fffffff ? 0 : r.i;
                a.next_tokens[m] = tok;
                if (a.advance) {
                    const bool active = a.context_lens[m] > 0;
                    if (active) a.tokens[m] = tok, a.offsets[m] += 1, a.context_lens[m] += 1;
                    if (step < a.log_capacity) a.out_log[static_cast<size_t>(step) * a.B + m] = active ? tok : -1;
                }
            }
        }
        if (lane == 0 && a.advance) *a.step_counter = step + 1;
    }
    prof.stamp(99999);
    // ---- leave the barrier counters clean for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(a.exit_counter, 1u);
        if (ticket == gridDim.x - 1) {
            *a.sync_counter = 0u;
            *a.exit_counter = 0u;
            __threadfence();
        }
    }
}

// -------------------------------------------------------------- host side --
size_t mk_dyn_bytes(const MkArgs &a, int grid, int MP) {
    // act staging + group sums + row stats + partial-sum entries, maximised over the five projection shapes
    const int MPA = MP < 8 ? 8 : MP;
    auto need = [&](int N, int K) {
        const size_t rows = static_cast<size_t>((static_cast<long long>(K) + grid - 1) / grid) + 1;
        const size_t chunks = (rows + 15) / 16 + 1;
        return static_cast<size_t>(N / 8) * MP * 16 + static_cast<size_t>(N / 128) * MPA * 4 + 128 + (chunks + MK_WARPS) * 512;
    };
    size_t best = need(a.H, (a.Hq + 2 * a.Hkv) * a.D);
    best = std::max(best, need(a.Hq * a.D, a.H));
    best = std::max(best, need(a.H, 2 * a.I));
    best = std::max(best, need(a.I, a.H));
    best = std::max(best, need(a.H, a.V));
    return std::max(best, MK_ATT_BYTES + 64);
}

template <int MP, int U>
static int mk_launch_t(const MkArgs &a, cudaStream_t st) {
    const int grid = sm_count();  // one CTA per SM
    const size_t smem = mk_dyn_bytes(a, grid, MP);
    constexpr size_t kMaxDyn = 226 * 1024;  // 227 KiB opt-in limit minus this kernel's static shared memory
    if (smem > kMaxDyn) return fail(TL_EINVAL, "decode_step: activations do not fit in shared memory for batch %d", a.B);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(decode_megakernel<MP, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxDyn)) != cudaSuccess)
            return fail(TL_ECUDA, "decode_step: cannot raise shared memory limit");
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_megakernel<MP, U>, MK_THREADS, smem);
        if (per_sm < 1) return fail(TL_ECUDA, "decode_step: kernel does not fit on an SM (smem %zu)", smem);
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(MK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, decode_megakernel<MP, U>, a);
    if (e != cudaSuccess) return fail(TL_ECUDA, "decode_step: launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return check_launch("decode_megakernel");
}

template <int U>
static int mk_launch_u(const MkArgs &a, cudaStream_t st) {
    switch (w4_pad_cols(a.B)) {
        case 1: return mk_launch_t<1, U>(a, st);
        case 2: return mk_launch_t<2, U>(a, st);
        case 4: return mk_launch_t<4, U>(a, st);
        default: return mk_launch_t<8, U>(a, st);
    }
}

int launch_decode_megakernel(const MkArgs &a, cudaStream_t st) {
    if (a.B < 1 || a.B > MK_MAXB) return fail(TL_EINVAL, "decode_step: batch must be 1..8");
    if (a.D != 128 || a.Hq % a.Hkv != 0 || a.Hq / a.Hkv > 4) return fail(TL_EINVAL, "decode_step: needs head_dim 128 and at most 4 query heads per KV head");
    if (a.H % 128 != 0 || a.I % 128 != 0 || (a.Hq * a.D) % 128 != 0) return fail(TL_EINVAL, "decode_step: widths must be multiples of 128");
    if (a.H % 16 != 0 || ((a.Hq + 2 * a.Hkv) * a.D) % 16 != 0) return fail(TL_EINVAL, "decode_step: projection heights must be multiples of 16");
    // pairs of groups per unit need 4-byte aligned scale pairs: every reduction width a multiple of 256
    const bool pairs = a.H % 256 == 0 && a.I % 256 == 0 && (a.Hq * a.D) % 256 == 0;
    return pairs ? mk_launch_u<2>(a, st) : mk_launch_u<1>(a, st);
}

// ---- the attention phase as a kernel of its own (CUDA-graph decode path) ----
__global__ void __launch_bounds__(MK_THREADS, 1) decode_attention_fused_kernel(const MkArgs a, const MkLayer l) {
    extern __shared__ __align__(128) unsigned char att_smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the o_proj stream may prefetch its weights now
    TL_TRACE_STAMP(20);
    Prof prof{nullptr, 0, 0};
    mk_attention<true>(a, l, att_smem_raw, prof);
    TL_TRACE_STAMP(29);
}
#if TL_TRACE
void trace_bind_attention(unsigned long long *buf, unsigned int *n, unsigned int cap) { trace_bind(buf, n, cap); }
#endif
__global__ void __launch_bounds__(MK_THREADS, 1) decode_attention_merge_kernel(const MkArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    mk_attention_merge(a);
}

static int attention_max_split(int batch, int num_kv_heads) {
    const int s = sm_count() / (batch * num_kv_heads);
    return s < 1 ? 1 : s;
}
size_t decode_attention_fused_workspace(int batch, int num_heads, int num_kv_heads) {
    return static_cast<size_t>(batch) * num_heads * attention_max_split(batch, num_kv_heads) * (128 + 2);
}

int launch_decode_attention_fused(const void *qkv, const void *q_norm_weight, const void *k_norm_weight, const int32_t *offsets,
                                  const int32_t *block_table, const int32_t *context_lens, const double *rope_inv_freq,
                                  void *key_pages, void *value_pages, void *out, float *workspace, int batch, int num_heads,
                                  int num_kv_heads, int head_dim, float eps, float scale, int num_pages, int page_size,
                                  int max_pages, int max_context, int dtype, cudaStream_t st) {
    if (batch == 0) return TL_OK;
    if (dtype != TL_BF16 || head_dim != 128 || num_kv_heads < 1 || num_heads % num_kv_heads != 0 || num_heads / num_kv_heads > 4)
        return fail(TL_EINVAL, "decode_attention_fused: needs bfloat16, head_dim 128 and at most 4 query heads per KV head");
    MkArgs a{};
    MkLayer l{};
    a.B = batch, a.Hq = num_heads, a.Hkv = num_kv_heads, a.D = head_dim;
    a.eps = eps, a.attn_scale = scale;
    a.page_size = page_size, a.max_pages = max_pages, a.num_pages = num_pages;
    a.offsets = const_cast<int32_t *>(offsets), a.context_lens = const_cast<int32_t *>(context_lens);
    a.rope_inv_freq = rope_inv_freq;
    a.qkv = const_cast<void *>(qkv), a.y = out, a.attn_ws = workspace;
    const int max_split = attention_max_split(batch, num_kv_heads);
    int tps = (max_context < 1 ? 1 : max_context + max_split - 1) / max_split;
    tps = tps < 2 * MK_ATT_TOK ? 2 * MK_ATT_TOK : tps;  // a split (extra merge launch) only pays beyond two rounds
    tps = (tps + 63) / 64 * 64;
    a.tokens_per_split = tps;
    a.nsplit = (max_context + tps - 1) / tps;
    a.nsplit = a.nsplit < 1 ? 1 : (a.nsplit > max_split ? max_split : a.nsplit);
    l.q_norm = q_norm_weight, l.k_norm = k_norm_weight, l.table = block_table;
    l.k_pages = key_pages, l.v_pages = value_pages;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(decode_attention_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(MK_ATT_BYTES + 64)) != cudaSuccess)
            return fail(TL_ECUDA, "decode_attention_fused: cannot raise shared memory limit");
        configured = true;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(batch * num_kv_heads * a.nsplit);
    cfg.blockDim = dim3(MK_THREADS);
    cfg.dynamicSmemBytes = MK_ATT_BYTES + 64;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 1 : 0;  // without the attribute griddepcontrol.wait returns at once
    cudaError_t e = cudaLaunchKernelEx(&cfg, decode_attention_fused_kernel, a, l);
    if (e != cudaSuccess) return fail(TL_ECUDA, "decode_attention_fused: launch failed: %s", cudaGetErrorString(e));
    TL_LAUNCH_CHECK("decode_attention_fused");
    if (a.nsplit > 1) {
        const int heads = batch * num_heads;
        cfg.gridDim = dim3((heads + MK_WARPS - 1) / MK_WARPS);
        cfg.dynamicSmemBytes = 0;
        e = cudaLaunchKernelEx(&cfg, decode_attention_merge_kernel, a);
        if (e != cudaSuccess) return fail(TL_ECUDA, "decode_attention_merge: launch failed: %s", cudaGetErrorString(e));
        TL_LAUNCH_CHECK("decode_attention_merge");
    }
    return TL_OK;
}

int mk_grid_size() { return sm_count(); }

}  // namespace tl
