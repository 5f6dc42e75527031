// extern "C" entry points of libtiny_llm_b200.so (see include/tiny_llm_b200.h).
// Each launcher checks what can be checked from pointers and sizes, picks the
// kernel, enqueues it on the caller's stream and returns; it never allocates,
// synchronises or throws.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"
#include "kernels.h"
#include "trace.cuh"

namespace tl {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) return TL_OK;
    cudaGetLastError();  // clear the sticky launch error so later calls can report their own
    return fail(TL_ECUDA, "%s: kernel launch failed: %s", what, cudaGetErrorString(e));
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

static bool float_dtype(int dtype) { return dtype == TL_F32 || dtype == TL_F16 || dtype == TL_BF16; }

}  // namespace tl

using namespace tl;

extern "C" {

int tl_abi_version(void) { return 1; }

const char *tl_last_error(void) { return g_error; }

long long tl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int tl_device_info(int *sms, int *major, int *minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(TL_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(TL_ENODEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (sms) *sms = prop.multiProcessorCount;
    if (major) *major = prop.major;
    if (minor) *minor = prop.minor;
    return TL_OK;
}

// ------------------------------------------------------------ W4A16 ------
// Dispatch by activation rows (use_simdgroup): M <= 8 weight-streaming matvec (the reference's matvec limit,
// quantize.py:162-163); 9..128 swap-AB tcgen05 GEMM with split reduction (w4a16_skinny.cu); above that the
// 128 x 128-tile tcgen05 GEMM.  The streaming kernel also takes what the tensor-core kernels cannot.
static bool use_skinny_kernel(int M, int N, int K, int dtype, int use_simdgroup) {
    return use_simdgroup && M > TL_MATVEC_REF_ROWS && w4a16_skinny_supported(M, N, K, dtype);
}
static bool use_stream_kernel(int M, int N, int K, int dtype, int use_simdgroup) {
    return use_simdgroup && !use_skinny_kernel(M, N, K, dtype, use_simdgroup) &&
           (M <= TL_MATVEC_MAX_ROWS || !w4a16_gemm_supported(M, N, K, dtype));
}

size_t tl_quantized_matmul_workspace(int M, int N, int K, int dtype, int use_simdgroup, int use_split_k) {
    if (!use_simdgroup) return 0;
    if (use_skinny_kernel(M, N, K, dtype, use_simdgroup)) return w4a16_skinny_workspace(M, N, K);
    if (use_stream_kernel(M, N, K, dtype, use_simdgroup)) return 0;
    return w4a16_gemm_workspace(M, N, K, dtype, use_split_k);
}

int tl_quantized_matmul(const void *scales, const void *biases, const void *a, const void *b, void *out, int M, int N,
                        int K, int dtype, int use_simdgroup, int use_split_k, void *workspace, size_t workspace_bytes,
                        void *stream) {
    if (dtype != TL_F16 && dtype != TL_BF16) return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
    if (M < 0 || N <= 0 || K < 0) return fail(TL_EINVAL, "quantized_matmul: negative dimension");
    if (N % 128 != 0) return fail(TL_EINVAL, "quantized_matmul: N must be divisible by group_size");
    if (!scales || !biases || !a || !b || !out) {
        if (M == 0 || K == 0) return TL_OK;
        return fail(TL_EINVAL, "quantized_matmul: null pointer");
    }
    cudaStream_t st = as_stream(stream);
    if (!use_simdgroup) return launch_w4a16_vanilla(scales, biases, a, b, out, M, N, K, dtype, st);
    if (M > 0 && K > 0 && use_skinny_kernel(M, N, K, dtype, use_simdgroup))
        return launch_w4a16_skinny(scales, biases, a, b, out, nullptr, M, N, K, TL_EPI_NONE, dtype, workspace, workspace_bytes, st);
    if (use_stream_kernel(M, N, K, dtype, use_simdgroup))
        return launch_w4a16_stream(scales, biases, a, b, out, M, N, K, dtype, st);
    return launch_w4a16_gemm(scales, biases, a, b, out, M, N, K, dtype, use_split_k, workspace, workspace_bytes, st);
}

int tl_quantized_embedding(const void *indices, const void *scales, const void *biases, const void *weight, void *out,
                           int tokens, int vocab, int dim, int dtype, void *stream) {
    if (tokens < 0 || vocab <= 0 || dim <= 0 || dim % 128 != 0)
        return fail(TL_EINVAL, "quantized_embedding: expected 4-bit weights with group size 128");
    if (tokens == 0) return TL_OK;
    if (!indices || !scales || !biases || !weight || !out) return fail(TL_EINVAL, "quantized_embedding: null pointer");
    return launch_quantized_embedding(indices, scales, biases, weight, out, tokens, vocab, dim, dtype, as_stream(stream));
}

// ----------------------------------------------------- fused model ops ----
int tl_rms_norm(const void *x, const void *weight, void *out, int rows, int dim, float eps, int dtype, void *stream) {
    if (!float_dtype(dtype)) return fail(TL_EDTYPE, "rms_norm: expected float32, float16, or bfloat16");
    if (rows < 0 || dim <= 0) return fail(TL_EINVAL, "rms_norm: bad shape");
    if (rows == 0) return TL_OK;
    if (!x || !weight || !out) return fail(TL_EINVAL, "rms_norm: null pointer");
    return launch_rms_norm(x, weight, out, rows, dim, eps, dtype, as_stream(stream));
}

int tl_rope(const void *x, const int32_t *offsets, void *out, int B, int L, int H, int D, int dims, float base,
            int traditional, int dtype, void *stream) {
    if (!float_dtype(dtype)) return fail(TL_EDTYPE, "rope: expected float32, float16, or bfloat16");
    if (B < 0 || L < 0 || H < 0 || D <= 0) return fail(TL_EINVAL, "rope: expected x=[B,L,H,D] and one int32 offset per batch row");
    if (dims <= 0 || dims > D || dims % 2 != 0)
        return fail(TL_EINVAL, "rope: dims must be positive, even, and no larger than the head dimension");
    if (static_cast<long long>(B) * L * H == 0) return TL_OK;
    if (!x || !offsets || !out) return fail(TL_EINVAL, "rope: null pointer");
    return launch_rope(x, offsets, out, B, L, H, D, dims, base, traditional, dtype, as_stream(stream));
}

int tl_swiglu(const void *gate, const void *up, void *out, long long size, int dtype, void *stream) {
    if (!float_dtype(dtype)) return fail(TL_EDTYPE, "swiglu: expected float32, float16, or bfloat16");
    if (size < 0) return fail(TL_EINVAL, "swiglu: negative size");
    if (size == 0) return TL_OK;
    if (!gate || !up || !out) return fail(TL_EINVAL, "swiglu: null pointer");
    return launch_swiglu(gate, up, out, size, dtype, as_stream(stream));
}

int tl_add(const void *a, const void *b, void *out, long long size, int dtype, void *stream) {
    if (!float_dtype(dtype)) return fail(TL_EDTYPE, "add: expected float32, float16, or bfloat16");
    if (size < 0) return fail(TL_EINVAL, "add: negative size");
    if (size == 0) return TL_OK;
    if (!a || !b || !out) return fail(TL_EINVAL, "add: null pointer");
    return launch_add(a, b, out, size, dtype, as_stream(stream));
}

int tl_decode_attention(const void *q, const void *k, const void *v, const float *mask, void *out, int q_rows, int L,
                        int S, int D, int num_heads, int num_kv_heads, float scale, int is_causal, int has_mask,
                        int dtype, void *stream) {
    if (!float_dtype(dtype)) return fail(TL_EDTYPE, "decode_attention: expected float32, float16, or bfloat16");
    if (q_rows < 0 || L < 0 || S < 0 || D <= 0 || D > 256 || num_heads <= 0 || num_kv_heads <= 0 ||
        num_heads % num_kv_heads != 0 || q_rows % num_heads != 0)
        return fail(TL_EINVAL, "decode_attention: incompatible attention shapes");
    if (static_cast<long long>(q_rows) * L == 0) return TL_OK;
    if (!q || !k || !v || !out || (has_mask && !mask)) return fail(TL_EINVAL, "decode_attention: null pointer");
    return launch_decode_attention(q, k, v, mask, out, q_rows, L, S, D, num_heads, num_kv_heads, scale, is_causal,
                                   has_mask, dtype, as_stream(stream));
}

// --------------------------------------------------------------- paged KV --
int tl_paged_cache_update(void *pages, const void *values, int num_pages, int heads, int page_size, int head_dim,
                          int length, int page_id, int start, int dtype, void *stream) {
    if (dtype != TL_F32 && dtype != TL_BF16)
        return fail(TL_EDTYPE, "paged_cache_update: pages and values must have the same float32 or bfloat16 dtype");
    if (num_pages <= 0 || heads <= 0 || page_size <= 0 || head_dim <= 0 || length < 0)
        return fail(TL_EINVAL, "paged_cache_update: expected pages [P, H, page_size, D] and values [1, H, length, D]");
    if (page_id < 0 || page_id >= num_pages || start < 0 || start + length > page_size)
        return fail(TL_EINVAL, "paged_cache_update: destination slice is outside page storage");
    if (length == 0) return TL_OK;
    if (!pages || !values) return fail(TL_EINVAL, "paged_cache_update: null pointer");
    return launch_paged_cache_update(pages, values, heads, page_size, head_dim, length, page_id, start, dtype,
                                     as_stream(stream));
}

int tl_paged_cache_append_decode(void *key_pages, void *value_pages, const void *keys, const void *values,
                                 const int32_t *block_table, const int32_t *context_lens, int batch, int num_pages,
                                 int heads, int page_size, int head_dim, int max_pages, int dtype, void *stream) {
    if (dtype != TL_F32 && dtype != TL_BF16)
        return fail(TL_EDTYPE, "paged_cache_append_decode: float32 or bfloat16 pages required");
    if (batch < 0 || num_pages <= 0 || heads <= 0 || page_size <= 0 || head_dim <= 0 || max_pages <= 0)
        return fail(TL_EINVAL, "paged_cache_append_decode: bad shape");
    if (batch == 0) return TL_OK;
    if (!key_pages || !value_pages || !keys || !values || !block_table || !context_lens)
        return fail(TL_EINVAL, "paged_cache_append_decode: null pointer");
    return launch_paged_cache_append_decode(key_pages, value_pages, keys, values, block_table, context_lens, batch,
                                            num_pages, heads, page_size, head_dim, max_pages, dtype, as_stream(stream));
}

int tl_paged_cache_append_chunk(void *key_pages, void *value_pages, const void *keys, const void *values,
                                const tl_page_span_list *spans, int num_pages, int heads, int page_size, int head_dim,
                                long long src_head_stride, long long src_token_stride, int dtype, void *stream) {
    if (dtype != TL_F32 && dtype != TL_BF16) return fail(TL_EDTYPE, "paged_cache_append_chunk: float32 or bfloat16 pages required");
    if (!spans || spans->n < 0 || spans->n > TL_PAGE_SPANS) return fail(TL_EINVAL, "paged_cache_append_chunk: bad span list");
    if (num_pages <= 0 || heads <= 0 || page_size <= 0 || head_dim <= 0 || src_token_stride < head_dim)
        return fail(TL_EINVAL, "paged_cache_append_chunk: bad shape");
    if (spans->n == 0) return TL_OK;
    if (!key_pages || !value_pages || !keys || !values) return fail(TL_EINVAL, "paged_cache_append_chunk: null pointer");
    for (int i = 0; i < spans->n; ++i)
        if (spans->page_id[i] < 0 || spans->page_id[i] >= num_pages || spans->start[i] < 0 || spans->count[i] < 0 || spans->src[i] < 0 ||
            spans->start[i] + spans->count[i] > page_size)
            return fail(TL_EINVAL, "paged_cache_append_chunk: destination slice is outside page storage");
    return launch_paged_cache_append_chunk(key_pages, value_pages, keys, values, *spans, heads, page_size, head_dim, src_head_stride,
                                           src_token_stride, dtype, as_stream(stream));
}

size_t tl_paged_attention_workspace(int rows, int L, int D, int num_kv_heads, int num_heads, int dtype) {
    if (L > 8) return 0;
    return paged_decode_workspace(rows, L, D, num_kv_heads, num_heads, dtype);
}

int tl_paged_attention(const void *q, const void *key_pages, const void *value_pages, const int32_t *block_table,
                       const int32_t *context_lens, void *out, int rows, int L, int D, int num_pages, int page_size,
                       int max_pages, float scale, int is_causal, int num_kv_heads, int num_heads, int dtype,
                       void *workspace, size_t workspace_bytes, void *stream) {
    if (dtype != TL_F32 && dtype != TL_BF16)
        return fail(TL_EDTYPE, "paged_attention: q, key_pages, and value_pages must have the same float32 or bfloat16 dtype");
    if (num_heads <= 0 || num_kv_heads <= 0 || num_heads % num_kv_heads != 0)
        return fail(TL_EINVAL, "paged_attention: num_heads must be divisible by num_kv_heads");
    if (rows < 0 || rows % num_heads != 0) return fail(TL_EINVAL, "paged_attention: q.shape[0] must be divisible by num_heads");
    if (D <= 0 || D > 128) return fail(TL_EINVAL, "paged_attention: head dimension must be in the range [1, 128]");
    if (L < 0 || num_pages <= 0 || page_size <= 0 || max_pages <= 0) return fail(TL_EINVAL, "paged_attention: bad shape");
    if (L > 8 && dtype == TL_BF16 && D != 128)
        return fail(TL_EINVAL, "paged_attention: bfloat16 prefill requires head dimension 128");
    if (static_cast<long long>(rows) * L == 0) return TL_OK;
    if (!q || !key_pages || !value_pages || !block_table || !context_lens || !out)
        return fail(TL_EINVAL, "paged_attention: null pointer");
    cudaStream_t st = as_stream(stream);
    if (L <= 8)
        return launch_paged_decode(q, key_pages, value_pages, block_table, context_lens, out, rows, L, D, num_pages,
                                   page_size, max_pages, scale, is_causal, num_kv_heads, num_heads, dtype, workspace,
                                   workspace_bytes, st);
    return launch_paged_prefill(q, key_pages, value_pages, block_table, context_lens, out, rows, L, D, num_pages,
                                page_size, max_pages, scale, is_causal, num_kv_heads, num_heads, dtype, st);
}

int tl_paged_attention_token_major(const void *q, const void *key_pages, const void *value_pages, const int32_t *block_table,
                                   const int32_t *context_lens, void *out, int rows, int L, int num_pages, int page_size, int max_pages,
                                   float scale, int is_causal, int num_kv_heads, int num_heads, void *stream) {
    if (num_heads <= 0 || num_kv_heads <= 0 || num_heads % num_kv_heads != 0 || rows < 0 || rows % num_heads != 0 || L < 0 || num_pages <= 0 ||
        page_size <= 0 || max_pages <= 0)
        return fail(TL_EINVAL, "paged_attention_token_major: bad shape");
    if (static_cast<long long>(rows) * L == 0) return TL_OK;
    if (!q || !key_pages || !value_pages || !block_table || !context_lens || !out) return fail(TL_EINVAL, "paged_attention: null pointer");
    if (!paged_prefill_tc_supported(L, num_pages, page_size, num_kv_heads, num_heads) || !aligned16(q) || !aligned16(out) || !aligned16(key_pages) ||
        !aligned16(value_pages))
        return fail(TL_EINVAL, "paged_attention_token_major: needs the tcgen05 kernel (bf16, D = 128, pages a multiple of 64 slots)");
    return launch_paged_prefill_tc(q, key_pages, value_pages, block_table, context_lens, out, rows, L, num_pages, page_size, max_pages, scale,
                                   is_causal, num_kv_heads, num_heads, false, nullptr, 0, as_stream(stream), true);
}

// ------------------------------------------------------------------ misc --
size_t tl_argmax_workspace(int rows, int vocab) { return argmax_workspace(rows, vocab); }

int tl_argmax(const void *logits, int32_t *out_tokens, int rows, int vocab, int dtype, void *workspace,
              size_t workspace_bytes, void *stream) {
    if (!float_dtype(dtype)) return fail(TL_EDTYPE, "argmax: expected float32, float16, or bfloat16");
    if (rows < 0 || vocab <= 0) return fail(TL_EINVAL, "argmax: bad shape");
    if (rows == 0) return TL_OK;
    if (!logits || !out_tokens) return fail(TL_EINVAL, "argmax: null pointer");
    return launch_argmax(logits, out_tokens, rows, vocab, dtype, workspace, workspace_bytes, as_stream(stream));
}

size_t tl_quantized_matmul_fused_workspace(int M, int N, int K, int lda, int prologue, int dtype) {
    if (prologue == TL_PRO_NONE && lda == N && use_skinny_kernel(M, N, K, dtype, 1)) return w4a16_skinny_workspace(M, N, K);
    return 0;
}

int tl_quantized_matmul_fused(const void *scales, const void *biases, const void *b, void *out, const void *p0,
                              const void *p1, const void *residual, int M, int N, int K, int lda, int prologue,
                              int epilogue, float eps, int dtype, void *workspace, size_t workspace_bytes, void *stream) {
    if (dtype != TL_F16 && dtype != TL_BF16) return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
    if (M < 0 || N <= 0 || K < 0 || lda < N) return fail(TL_EINVAL, "quantized_matmul_fused: bad shape");
    if (N % 128 != 0) return fail(TL_EINVAL, "quantized_matmul: N must be divisible by group_size");
    if (prologue < TL_PRO_NONE || prologue > TL_PRO_SWIGLU || epilogue < TL_EPI_NONE || epilogue > TL_EPI_SWIGLU_PAIRS)
        return fail(TL_EINVAL, "quantized_matmul_fused: unknown prologue/epilogue");
    if (epilogue == TL_EPI_SWIGLU_PAIRS && K % 16 != 0)
        return fail(TL_EINVAL, "quantized_matmul_fused: interleaved gate|up rows need K %% 16 == 0");
    if (M == 0 || K == 0) return TL_OK;
    if (!scales || !biases || !b || !out || !p0 || (prologue != TL_PRO_NONE && !p1) || (epilogue == TL_EPI_RESIDUAL && !residual))
        return fail(TL_EINVAL, "quantized_matmul_fused: null pointer");
    if (prologue == TL_PRO_NONE && lda == N && use_skinny_kernel(M, N, K, dtype, 1))
        return launch_w4a16_skinny(scales, biases, p0, b, out, residual, M, N, K, epilogue, dtype, workspace, workspace_bytes, as_stream(stream));
    return launch_w4a16_fused(scales, biases, b, out, p0, p1, residual, M, N, K, lda, prologue, epilogue, eps, dtype,
                              as_stream(stream));
}

int tl_quantized_matmul_residual_norm(const void *scales, const void *biases, const void *b, void *out, const void *p0, const void *residual,
                                      const void *norm_weight, void *normed_out, int M, int N, int K, float norm_eps, int dtype, void *workspace,
                                      size_t workspace_bytes, void *stream) {
    if (!norm_weight || !normed_out) return fail(TL_EINVAL, "quantized_matmul_residual_norm: null pointer");
    if (dtype != TL_F16 && dtype != TL_BF16) return fail(TL_EDTYPE, "quantized_matmul: scales must be float16 or bfloat16");
    if (M < 0 || N <= 0 || K < 0 || N % 128 != 0) return fail(TL_EINVAL, "quantized_matmul_residual_norm: bad shape");
    if (M == 0 || K == 0) return TL_OK;
    if (!scales || !biases || !b || !out || !p0 || !residual) return fail(TL_EINVAL, "quantized_matmul_residual_norm: null pointer");
    bool norm_done = false;
    int rc;
    if (use_skinny_kernel(M, N, K, dtype, 1))
        rc = launch_w4a16_skinny(scales, biases, p0, b, out, residual, M, N, K, TL_EPI_RESIDUAL, dtype, workspace, workspace_bytes, as_stream(stream),
                                 norm_weight, norm_eps, normed_out, &norm_done);
    else
        rc = launch_w4a16_fused(scales, biases, b, out, p0, nullptr, residual, M, N, K, N, TL_PRO_NONE, TL_EPI_RESIDUAL, 0.f, dtype, as_stream(stream));
    if (rc != TL_OK || norm_done) return rc;
    return launch_rms_norm(out, norm_weight, normed_out, M, K, norm_eps, dtype, as_stream(stream));
}

int tl_qkv_project_rope_append(const void *scales, const void *biases, const void *b, const void *p0, void *qkv_scratch, const void *q_norm_weight,
                               const void *k_norm_weight, const int32_t *offsets, const int32_t *block_table, const int32_t *context_lens, void *q_out,
                               void *key_pages, void *value_pages, int rows, int N, int num_heads, int num_kv_heads, int head_dim, float base, float eps,
                               int num_pages, int page_size, int max_pages, int chunk, int dtype, void *workspace, size_t workspace_bytes, void *stream) {
    if (dtype != TL_BF16) return fail(TL_EDTYPE, "qkv_project_rope_append: bfloat16 required");
    if (rows < 0 || N <= 0 || N % 128 != 0 || num_heads <= 0 || num_kv_heads <= 0 || head_dim <= 0 || head_dim % 2 != 0 || num_pages <= 0 ||
        page_size <= 0 || max_pages <= 0)
        return fail(TL_EINVAL, "qkv_project_rope_append: bad shape");
    if (rows == 0) return TL_OK;
    if (!scales || !biases || !b || !p0 || !qkv_scratch || !q_norm_weight || !k_norm_weight || !offsets || !block_table || !context_lens || !q_out ||
        !key_pages || !value_pages)
        return fail(TL_EINVAL, "qkv_project_rope_append: null pointer");
    const int K = (num_heads + 2 * num_kv_heads) * head_dim;
    cudaStream_t st = as_stream(stream);
    int planes = 1;
    int rc;
    if (use_skinny_kernel(rows, N, K, dtype, 1) && qkv_planes_rope_supported(num_heads, num_kv_heads, head_dim, dtype))
        rc = launch_w4a16_skinny(scales, biases, p0, b, qkv_scratch, nullptr, rows, N, K, TL_EPI_NONE, dtype, workspace, workspace_bytes, st, nullptr, 0.f,
                                 nullptr, nullptr, &planes);
    else
        rc = tl_quantized_matmul_fused(scales, biases, b, qkv_scratch, p0, nullptr, nullptr, rows, N, K, N, TL_PRO_NONE, TL_EPI_NONE, 0.f, dtype,
                                       workspace, workspace_bytes, stream);
    if (rc != TL_OK) return rc;
    if (planes > 1)  // the split-reduction planes go straight into the norm / RoPE / append kernel: q|k|v is never written
        return launch_qkv_planes_rope_append(static_cast<const float *>(workspace), planes, q_norm_weight, k_norm_weight, offsets, block_table,
                                             context_lens, q_out, key_pages, value_pages, rows, num_heads, num_kv_heads, base, eps, num_pages,
                                             page_size, max_pages, st, chunk != 0);
    return launch_decode_qk_norm_rope_append(qkv_scratch, q_norm_weight, k_norm_weight, offsets, block_table, context_lens, q_out, key_pages,
                                             value_pages, rows, num_heads, num_kv_heads, head_dim, base, eps, num_pages, page_size, max_pages, dtype,
                                             st, chunk != 0);
}

int tl_decode_qk_norm_rope_append(const void *qkv, const void *q_norm_weight, const void *k_norm_weight,
                                  const int32_t *offsets, const int32_t *block_table, const int32_t *context_lens,
                                  void *q_out, void *key_pages, void *value_pages, int batch, int num_heads,
                                  int num_kv_heads, int head_dim, float base, float eps, int num_pages, int page_size,
                                  int max_pages, int dtype, void *stream) {
    if (dtype != TL_F32 && dtype != TL_BF16) return fail(TL_EDTYPE, "decode_qk_norm_rope_append: bfloat16 or float32 required");
    if (batch < 0 || num_heads <= 0 || num_kv_heads <= 0 || head_dim <= 0 || head_dim % 2 != 0 || head_dim > 512 ||
        num_pages <= 0 || page_size <= 0 || max_pages <= 0)
        return fail(TL_EINVAL, "decode_qk_norm_rope_append: bad shape");
    if (batch == 0) return TL_OK;
    if (!qkv || !q_norm_weight || !k_norm_weight || !offsets || !block_table || !context_lens || !q_out || !key_pages || !value_pages)
        return fail(TL_EINVAL, "decode_qk_norm_rope_append: null pointer");
    return launch_decode_qk_norm_rope_append(qkv, q_norm_weight, k_norm_weight, offsets, block_table, context_lens, q_out,
                                             key_pages, value_pages, batch, num_heads, num_kv_heads, head_dim, base, eps,
                                             num_pages, page_size, max_pages, dtype, as_stream(stream));
}

int tl_chunk_qk_norm_rope_append(const void *qkv, const void *q_norm_weight, const void *k_norm_weight, const int32_t *offsets,
                                 const int32_t *block_table_row, const int32_t *context_lens, void *q_out, void *key_pages,
                                 void *value_pages, int tokens, int num_heads, int num_kv_heads, int head_dim, float base, float eps,
                                 int num_pages, int page_size, int max_pages, int dtype, void *stream) {
    if (dtype != TL_F32 && dtype != TL_BF16) return fail(TL_EDTYPE, "chunk_qk_norm_rope_append: bfloat16 or float32 required");
    if (tokens < 0 || num_heads <= 0 || num_kv_heads <= 0 || head_dim <= 0 || head_dim % 2 != 0 || head_dim > 512 || num_pages <= 0 ||
        page_size <= 0 || max_pages <= 0)
        return fail(TL_EINVAL, "chunk_qk_norm_rope_append: bad shape");
    if (tokens == 0) return TL_OK;
    if (!qkv || !q_norm_weight || !k_norm_weight || !offsets || !block_table_row || !context_lens || !q_out || !key_pages || !value_pages)
        return fail(TL_EINVAL, "chunk_qk_norm_rope_append: null pointer");
    return launch_decode_qk_norm_rope_append(qkv, q_norm_weight, k_norm_weight, offsets, block_table_row, context_lens, q_out, key_pages,
                                             value_pages, tokens, num_heads, num_kv_heads, head_dim, base, eps, num_pages, page_size,
                                             max_pages, dtype, as_stream(stream), true);
}

size_t tl_decode_attention_fused_workspace(int batch, int num_heads, int num_kv_heads) {
    if (batch < 1 || num_heads < 1 || num_kv_heads < 1) return 0;
    return decode_attention_fused_workspace(batch, num_heads, num_kv_heads);
}

int tl_decode_attention_fused(const void *qkv, const void *q_norm_weight, const void *k_norm_weight, const int32_t *offsets,
                              const int32_t *block_table, const int32_t *context_lens, const double *rope_inv_freq,
                              void *key_pages, void *value_pages, void *out, float *workspace, int batch, int num_heads,
                              int num_kv_heads, int head_dim, float eps, float scale, int num_pages, int page_size,
                              int max_pages, int max_context, int dtype, void *stream) {
    if (batch == 0) return TL_OK;
    if (!qkv || !q_norm_weight || !k_norm_weight || !offsets || !block_table || !context_lens || !rope_inv_freq || !key_pages ||
        !value_pages || !out || !workspace)
        return fail(TL_EINVAL, "decode_attention_fused: null pointer");
    if (batch < 0 || page_size < 1 || max_pages < 1 || num_pages < 0) return fail(TL_EINVAL, "decode_attention_fused: bad sizes");
    return launch_decode_attention_fused(qkv, q_norm_weight, k_norm_weight, offsets, block_table, context_lens, rope_inv_freq,
                                         key_pages, value_pages, out, workspace, batch, num_heads, num_kv_heads, head_dim, eps,
                                         scale, num_pages, page_size, max_pages, max_context, dtype, as_stream(stream));
}

#if TL_TRACE
extern "C" int tl_debug_trace(unsigned long long *device_events, unsigned int *device_count, unsigned int capacity) {
    trace_bind_matvec(device_events, device_count, capacity);
    trace_bind_attention(device_events, device_count, capacity);
    trace_bind_skinny(device_events, device_count, capacity);
    trace_bind_gemm(device_events, device_count, capacity);
    trace_bind_gemm2(device_events, device_count, capacity);
    return TL_OK;
}
#endif

int tl_set_gemm_pairs(int mode) {
    set_gemm_pairs(mode);
    return TL_OK;
}

int tl_set_pdl(int enabled) {
    set_use_pdl(enabled != 0);
    return TL_OK;
}

int tl_decode_advance(int32_t *tokens, const int32_t *next_tokens, int32_t *offsets, int32_t *context_lens,
                      int32_t *out_log, int32_t *step_counter, int batch, int log_capacity, void *stream) {
    if (batch <= 0 || log_capacity < 0) return fail(TL_EINVAL, "decode_advance: bad shape");
    if (!tokens || !next_tokens || !offsets || !context_lens || !out_log || !step_counter)
        return fail(TL_EINVAL, "decode_advance: null pointer");
    return launch_decode_advance(tokens, next_tokens, offsets, context_lens, out_log, step_counter, batch, log_capacity,
                                 as_stream(stream));
}

}  // extern "C"
