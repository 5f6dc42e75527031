/*
 * tiny_llm_b200.h - C ABI of the B200 (sm_100a) backend for tiny-llm's Qwen3
 * W4A16 inference hot path.
 *
 * This is the drop-in boundary: one launcher per native primitive of the
 * reference's extension (declared in
 * /root/reference/src/extensions_ref/src/tiny_llm_ext.h:10-141 and exported to
 * Python by /root/reference/src/extensions_ref/bindings.cpp:14-46).  Where the
 * reference builds a lazy `mx::array` whose `eval_gpu` encodes a Metal
 * dispatch, a launcher here enqueues sm_100a kernels on the caller's CUDA
 * stream.  Conventions:
 *
 *   - plain device pointers and sizes; no torch / MLX types;
 *   - the caller owns every buffer (outputs and workspaces included); a
 *     launcher never allocates, frees or synchronises, so it is legal inside
 *     CUDA-graph capture;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default);
 *   - return 0 on success, a negative TL_E* code otherwise; the message is
 *     available from tl_last_error() (thread-local); nothing throws across
 *     the boundary;
 *   - `dtype` is one of TL_F32 / TL_F16 / TL_BF16 and names the activation /
 *     storage type; all accumulation is fp32.
 *
 * Naming quirk kept from the reference (quantized_matmul.cpp:125-127):
 * in quantized_matmul `a` is [M, N] with N the REDUCTION length, `b` is
 * [K, N/8] packed words with K the number of OUTPUT features, out is [M, K].
 */
#ifndef TINY_LLM_B200_H
#define TINY_LLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { TL_F32 = 0, TL_F16 = 1, TL_BF16 = 2 };

enum {
    TL_OK = 0,
    TL_EINVAL = -1,      /* bad argument (shape, alignment, unsupported combination) */
    TL_EDTYPE = -2,      /* dtype not supported by this primitive */
    TL_EWORKSPACE = -3,  /* workspace missing or too small */
    TL_ECUDA = -4,       /* CUDA runtime / driver error at launch */
    TL_ENODEVICE = -5    /* no sm_100 device */
};

/* Library identity / diagnostics. */
int tl_abi_version(void);
const char *tl_last_error(void);
/* Fills SM count and compute capability of the current device. */
int tl_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* Number of kernels this library has launched since load (bench bookkeeping). */
long long tl_launch_count(void);

/* ---- W4A16 (4-bit, group 128) projections -------------------------------
 * Replaces quantized_matmul (tiny_llm_ext.h:12-21, quantized_matmul.cpp:14-80,
 * eval_gpu :111-240).  scales/biases [K, N/128] (f16|bf16), a [M, N],
 * b [K, N/8] u32, out [M, K].  Kernel selection:
 *   use_simdgroup && M <= TL_MATVEC_REF_ROWS : weight-streaming tensor-core
 *       matvec, weights kept fp32-exact (reference: M <= 8 matvec, quantize.py:162-163);
 *   use_simdgroup && M <= 128                : swap-AB tcgen05 GEMM with the reduction split over
 *       CTAs (reference: quantized_matmul_splitk, quantized_matmul.metal:251-293), weights rounded to
 *       the activation dtype before the MMA, fp32 partial planes added in split order;
 *   use_simdgroup                            : 128 x 128-tile tcgen05 GEMM, weights rounded to
 *       the activation dtype before the MMA (reference: simdgroup tile);
 *   !use_simdgroup                           : scalar control kernel.
 * (Shapes a tensor-core kernel cannot take - odd alignments - fall back to the streaming kernel, which
 * handles up to TL_MATVEC_MAX_ROWS rows per pass.)
 * The split of the reduction is a scheduling decision of this backend (SURVEY 8a' item 12): it depends only on
 * (N, K), never on use_split_k, so a split request and a plain request run the very same kernel with
 * bit-identical results (tests_refsol/test_week_2_day_7.py:80-109).
 * workspace: tl_quantized_matmul_workspace() bytes (0 = none; fp32 partial planes of a split reduction,
 * no initialisation needed). */
#define TL_MATVEC_REF_ROWS 8
#define TL_MATVEC_MAX_ROWS 32
size_t tl_quantized_matmul_workspace(int M, int N, int K, int dtype, int use_simdgroup, int use_split_k);
int tl_quantized_matmul(const void *scales, const void *biases, const void *a, const void *b, void *out, int M,
                        int N, int K, int dtype, int use_simdgroup, int use_split_k, void *workspace,
                        size_t workspace_bytes, void *stream);

/* Replaces quantized_embedding (tiny_llm_ext.h:40-41, quantized_matmul.cpp:82-101,
 * :242-273).  indices int32/uint32 [tokens] (bit pattern), weight [vocab, dim/8] u32,
 * scales/biases [vocab, dim/128], out [tokens, dim]. */
int tl_quantized_embedding(const void *indices, const void *scales, const void *biases, const void *weight,
                           void *out, int tokens, int vocab, int dim, int dtype, void *stream);

/* ---- fused model kernels (week2_kernels.cpp:36-84, :104-211) ------------- */
/* x [rows, dim], weight [dim], out [rows, dim]:  x * rsqrt(mean(x^2)+eps) * w */
int tl_rms_norm(const void *x, const void *weight, void *out, int rows, int dim, float eps, int dtype,
                void *stream);
/* x,out [B, L, H, D]; offsets int32 [B]; rotates the first `dims` of D. */
int tl_rope(const void *x, const int32_t *offsets, void *out, int B, int L, int H, int D, int dims, float base,
            int traditional, int dtype, void *stream);
/* out = gate / (1 + exp(-gate)) * up, `size` elements. */
int tl_swiglu(const void *gate, const void *up, void *out, long long size, int dtype, void *stream);
/* Dense-KV GQA attention: q,out [q_rows, L, D] with q_rows = B*Hq; k,v [B*Hkv, S, D];
 * mask fp32 [q_rows, L, S] when has_mask (ignored otherwise).  D <= 256. */
int tl_decode_attention(const void *q, const void *k, const void *v, const float *mask, void *out, int q_rows,
                        int L, int S, int D, int num_heads, int num_kv_heads, float scale, int is_causal,
                        int has_mask, int dtype, void *stream);

/* ---- paged KV (paged_attention.cpp:14-31, :38-70, :77-122, :129-225) ------ */
/* In-place: pages[page_id, :, start:start+length, :] = values[0]; pages [P,H,page,D],
 * values [1,H,length,D]. */
int tl_paged_cache_update(void *pages, const void *values, int num_pages, int heads, int page_size, int head_dim,
                          int length, int page_id, int start, int dtype, void *stream);
/* q,out [B*Hq, L, D]; pages [P, Hkv, page, D]; block_table int32 [B, max_pages]
 * (-1 padded); context_lens int32 [B] (post-append).  D <= 128; the tensor-core
 * prefill branch (L > 8, bf16) requires D == 128.  Rows that see no key are
 * written as zeros.  Workspace holds split-KV partials for the decode branch. */
size_t tl_paged_attention_workspace(int rows, int L, int D, int num_kv_heads, int num_heads, int dtype);
int tl_paged_attention(const void *q, const void *key_pages, const void *value_pages, const int32_t *block_table,
                       const int32_t *context_lens, void *out, int rows, int L, int D, int num_pages,
                       int page_size, int max_pages, float scale, int is_causal, int num_kv_heads, int num_heads,
                       int dtype, void *workspace, size_t workspace_bytes, void *stream);
/* The prefill form with the output written token-major: out [B, L, Hq * D] (the layout the o-projection takes; saves
 * the transpose copy of a chunked-prefill step).  bf16, D == 128, pages a multiple of 64 slots (the tcgen05 kernel);
 * TL_EINVAL otherwise - callers then use tl_paged_attention and transpose. */
int tl_paged_attention_token_major(const void *q, const void *key_pages, const void *value_pages, const int32_t *block_table,
                                   const int32_t *context_lens, void *out, int rows, int L, int num_pages, int page_size, int max_pages,
                                   float scale, int is_causal, int num_kv_heads, int num_heads, void *stream);

/* ---- B200 extensions behind the same per-op semantics ---------------------
 * Device-driven K/V append for a decode batch (the batched form of
 * paged_kv_cache.py:196-234 called once per request, kv_cache.py:191-199):
 * row b writes keys[b]/values[b] ([B, Hkv, 1, D]) to the slot of token
 * context_lens[b]-1, resolved through block_table; rows with context 0 are
 * skipped.  Page ids and slots come from device memory, so the call is
 * CUDA-graph replayable. */
int tl_paged_cache_append_decode(void *key_pages, void *value_pages, const void *keys, const void *values,
                                 const int32_t *block_table, const int32_t *context_lens, int batch,
                                 int num_pages, int heads, int page_size, int head_dim, int max_pages, int dtype,
                                 void *stream);
/* out = a + b (residual adds of qwen3_week3.py:204-206), `size` elements. */
int tl_add(const void *a, const void *b, void *out, long long size, int dtype, void *stream);
/* Greedy sampler of batch.py:8-13: out_tokens[r] = argmax(logits[r, :]) (the
 * log-softmax shift is argmax-invariant).  logits [rows, vocab], out int32 [rows]. */
int tl_argmax(const void *logits, int32_t *out_tokens, int rows, int vocab, int dtype, void *workspace,
              size_t workspace_bytes, void *stream);
size_t tl_argmax_workspace(int rows, int vocab);
/* Fused W4A16 projection for the decode hot loop: the weight-streaming kernel of
 * tl_quantized_matmul with the neighbouring element-wise operator folded in.
 * Rounding points are those of the unfused call sequence, so results are
 * bit-identical to it.
 *   prologue TL_PRO_NONE    : a = p0 [M, N]
 *            TL_PRO_RMSNORM : a = rms_norm(p0 [M, N], p1 [N], eps)      (then projected)
 *            TL_PRO_SWIGLU  : a = swiglu(p0 = gate [M, N], p1 = up [M, N])
 *   epilogue TL_EPI_NONE    : out = result
 *            TL_EPI_RESIDUAL: out = residual [M, K] + result
 *            TL_EPI_SWIGLU_PAIRS: out [M, K/2]; the weight rows are gate and up rows interleaved in
 *                             blocks of 8 (rows 16c..16c+7 = gate 8c..8c+7, rows 16c+8..16c+15 = up
 *                             8c..8c+7; K % 16 == 0) and out[m, 8c+r] = swiglu(T(gate), T(up)): the
 *                             MLP activation leaves the gate|up projection already combined
 *                             (week2_kernels.metal:115-116 applied to the rounded projection outputs)
 * lda is the row stride (elements) of p0 (and of p1 for SWIGLU), so gate/up may
 * be the two halves of one [M, 2N] buffer.
 * With 9 <= M <= 128, no prologue and lda == N the launch goes to the swap-AB tcgen05 kernel (same
 * epilogues; weights rounded to the activation dtype like every tensor-core path) and needs
 * tl_quantized_matmul_fused_workspace() bytes of workspace. */
enum { TL_PRO_NONE = 0, TL_PRO_RMSNORM = 1, TL_PRO_SWIGLU = 2 };
enum { TL_EPI_NONE = 0, TL_EPI_RESIDUAL = 1, TL_EPI_SWIGLU_PAIRS = 2 };
size_t tl_quantized_matmul_fused_workspace(int M, int N, int K, int lda, int prologue, int dtype);
int tl_quantized_matmul_fused(const void *scales, const void *biases, const void *b, void *out, const void *p0,
                              const void *p1, const void *residual, int M, int N, int K, int lda, int prologue,
                              int epilogue, float eps, int dtype, void *workspace, size_t workspace_bytes, void *stream);
/* out = residual + projection(p0) and, in the same call, normed_out = rms_norm(out, norm_weight, norm_eps): the
 * residual-stream projections (o, down) of qwen3_week3.py:204-206 followed by the RMSNorm that opens the next block
 * (:195-199).  Same rounding points as tl_quantized_matmul_fused(TL_EPI_RESIDUAL) then tl_rms_norm; with a split
 * reduction (9 <= M <= 128) the normalisation happens inside the kernel that adds the partial planes.  p0 is contiguous
 * [M, N]; workspace as for tl_quantized_matmul_fused(prologue TL_PRO_NONE). */
int tl_quantized_matmul_residual_norm(const void *scales, const void *biases, const void *b, void *out, const void *p0, const void *residual,
                                      const void *norm_weight, void *normed_out, int M, int N, int K, float norm_eps, int dtype, void *workspace,
                                      size_t workspace_bytes, void *stream);
/* q|k|v projection of `rows` activation rows (p0 [rows, N], packed weight rows = q heads | k heads | v heads) followed by
 * per-head q/k RMSNorm + RoPE + K/V append (tl_decode_qk_norm_rope_append, or its chunk form when chunk != 0).  With a
 * split reduction (9..128 rows) the partial planes feed the second kernel directly and q|k|v is never materialised;
 * otherwise qkv_scratch [rows, (Hq + 2 Hkv) * D] receives it.  Same results as the two calls.  bfloat16; workspace as
 * for tl_quantized_matmul_fused(prologue TL_PRO_NONE). */
int tl_qkv_project_rope_append(const void *scales, const void *biases, const void *b, const void *p0, void *qkv_scratch, const void *q_norm_weight,
                               const void *k_norm_weight, const int32_t *offsets, const int32_t *block_table, const int32_t *context_lens, void *q_out,
                               void *key_pages, void *value_pages, int rows, int N, int num_heads, int num_kv_heads, int head_dim, float base, float eps,
                               int num_pages, int page_size, int max_pages, int chunk, int dtype, void *workspace, size_t workspace_bytes, void *stream);
/* Decode step, L == 1: per-head q/k RMSNorm + RoPE + K/V append in one launch.
 * qkv [B, (Hq + 2*Hkv) * D] (q heads | k heads | v heads); q_out [B, Hq, D];
 * K/V rows land in the page slot of token context_lens[b]-1 (rows with context
 * 0 are skipped).  Non-traditional RoPE over the full head dimension. */
int tl_decode_qk_norm_rope_append(const void *qkv, const void *q_norm_weight, const void *k_norm_weight,
                                  const int32_t *offsets, const int32_t *block_table, const int32_t *context_lens,
                                  void *q_out, void *key_pages, void *value_pages, int batch, int num_heads,
                                  int num_kv_heads, int head_dim, float base, float eps, int num_pages, int page_size,
                                  int max_pages, int dtype, void *stream);
/* Prefill-chunk form of the same kernel: the rows of qkv [tokens, (Hq + 2*Hkv) * D] are consecutive tokens of ONE
 * request (one shared block-table row, int32 [max_pages]); offsets[t] is the RoPE position and context_lens[t]
 * the post-append length of token t (0 = padding row: nothing is appended), all DEVICE data, so a captured
 * chunk replays for any position.  q_out is [Hq, tokens, D] - the layout tl_paged_attention takes.  Replaces,
 * for one chunk, rms_norm x2 + rope x2 + the paged_cache_update calls of qwen3_week3.py:62-84. */
int tl_chunk_qk_norm_rope_append(const void *qkv, const void *q_norm_weight, const void *k_norm_weight,
                                 const int32_t *offsets, const int32_t *block_table_row, const int32_t *context_lens,
                                 void *q_out, void *key_pages, void *value_pages, int tokens, int num_heads,
                                 int num_kv_heads, int head_dim, float base, float eps, int num_pages, int page_size,
                                 int max_pages, int dtype, void *stream);
/* Chunk append (B200 extension of paged_cache_update, paged_attention.cpp:14-31): the rows of ONE
 * request's key/value chunk [1, H, L, D] (element strides src_head_stride / src_token_stride, unit
 * inner stride) are written into up to TL_PAGE_SPANS page slices in one launch:
 * pages[page_id[i], :, start[i]:start[i]+count[i], :] = chunk[0, :, src[i]:src[i]+count[i], :].
 * `spans` is a HOST pointer (passed by value to the kernel). */
#define TL_PAGE_SPANS 64
typedef struct tl_page_span_list {
    int32_t page_id[TL_PAGE_SPANS], start[TL_PAGE_SPANS], count[TL_PAGE_SPANS], src[TL_PAGE_SPANS];
    int32_t n;
} tl_page_span_list;
int tl_paged_cache_append_chunk(void *key_pages, void *value_pages, const void *keys, const void *values,
                                const tl_page_span_list *spans, int num_pages, int heads, int page_size, int head_dim,
                                long long src_head_stride, long long src_token_stride, int dtype, void *stream);
/* Fused decode attention, L == 1 (bf16, head_dim 128, <= 4 query heads per KV head): per-head
 * q/k RMSNorm + RoPE, append of the newest K/V row and paged GQA attention in ONE launch (plus a
 * merge launch when the context is split over several CTAs).  Replaces, with the same rounding
 * points, the call sequence rms_norm, rms_norm, rope, rope, paged_cache_update, paged_attention
 * of /root/reference/src/tiny_llm_ref/qwen3_week3.py:62-105.  qkv [B, (Hq + 2 Hkv) * 128] is the
 * row-concatenated projection output; context_lens are post-append; rope_inv_freq holds the 64
 * float64 frequencies base^(-i/64); out [B, Hq * 128]; max_context bounds every context_lens[b]
 * (it fixes the split count, so a captured launch stays valid as the contexts grow);
 * workspace: tl_decode_attention_fused_workspace() floats. */
size_t tl_decode_attention_fused_workspace(int batch, int num_heads, int num_kv_heads);
int tl_decode_attention_fused(const void *qkv, const void *q_norm_weight, const void *k_norm_weight,
                              const int32_t *offsets, const int32_t *block_table, const int32_t *context_lens,
                              const double *rope_inv_freq, void *key_pages, void *value_pages, void *out,
                              float *workspace, int batch, int num_heads, int num_kv_heads, int head_dim, float eps,
                              float scale, int num_pages, int page_size, int max_pages, int max_context, int dtype,
                              void *stream);
/* Programmatic dependent launch for the streaming kernels (on by default; 0 turns it off,
 * TL_PDL=0 in the environment does the same). */
int tl_set_pdl(int enabled);
/* Prefill GEMM on CTA pairs (tcgen05 cta_group::2, w4a16_gemm2.cu): 0 never, 1 where the pair grid fills the SMs
 * (default), 2 for every M > 256.  Same results either way (same rounding points); TL_GEMM2 sets the initial mode. */
int tl_set_gemm_pairs(int mode);
/* Device-side bookkeeping between two decode steps of a CUDA-graph loop (the
 * per-request `req.decode_done(token)` + `offset += 1` of batch.py:241-247 done
 * without a host round trip): for every ACTIVE row (context_lens > 0)
 *   tokens[b] = next_tokens[b]; offsets[b] += 1; context_lens[b] += 1;
 * and the sampled row is logged at out_log[*step_counter * batch + b]
 * (idle rows log -1); *step_counter is then incremented. */
int tl_decode_advance(int32_t *tokens, const int32_t *next_tokens, int32_t *offsets, int32_t *context_lens,
                      int32_t *out_log, int32_t *step_counter, int batch, int log_capacity, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TINY_LLM_B200_H */
