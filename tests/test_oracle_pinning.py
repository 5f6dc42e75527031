"""Pin the CPU oracle before trusting it (CPU-only, no GPU needed).

MLX cannot be installed here and every reference primitive throws on CPU, so
there is no reference-produced float output.  What pins the oracle instead:

* the in-tree layout spec (quantize.py:103-121) through identity products;
* the cross-implementation equalities the reference asserts in its own tests,
  evaluated between two independently written restatements (``oracle.ops``,
  kernel-level, vs ``oracle.readable``, Week-1 readable path), with the
  reference's tolerances and - where it has them - its deterministic fixtures;
* closed-form cases.
Integer/structural literals of the reference tests are pinned in
``test_paged_kv_host.py`` / ``test_scheduler_host.py`` against the product's
host code (they involve no arithmetic the oracle could get wrong).
"""

import json
import math
from math import prod
from pathlib import Path

import pytest
import torch

from oracle import ops, readable

GOLDEN = Path(__file__).parent / "golden"
BF16 = torch.bfloat16


def close(a, b, rtol, atol):
    torch.testing.assert_close(a.to(torch.float32), b.to(torch.float32), rtol=rtol, atol=atol)


def pack_codes(codes: torch.Tensor) -> torch.Tensor:
    """[K, N] integer codes 0..15 -> [K, N/8] packed words (code i at bits 4i)."""
    K, N = codes.shape
    c = codes.to(torch.int64).reshape(K, N // 8, 8)
    words = (c << torch.arange(0, 32, 4)).sum(-1)
    return torch.where(words >= 2**31, words - 2**32, words).to(torch.int32).view(torch.uint32)


# ---- W4 layout ---------------------------------------------------------------
def test_nibble_order_is_least_significant_first():
    # word 0x76543210 must decode to 0,1,2,...,7 (quantize.py:113-115, metal :41-48)
    w = torch.tensor([[0x76543210]], dtype=torch.int32).view(torch.uint32)
    assert ops.unpack_nibbles(w).tolist() == [[0, 1, 2, 3, 4, 5, 6, 7]]
    w = torch.tensor([[-1]], dtype=torch.int32)  # 0xFFFFFFFF
    assert ops.unpack_nibbles(w).tolist() == [[15] * 8]


def test_dequantize_matches_the_affine_spec_exactly():
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(0, 16, (5, 256), generator=g)
    scales = (torch.randn(5, 2, generator=g) * 0.1).to(BF16)
    biases = torch.randn(5, 2, generator=g).to(BF16)
    got = ops.dequantize_weights(pack_codes(codes), scales, biases, 128, 4)
    want = (codes.float() * scales.float().repeat_interleave(128, 1) + biases.float().repeat_interleave(128, 1)).to(BF16)
    assert got.dtype == BF16 and torch.equal(got, want)


@pytest.mark.parametrize("flags", [dict(use_simdgroup=False), dict(use_simdgroup=True), dict(use_simdgroup=True, use_split_k=True)])
def test_identity_activations_return_the_dequantised_weights(flags):
    # test_week_2_day_3.py:74-118 (identity_matrix=True): eye(128) @ W^T == dequantised W^T
    g = torch.Generator().manual_seed(1)
    codes = torch.randint(0, 16, (5, 128), generator=g)
    scales = (torch.randn(5, 1, generator=g) * 0.1).to(BF16)
    biases = torch.randn(5, 1, generator=g).to(BF16)
    b = pack_codes(codes)
    out = ops.quantized_matmul(scales, biases, 128, 4, torch.eye(128, dtype=BF16), b, True, **flags)
    assert out.dtype == BF16 and tuple(out.shape) == (128, 5)
    assert torch.equal(out, ops.dequantize_weights(b, scales, biases, 128, 4).T)


def test_matvec_tiled_and_splitk_rounding_variants_agree_within_reference_tolerance():
    # test_week_2_day_6.py:30-48 (atol 1.0) and test_week_2_day_7.py:19-47 (atol 1.5)
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, 16, (1024, 2560), generator=g)
    scales = (torch.rand(1024, 20, generator=g) * 0.4).to(BF16)
    biases = (-7.5 * scales.float()).to(BF16)
    a = torch.randn(32, 2560, generator=g).to(BF16)
    b = pack_codes(codes)
    vanilla = ops.quantized_matmul(scales, biases, 128, 4, a, b, True, use_simdgroup=False)
    tiled = ops.quantized_matmul(scales, biases, 128, 4, a, b, True, use_simdgroup=True)
    split = ops.quantized_matmul(scales, biases, 128, 4, a, b, True, use_simdgroup=True, use_split_k=True)
    assert ops.reference_split_k(32, 2560, 1024) == 10  # 320 // 32 tiles, 2560 % (10*128) == 0
    close(tiled, vanilla, rtol=2e-2, atol=1.0)
    close(split, vanilla, rtol=2e-2, atol=1.5)


def test_reference_split_policy_falls_back_when_grid_is_full():
    # test_week_2_day_7.py:80-109: 128x256 @ 2560 rows -> 4*80 tiles >= 320 -> no split
    assert ops.reference_split_k(128, 256, 2560) == 1
    # 17 x 1032: 1 * 33 tiles -> 9, reduced until 2560 % (s*128) == 0 -> 5
    assert ops.reference_split_k(17, 2560, 1032) == 5


def test_quantized_matmul_validation_messages():
    s = torch.zeros(4, 1, dtype=BF16)
    a = torch.zeros(2, 128, dtype=BF16)
    b = torch.zeros(4, 16, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="b must be transposed"):
        ops.quantized_matmul(s, s, 128, 4, a, b, False)
    with pytest.raises(RuntimeError, match="bits must be 4"):
        ops.quantized_matmul(s, s, 128, 8, a, b, True)
    with pytest.raises(RuntimeError, match="group_size must be 128"):
        ops.quantized_matmul(s, s, 64, 4, a, b, True)
    with pytest.raises(RuntimeError, match="same dtype as scales"):
        ops.quantized_matmul(s, s, 128, 4, a.to(torch.float16), b, True)
    with pytest.raises(RuntimeError, match="one column per input group"):
        ops.quantized_matmul(torch.zeros(4, 2, dtype=BF16), torch.zeros(4, 2, dtype=BF16), 128, 4, a, b, True)


def test_quantized_embedding_equals_gather_then_dequantize():
    # test_week_2_day_3.py:24-51 with int32 and uint32 indices
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 16, (7, 256), generator=g)
    scales = (torch.randn(7, 2, generator=g) * 0.1).to(BF16)
    biases = torch.randn(7, 2, generator=g).to(BF16)
    w = pack_codes(codes)
    full = ops.dequantize_weights(w, scales, biases, 128, 4)
    for dtype in (torch.int32, torch.uint32):
        idx = torch.tensor([[1, 4]], dtype=torch.int32).view(dtype) if dtype == torch.uint32 else torch.tensor([[1, 4]], dtype=dtype)
        got = ops.quantized_embedding(idx, scales, biases, w, 128, 4)
        assert tuple(got.shape) == (1, 2, 256)
        assert torch.equal(got[0], full[[1, 4]])


# ---- fused kernels vs the readable path ---------------------------------------
def test_fast_rms_norm_matches_readable():
    # test_week_2_day_4.py:29-34
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 16, generator=g).to(BF16)
    w = torch.randn(16, generator=g).to(BF16)
    close(ops.rms_norm(x, w, 1e-5), readable.RMSNorm(16, w, eps=1e-5)(x), rtol=2e-2, atol=2e-2)


def test_rms_norm_closed_form():
    x = torch.tensor([[3.0, 4.0]], dtype=torch.float32)
    w = torch.tensor([2.0, 0.5], dtype=torch.float32)
    inv = 1.0 / math.sqrt((9 + 16) / 2 + 1e-6)
    close(ops.rms_norm(x, w, 1e-6), torch.tensor([[3 * inv * 2, 4 * inv * 0.5]]), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("offsets", [3, [3, 7]])
@pytest.mark.parametrize("traditional", [False, True])
def test_fast_rope_matches_readable(offsets, traditional):
    # test_week_2_day_4.py:37-51 (+ traditional layout of test_week_3_day_1.py:12-49)
    g = torch.Generator().manual_seed(5)
    B = 1 if isinstance(offsets, int) else len(offsets)
    x = torch.randn(B, 4, 2, 16, generator=g).to(BF16)
    off = [offsets] * B if isinstance(offsets, int) else offsets
    got = ops.rope(x, torch.tensor(off, dtype=torch.int32), 16, 10000, traditional)
    want = readable.RoPE(16, 32, base=10000, traditional=traditional)(x, [slice(o, o + 4) for o in off])
    close(got, want, rtol=2e-2, atol=2e-2)


def test_rope_closed_form_and_tail_passthrough():
    # one pair, angle = pos * base^0 = pos; dims < D copies the tail
    x = torch.tensor([[[[1.0, 0.0, 5.0, 6.0]]]], dtype=torch.float32)  # [1,1,1,4], dims=2
    out = ops.rope(x, torch.tensor([2], dtype=torch.int32), 2, 10000.0, False)
    close(out, torch.tensor([[[[math.cos(2.0), math.sin(2.0), 5.0, 6.0]]]]), rtol=1e-6, atol=1e-6)


def test_swiglu_matches_readable_expression():
    # test_week_2_day_4.py:54-57
    g = torch.Generator().manual_seed(6)
    gate = torch.randn(2, 4, 16, generator=g).to(BF16)
    up = torch.randn(2, 4, 16, generator=g).to(BF16)
    close(ops.swiglu(gate, up), readable.silu(gate) * up, rtol=5e-2, atol=1e-2)
    assert float(ops.swiglu(torch.zeros(1), torch.ones(1))) == 0.0


def test_decode_attention_matches_grouped_attention_on_the_reference_fixture_sweep():
    # test_week_2_day_5.py:119-163 - deterministic sin fixtures, no RNG involved.
    head_dim, query_heads = 128, 4

    def fixture(shape, phase):
        return torch.sin(torch.arange(prod(shape), dtype=torch.float32) * 0.017 + phase).reshape(shape).to(BF16)

    shapes = [(1, s) for s in (1, 31, 32, 127, 128, 129, 255, 256)] + [(8, s) for s in (8, 31, 32, 127, 128, 129, 255, 256)]
    golden = {}
    for L, S in shapes:
        for ratio in (1, 4):
            kv_heads = query_heads // ratio
            q = fixture((1, query_heads, L, head_dim), 0.1)
            k = fixture((1, kv_heads, S, head_dim), 0.7)
            v = fixture(tuple(k.shape), 1.3)
            explicit = torch.where(torch.arange(S) % 5 == 0, -2.0, 0.0).reshape(1, 1, 1, S)
            for mask in ("causal", explicit):
                causal = isinstance(mask, str)
                m3 = (
                    torch.zeros(1)
                    if causal
                    else torch.broadcast_to(mask, (1, query_heads, L, S)).reshape(query_heads, L, S).contiguous().float()
                )
                got = ops.decode_attention(
                    q.reshape(query_heads, L, head_dim), k.reshape(kv_heads, S, head_dim), v.reshape(kv_heads, S, head_dim),
                    m3, head_dim**-0.5, causal, not causal, query_heads, kv_heads,
                ).reshape(1, query_heads, L, head_dim)
                want = readable.scaled_dot_product_attention_grouped(q, k, v, head_dim**-0.5, mask)
                close(got, want, rtol=3e-2, atol=3e-2)
                golden[f"L{L}_S{S}_g{ratio}_{'causal' if causal else 'mask'}"] = float(got.float().sum())
    # regression anchor: checksums of the kernel-level restatement (self-generated, see golden/README)
    ref = json.loads((GOLDEN / "decode_attention_fixture_checksums.json").read_text())
    for key, value in golden.items():
        assert abs(value - ref[key]) <= 2e-2 * max(1.0, abs(ref[key])), key


# ---- paged attention ------------------------------------------------------------
def _paged_case(seed, page_size, lens, H=2, D=4, Hq=4, dtype=torch.float32, holes=False):
    g = torch.Generator().manual_seed(seed)
    pages_needed = [(n + page_size - 1) // page_size for n in lens]
    ids, nxt = [], 0
    for n in pages_needed:
        row = []
        for _ in range(n):
            row.append(nxt)
            nxt += 2 if holes else 1
        ids.append(row)
    P = max(nxt, 1)
    kp = torch.randn(P, H, page_size, D, generator=g).to(dtype)
    vp = torch.randn(P, H, page_size, D, generator=g).to(dtype)
    width = max(1, max(pages_needed))
    bt = torch.full((len(lens), width), -1, dtype=torch.int32)
    for b, row in enumerate(ids):
        bt[b, : len(row)] = torch.tensor(row, dtype=torch.int32)
    return kp, vp, bt, torch.tensor(lens, dtype=torch.int32), ids


def _dense(kp, vp, ids, n):
    k = torch.cat([kp[i] for i in ids], dim=1)[:, :n][None]
    v = torch.cat([vp[i] for i in ids], dim=1)[:, :n][None]
    return k, v


@pytest.mark.parametrize("L", [1, 3, 9])
def test_paged_attention_equals_dense_attention_on_noncontiguous_pages(L):
    # test_week_3_day_4.py:118-150, test_week_3_day_5.py:23-61 (non-contiguous page ids)
    kp, vp, bt, cl, ids = _paged_case(7, 4, [11], holes=True)
    g = torch.Generator().manual_seed(8)
    q = torch.randn(1, 4, L, 4, generator=g)
    got = ops.paged_attention(q.reshape(4, L, 4), kp, vp, bt, cl, 4**-0.5, True, 2, 4).reshape(1, 4, L, 4)
    k, v = _dense(kp, vp, ids[0], 11)
    want = readable.scaled_dot_product_attention_grouped(q, k, v, mask="causal")
    close(got, want, rtol=1e-5, atol=1e-5)


def test_paged_attention_idle_slot_is_exact_zero_and_rows_are_independent():
    # test_week_3_day_4.py:153-201: context_lens == [4, 0, 7], row [-1, -1]
    kp, vp, bt, cl, ids = _paged_case(9, 4, [4, 0, 7])
    g = torch.Generator().manual_seed(10)
    q = torch.randn(3, 4, 1, 4, generator=g)
    got = ops.paged_attention(q.reshape(12, 1, 4), kp, vp, bt, cl, 0.5, True, 2, 4).reshape(3, 4, 1, 4)
    assert bt[1].tolist() == [-1, -1]
    assert torch.count_nonzero(got[1]) == 0
    for b, n in ((0, 4), (2, 7)):
        k, v = _dense(kp, vp, ids[b], n)
        close(got[b : b + 1], readable.scaled_dot_product_attention_grouped(q[b : b + 1], k, v, 0.5, "causal"), 1e-5, 1e-5)


def test_paged_cache_update_writes_one_slice_in_place():
    pages = torch.zeros(3, 2, 4, 2)
    vals = torch.arange(8, dtype=torch.float32).reshape(1, 2, 2, 2)
    out = ops.paged_cache_update(pages, vals, 1, 1)
    assert out is pages
    assert torch.equal(pages[1, :, 1:3, :], vals[0])
    assert pages.sum() == vals.sum()
    with pytest.raises(RuntimeError, match="outside page storage"):
        ops.paged_cache_update(pages, vals, 1, 3)
    with pytest.raises(RuntimeError, match="outside page storage"):
        ops.paged_cache_update(pages, vals, 3, 0)


# ---------------------------------------------------------------------------------------------
# fp64 referee: the float operators of the oracle are fp32 restatements of
# the Metal kernels; MLX cannot run here, so their only other witness was a second restatement.
# Each operator is evaluated once more in float64 straight from its mathematical definition (no
# shared code with oracle/ops.py beyond the nibble unpacking spec of quantize.py:113-115) - an
# arithmetic-free third opinion: the oracle's bf16 result must be the correctly rounded fp64 value
# up to one output ulp: |err| <= 2^-8 |x| is half an ulp at the bottom of a bf16 binade, a full ulp is
# allowed where fp32 accumulation noise can flip a rounding.
def _ulp_close(got, want64, ulps=2.0, floor=1e-6):
    want = want64.to(torch.float64)
    err = (got.to(torch.float64) - want).abs()
    bound = ulps * 2.0**-8 * want.abs() + floor
    assert bool((err <= bound).all()), f"max excess {float((err - bound).max()):.3e}"


def _codes64(words, N):
    w = words.to(torch.int64) & 0xFFFFFFFF
    shifts = torch.arange(0, 32, 4, dtype=torch.int64)
    return ((w[..., None] >> shifts) & 0xF).reshape(words.shape[0], N).to(torch.float64)


def test_fp64_referee_quantized_matmul_and_embedding():
    g = torch.Generator().manual_seed(64)
    K, N, M = 48, 512, 5
    words = torch.randint(-(2**31), 2**31, (K, N // 8), dtype=torch.int64, generator=g).to(torch.int32)
    scales = (torch.randn(K, N // 128, generator=g) * 0.02).to(torch.bfloat16)
    biases = (-7.5 * scales.float() + torch.randn(K, N // 128, generator=g) * 0.02).to(torch.bfloat16)
    a = torch.randn(M, N, generator=g).to(torch.bfloat16)
    w64 = _codes64(words, N) * scales.double().repeat_interleave(128, dim=1) + biases.double().repeat_interleave(128, dim=1)
    got = ops.quantized_matmul(scales, biases, 128, 4, a, words, True, use_simdgroup=False)
    _ulp_close(got, a.double() @ w64.T, ulps=2.0, floor=2e-3 * float((a.double() @ w64.T).abs().max()) * 2.0**-8)
    idx = torch.tensor([[3, 47, 0]], dtype=torch.int32)
    emb = ops.quantized_embedding(idx, scales, biases, words, 128, 4)
    _ulp_close(emb[0], w64[idx[0].long()], ulps=1.01)


def test_fp64_referee_rms_norm_rope_swiglu():
    g = torch.Generator().manual_seed(65)
    x = (torch.randn(3, 7, 256, generator=g) * 3).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(256, generator=g)).to(torch.bfloat16)
    want = x.double() * torch.rsqrt((x.double() ** 2).mean(dim=-1, keepdim=True) + 1e-6) * w.double()
    _ulp_close(ops.rms_norm(x, w, 1e-6), want, ulps=1.01)
    gate = (torch.randn(4, 96, generator=g) * 4).to(torch.bfloat16)
    up = torch.randn(4, 96, generator=g).to(torch.bfloat16)
    want = gate.double() / (1 + torch.exp(-gate.double())) * up.double()
    _ulp_close(ops.swiglu(gate, up), want, ulps=1.01)
    B, L, H, D = 2, 5, 3, 64
    q = torch.randn(B, L, H, D, generator=g).to(torch.bfloat16)
    offsets = torch.tensor([0, 37], dtype=torch.int32)
    half = D // 2
    freq = 10000.0 ** (-torch.arange(half, dtype=torch.float64) / half)
    pos = offsets.double()[:, None] + torch.arange(L, dtype=torch.float64)[None, :]
    ang = pos[:, :, None, None] * freq[None, None, None, :]
    re, im = q.double()[..., :half], q.double()[..., half:]
    want = torch.cat([re * torch.cos(ang) - im * torch.sin(ang), im * torch.cos(ang) + re * torch.sin(ang)], dim=-1)
    # the reference forms the angle in fp32 (week2_kernels.metal:78-104): ~1e-6 relative angle noise at these positions
    _ulp_close(ops.rope(q, offsets, D, 10000.0), want, ulps=2.0, floor=2e-5)


def test_fp64_referee_attention_paged_and_dense():
    g = torch.Generator().manual_seed(66)
    Hq, Hkv, D, page = 4, 2, 32, 8
    lens = [19, 8]
    L = 3
    q = torch.randn(len(lens) * Hq, L, D, generator=g).to(torch.bfloat16)
    total_pages = 6
    kp = torch.randn(total_pages, Hkv, page, D, generator=g).to(torch.bfloat16)
    vp = torch.randn(total_pages, Hkv, page, D, generator=g).to(torch.bfloat16)
    bt = torch.tensor([[4, 1, 5], [2, -1, -1]], dtype=torch.int32)
    cl = torch.tensor(lens, dtype=torch.int32)
    scale = D**-0.5
    got = ops.paged_attention(q, kp, vp, bt, cl, scale, True, Hkv, Hq)
    for b, ctx in enumerate(lens):
        ids = bt[b, : (ctx + page - 1) // page].long()
        k = kp[ids].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :ctx].double()
        v = vp[ids].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :ctx].double()
        for h in range(Hq):
            for l in range(L):
                seen = ctx - L + l + 1  # bottom-right causal alignment (attention.py:24-27)
                s = (q[b * Hq + h, l].double() @ k[h // (Hq // Hkv), :seen].T) * scale
                p = torch.softmax(s, dim=-1)
                _ulp_close(got[b * Hq + h, l], p @ v[h // (Hq // Hkv), :seen], ulps=2.0, floor=1e-4)
    # dense decode attention with an explicit additive mask (week2_kernels.metal:119-235)
    S = 11
    kd = torch.randn(Hkv, S, D, generator=g).to(torch.bfloat16)
    vd = torch.randn(Hkv, S, D, generator=g).to(torch.bfloat16)
    qd = torch.randn(Hq, 2, D, generator=g).to(torch.bfloat16)
    mask = torch.where(torch.arange(S) % 3 == 0, -1.5, 0.0).reshape(1, 1, S).expand(Hq, 2, S).contiguous()
    got = ops.decode_attention(qd, kd, vd, mask, scale, False, True, Hq, Hkv)
    for h in range(Hq):
        s = (qd[h].double() @ kd[h // 2].double().T) * scale + mask[h].double()
        _ulp_close(got[h], torch.softmax(s, dim=-1) @ vd[h // 2].double(), ulps=2.0, floor=1e-4)
