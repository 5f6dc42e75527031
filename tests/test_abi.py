"""The C-ABI shared library: it loads on a machine without a GPU, exports every
symbol include/tiny_llm_b200.h declares, and the Python shim refuses CPU
tensors (no CPU fallback).  No kernel is launched here."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

from extensions_b200 import tiny_llm_ext_b200 as ext

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "tiny_llm_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(tl_[a-z0-9_]+)\s*\(", body)))


def test_library_is_in_tree_and_loaded():
    path = ext.current_library_path()
    assert path is not None and path.exists()
    assert ROOT in path.parents, "the extension must be built in-tree (not in a JIT cache)"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(str(ext.current_library_path()))
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(ext.EXPORTED_SYMBOLS), "shim bindings and header disagree"
    assert lib.tl_abi_version() == 1


def test_reference_entry_points_keep_names_and_defaults():
    # /root/reference/src/extensions_ref/bindings.cpp:14-46
    import inspect

    expected = {
        "quantized_matmul": ["scales", "biases", "group_size", "bits", "a", "b", "transpose_b", "use_simdgroup", "use_split_k", "stream"],
        "quantized_embedding": ["indices", "scales", "biases", "weight", "group_size", "bits", "stream"],
        "rms_norm": ["x", "weight", "eps", "stream"],
        "rope": ["x", "offsets", "dims", "base", "traditional", "stream"],
        "swiglu": ["gate", "up", "stream"],
        "decode_attention": ["query", "key", "value", "mask", "scale", "is_causal", "has_mask", "num_heads", "num_kv_heads", "stream"],
        "paged_cache_update": ["pages", "values", "page_id", "start", "stream"],
        "paged_attention": ["query", "key_pages", "value_pages", "block_table", "context_lens", "scale", "is_causal", "num_kv_heads", "num_heads", "stream"],
    }
    for name, params in expected.items():
        sig = inspect.signature(getattr(ext, name))
        assert list(sig.parameters) == params, name
        assert sig.parameters["stream"].default is None
    qm = inspect.signature(ext.quantized_matmul).parameters
    assert (qm["transpose_b"].default, qm["use_simdgroup"].default, qm["use_split_k"].default) == (False, True, False)
    pa = inspect.signature(ext.paged_attention).parameters
    assert (pa["scale"].default, pa["is_causal"].default) == (1.0, False)
    assert inspect.signature(ext.rope).parameters["traditional"].default is False
    assert callable(ext.load_library)


def test_cpu_tensors_are_refused_gpu_only():
    bf = torch.bfloat16
    with pytest.raises(RuntimeError, match="rms_norm: the course extension is GPU-only"):
        ext.rms_norm(torch.zeros(2, 8, dtype=bf), torch.ones(8, dtype=bf), 1e-5)
    with pytest.raises(RuntimeError, match="swiglu: the course extension is GPU-only"):
        ext.swiglu(torch.zeros(4), torch.zeros(4))
    with pytest.raises(RuntimeError, match="quantized_matmul: the course extension is GPU-only"):
        ext.quantized_matmul(torch.zeros(4, 1, dtype=bf), torch.zeros(4, 1, dtype=bf), 128, 4, torch.zeros(1, 128, dtype=bf),
                             torch.zeros(4, 16, dtype=torch.int32), True)
    with pytest.raises(RuntimeError, match="paged_cache_update: the course extension is GPU-only"):
        ext.paged_cache_update(torch.zeros(2, 1, 4, 2), torch.zeros(1, 1, 1, 2), 0, 0)


def test_builder_checks_run_before_the_device_check():
    bf = torch.bfloat16
    with pytest.raises(RuntimeError, match="b must be transposed"):
        ext.quantized_matmul(torch.zeros(4, 1, dtype=bf), torch.zeros(4, 1, dtype=bf), 128, 4, torch.zeros(1, 128, dtype=bf),
                             torch.zeros(4, 16, dtype=torch.int32), False)
    with pytest.raises(RuntimeError, match="dims must be positive, even"):
        ext.rope(torch.zeros(1, 1, 1, 4), torch.zeros(1, dtype=torch.int32), 3, 10000.0)
    with pytest.raises(RuntimeError, match="destination slice is outside page storage"):
        ext.paged_cache_update(torch.zeros(2, 1, 4, 2), torch.zeros(1, 1, 2, 2), 0, 3)
    with pytest.raises(RuntimeError, match="mask must be float32"):
        ext.decode_attention(torch.zeros(1, 1, 4), torch.zeros(1, 1, 4), torch.zeros(1, 1, 4), torch.zeros(1, dtype=bf), 1.0, False, False, 1, 1)


def test_c_abi_reports_argument_errors_without_a_device():
    lib = ctypes.CDLL(str(ext.current_library_path()))
    lib.tl_last_error.restype = ctypes.c_char_p
    lib.tl_rms_norm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    assert lib.tl_rms_norm(None, None, None, 4, 8, 1e-5, 7, None) == -2  # TL_EDTYPE
    assert b"expected float32, float16, or bfloat16" in lib.tl_last_error()
    assert lib.tl_rms_norm(None, None, None, 4, 8, 1e-5, 2, None) == -1  # TL_EINVAL: null pointers
    assert lib.tl_rms_norm(None, None, None, 0, 8, 1e-5, 2, None) == 0  # empty input is a no-op


def test_gate_up_interleave_layout_and_span_record():
    """Host-side helpers of the B200 extensions: the row layout the SWIGLU_PAIRS epilogue expects,
    and the by-value span record of tl_paged_cache_append_chunk (4 x 64 int32 + count)."""
    import ctypes

    import torch

    from extensions_b200 import tiny_llm_ext_b200 as ext

    gate = torch.arange(32 * 3, dtype=torch.int32).reshape(32, 3)
    up = -gate - 1
    both = ext.interleave_gate_up(gate, up)
    assert both.shape == (64, 3)
    for c in range(4):
        assert torch.equal(both[16 * c : 16 * c + 8], gate[8 * c : 8 * c + 8])
        assert torch.equal(both[16 * c + 8 : 16 * c + 16], up[8 * c : 8 * c + 8])
    try:
        ext.interleave_gate_up(gate[:30], up[:30])
    except RuntimeError as exc:
        assert "rows % 8" in str(exc)
    else:
        raise AssertionError("rows not divisible by 8 must be rejected")
    assert ctypes.sizeof(ext.PageSpanList) == 4 * 64 * 4 + 4
    assert ext.EPI_SWIGLU_PAIRS == 2 and ext.PAGE_SPANS == 64
