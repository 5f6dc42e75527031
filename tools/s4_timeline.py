"""In-kernel timeline of the W4A16 streaming kernel (debug build with -DS4_PROF=1):

  TL_DEFINES="-DS4_PROF=1" TL_LIB_SUFFIX=_prof python tiny-llm_b200/csrc/build.py
  TL_LIB=.../libtiny_llm_b200_prof.so python tools/s4_timeline.py

Runs the five decode projections of one Qwen3-4B layer back to back inside a CUDA graph (PDL on),
many layers deep, and prints the median globaltimer deltas of CTA 0: entry -> prefetch issued ->
dependency wait over -> activations staged -> last unit consumed -> CTA done, plus the gap between
consecutive launches.
"""
from __future__ import annotations

import ctypes
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]

from extensions_b200 import tiny_llm_ext_b200 as ext  # noqa: E402

SHAPES = [("qkv", 2560, 6144, "rms"), ("o", 4096, 2560, "res"), ("gate_up", 2560, 19456, "rms"), ("down", 9728, 2560, "res")]


def main():
    dev = torch.device("cuda:0")
    ext.set_pdl(True)
    lib = ctypes.CDLL(str(ext.current_library_path()))
    g = torch.Generator(device=dev).manual_seed(0)
    layers = 12
    ws = {}
    for name, N, K, _ in SHAPES:
        ws[name] = []
        for _ in range(layers):
            w = torch.randint(-(2**31), 2**31, (K, N // 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
            s = (torch.randn(K, N // 128, device=dev, generator=g) * 0.01).to(torch.bfloat16)
            ws[name].append((w, s, (-7.5 * s.float()).to(torch.bfloat16)))
    x = torch.randn(1, 2560, device=dev, generator=g).to(torch.bfloat16)
    y = torch.randn(1, 4096, device=dev, generator=g).to(torch.bfloat16)
    nw = torch.ones(2560, device=dev, dtype=torch.bfloat16)

    def step():
        h = x
        for i in range(layers):
            w, s, b = ws["qkv"][i]
            ext.quantized_matmul_fused(s, b, w, h, nw, prologue=ext.PRO_RMSNORM, eps=1e-6)
            w, s, b = ws["o"][i]
            h2 = ext.quantized_matmul_fused(s, b, w, y, residual=h, epilogue=ext.EPI_RESIDUAL)
            w, s, b = ws["gate_up"][i]
            act = ext.quantized_matmul_fused(s, b, w, h2, nw, prologue=ext.PRO_RMSNORM, eps=1e-6, epilogue=ext.EPI_SWIGLU_PAIRS)
            w, s, b = ws["down"][i]
            h = ext.quantized_matmul_fused(s, b, w, act, residual=h2, epilogue=ext.EPI_RESIDUAL)

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (4096 * 8))()
        lib.tl_debug_s4_prof(buf, 4096)  # drop warm-up stamps
        graph.replay()
        torch.cuda.synchronize()
        n = lib.tl_debug_s4_prof(buf, 4096)
    rows = [list(buf[i * 8 : i * 8 + 8]) for i in range(n)]
    rows.sort(key=lambda r: r[0])
    # stamps: 0 entry, 1 prefetch issued, 2 dependency wait over, 3 staged, 4 last unit consumed,
    #         6 block barrier passed, 5 outputs stored
    order = [(0, 1, "prefetch-issue"), (1, 2, "dep-wait"), (2, 3, "stage"), (3, 4, "consume"), (4, 6, "block-barrier"), (6, 5, "reduce+store")]
    per = {s[0]: {k: [] for _, _, k in order} for s in SHAPES}
    for i, r in enumerate(rows):
        shape = SHAPES[i % 4][0]
        for a, b, k in order:
            per[shape][k].append((r[b] - r[a]) / 1e3)
        per[shape].setdefault("total", []).append((r[5] - r[0]) / 1e3)
        if i > 0:
            per[shape].setdefault("gap-from-prev-end", []).append((r[0] - rows[i - 1][5]) / 1e3)
    print(f"{n} launches; medians in us (CTA 0)")
    for shape, d in per.items():
        print(shape, {k: round(statistics.median(v), 2) for k, v in d.items() if v})
    span = (rows[-1][5] - rows[0][0]) / 1e3
    print(f"span of {n} launches: {span:.1f} us -> {span / n:.2f} us per launch")


if __name__ == "__main__":
    main()
