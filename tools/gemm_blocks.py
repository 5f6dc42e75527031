"""Per-block cycle timeline of CTA (0,0) of the tiled W4A16 GEMM (debug build with -DTL_TRACE=1):
  TL_LIB=.../libtiny_llm_b200_trace.so python tools/gemm_blocks.py [M N K]        (default 4096 x 2560 -> 19456)
Columns (SM cycles since the first stamp): MMA warp: activation tile ready / weight tile ready / MMAs issued; first
dequantiser warp: loop top / math done / stage free / handed over."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]
from extensions_b200 import tiny_llm_ext_b200 as ext  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 2560, 19456)
dev = torch.device("cuda:0")
lib = ctypes.CDLL(str(ext.current_library_path()))
g = torch.Generator(device=dev).manual_seed(0)
w = torch.randint(-(2**31), 2**31, (K, N // 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
s = (torch.randn(K, N // 128, device=dev, generator=g) * 0.01).to(torch.bfloat16)
b = (-7.5 * s.float()).to(torch.bfloat16)
a = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
for _ in range(3):
    ext.quantized_matmul(s, b, 128, 4, a, w, True)
torch.cuda.synchronize()
cap = 8192
events = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
count = torch.zeros(1, dtype=torch.int32, device=dev)
lib.tl_debug_trace(ctypes.c_void_p(events.data_ptr()), ctypes.c_void_p(count.data_ptr()), ctypes.c_uint(cap))
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
ext.quantized_matmul(s, b, 128, 4, a, w, True)
end.record()
torch.cuda.synchronize()
lib.tl_debug_trace(None, None, ctypes.c_uint(0))
n = min(int(count[0]), cap)
tab = {}
for tag, t in events[: 2 * n].cpu().reshape(-1, 2).tolist():
    if 20000 <= tag < 30000:
        role, rest = divmod(tag - 20000, 1000)
        tab[(role, rest // 4, rest % 4)] = t
if not tab:
    sys.exit("no per-block stamps (not the trace build, or not the tiled GEMM path)")
t0 = min(tab.values())
us = start.elapsed_time(end) * 1e3
print(f"M={M} N={N} K={K}: {us:.1f} us ({2.0 * M * N * K / us / 1e6:.0f} TF/s under the trace build); CTA (0,0); cycles since first stamp")
print("blk |  A-ready  W-ready   issued (d) | deq: top  math-done stage-free   handed")
prev = None
for i in sorted({k[1] for k in tab}):
    m = [tab.get((0, i, k), 0) - t0 for k in range(3)]
    d = [tab.get((1, i, k), 0) - t0 for k in range(4)]
    print(f"{i:3d} | {m[0]:8d} {m[1]:8d} {m[2]:8d} ({m[2] - prev if prev is not None else 0:5d}) | {d[0]:8d} {d[1]:9d} {d[2]:10d} {d[3]:8d}")
    prev = m[2]
