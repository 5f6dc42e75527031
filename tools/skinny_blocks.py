"""Per-block cycle timeline of CTA 0 of the swap-AB GEMM (debug build with -DTL_TRACE=1):
  TL_LIB=.../libtiny_llm_b200_trace.so python tools/skinny_blocks.py [M N K]
Columns (SM cycles since the first stamp): MMA thread: A tile ready / B tile ready / MMAs issued; dequantiser group
owning the block: box ready / math done / stage free / handed over."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]
from extensions_b200 import tiny_llm_ext_b200 as ext  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 2560, 151936)
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
ext.quantized_matmul(s, b, 128, 4, a, w, True)
torch.cuda.synchronize()
lib.tl_debug_trace(None, None, ctypes.c_uint(0))
n = min(int(count[0]), cap)
ev = events[: 2 * n].cpu().reshape(-1, 2).tolist()
tab = {}
for tag, t in ev:
    if tag >= 10000:
        role, rest = divmod(tag - 10000, 1000)
        if not (role == 0 and rest % 4 == 3):
            tab[(role, rest // 4, rest % 4)] = t
if not tab:
    sys.exit("no per-block stamps (not the trace build?)")
t0 = min(tab.values())
blocks = sorted({k[1] for k in tab})
print(f"M={M} N={N} K={K}: CTA 0, {len(blocks)} blocks; cycles since first stamp")
print("blk |  A-ready  B-ready   issued | grp  box-rdy math-done stage-free   handed | issued-prev")
prev = None
for i in blocks:
    m = [tab.get((0, i, k), 0) - t0 for k in range(3)]
    role = 1 + i % 2
    d = [tab.get((role, i, k), 0) - t0 for k in range(4)]
    print(f"{i:3d} | {m[0]:8d} {m[1]:8d} {m[2]:8d} |  {i % 2}  {d[0]:8d} {d[1]:9d} {d[2]:10d} {d[3]:8d} | {m[2] - prev if prev is not None else 0:6d}")
    prev = m[2]
