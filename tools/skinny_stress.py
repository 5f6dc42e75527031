"""Determinism / correctness stress of the swap-AB split-reduction GEMM: every shape is run `reps`
times on the same inputs; all runs must be bit-identical and match the tiled arithmetic restated with
torch ops on the GPU (weights rounded to the activation dtype, fp32 accumulation)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]
from extensions_b200 import tiny_llm_ext_b200 as ext  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for dtype in (torch.bfloat16, torch.float16):
    for M, N, K in [(100, 9728, 2560), (64, 2560, 6144), (128, 4096, 2560), (16, 2560, 19456), (48, 2560, 151936 // 8 * 2)]:
        g = torch.Generator(device=dev).manual_seed(M + N + K)
        w = torch.randint(-(2**31), 2**31, (K, N // 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        s = (torch.randn(K, N // 128, device=dev, generator=g) * (1.0 / (4.717 * N**0.5))).to(dtype)
        b = (-7.5 * s.float()).to(dtype)
        a = torch.randn(M, N, device=dev, generator=g).to(dtype)
        shifts = torch.arange(0, 32, 4, device=dev, dtype=torch.int64)
        codes = ((w.to(torch.int64)[..., None] >> shifts) & 0xF).reshape(K, N).float()
        wd = (codes * s.float().repeat_interleave(128, dim=1) + b.float().repeat_interleave(128, dim=1)).to(dtype).float()
        want = (a.float() @ wd.T)
        first = ext.quantized_matmul(s, b, 128, 4, a, w, True)
        torch.cuda.synchronize()
        err = (first.float() - want).abs().max().item() / (want.abs().max().item() + 1e-9)
        diffs = 0
        for r in range(reps):
            again = ext.quantized_matmul(s, b, 128, 4, a, w, True)
            n = int((again != first).sum())
            if n:
                diffs += 1
                idx = (again != first).nonzero()[:4].tolist()
                print(f"  run {r}: {n} elements differ, e.g. {idx}", flush=True)
        bad += diffs + (err > 5e-3)
        print(f"{str(dtype)[6:]:9s} M={M:3d} N={N:5d} K={K:6d}: rel err vs torch {err:.2e}, {diffs}/{reps} runs differ", flush=True)
print("STRESS", "FAILED" if bad else "OK")
