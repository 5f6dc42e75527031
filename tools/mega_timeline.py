"""In-kernel timeline of the whole-step decode kernel (CTA 0): where one token's time goes.

  python tools/mega_timeline.py [--layers 36] [--context 128] [--out gpurun_out/mega_timeline.json]
"""
from __future__ import annotations

import argparse
import collections
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]

from tiny_llm_b200 import Qwen3ModelWeek3  # noqa: E402
from tiny_llm_b200.engine import DecodeEngine  # noqa: E402
from tiny_llm_b200.synthetic import synthetic_qwen3  # noqa: E402

KIND = {1: "stage", 2: "items", 3: "cta-sync", 4: "reduce+store", 5: "grid-sync", 6: "attention", 7: "grid-sync(attn)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=36)
    ap.add_argument("--context", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "mega_timeline.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ns = synthetic_qwen3("qwen3-4b", seed=0, device=dev, num_hidden_layers=args.layers)
    model = Qwen3ModelWeek3(ns, page_size=128)
    engine = DecodeEngine(model, args.batch, args.context + 256, dev, persistent=True)
    engine.reserve_pools()
    cache = model.create_kv_cache()
    for layer_cache in cache:
        for _ in range(args.context):
            layer_cache.append_token_slot()
    assert args.batch == 1
    cap = 4096
    prof = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
    offset = args.context
    for _ in range(3):
        engine.step([1000], [offset], cache)
        offset += 1
    torch.cuda.synchronize()
    engine._mega.args.prof = prof.data_ptr()
    engine._mega.args.prof_capacity = cap
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    engine.step([1000], [offset], cache)
    end.record()
    torch.cuda.synchronize()
    total_ms = start.elapsed_time(end)
    stamps = prof.cpu().reshape(-1, 2).tolist()
    stamps = [s for s in stamps if s[0] != 0]
    mhz = 1965.0
    t0 = stamps[0][1]
    agg = collections.defaultdict(float)
    rows = []
    prev = t0
    for tag, clk in stamps:
        dt = (clk - prev) / mhz
        prev = clk
        if 50000 <= tag < 60000:
            name = "attn:" + {1: "loads-issued", 2: "page-ids+sync", 3: "kv-copies-issued", 4: "q-path", 5: "wait+sync", 6: "PV+sync", 7: "absorb-round", 8: "scores(mma)+sync", 9: "softmax-stats+sync"}.get(tag - 50000, str(tag))
        elif tag >= 100 and tag < 90000:
            sp, kind = (tag - 100) // 10, (tag - 100) % 10
            name = ["qkv", "o", "gate_up", "down"][sp % 4] + ":" + KIND.get(kind, str(kind))
        else:
            name = {1: "start", 2: "ring-fill+embed", 3: "grid-sync(embed)", 90001: "head", 99999: "argmax+sync"}.get(tag, str(tag))
        agg[name] += dt
        if len(rows) < 60:
            rows.append((name, round(dt, 2)))
    print("event-timed step: %.3f ms; stamped span: %.1f us" % (total_ms, (stamps[-1][1] - t0) / mhz))
    print("first events (us):", rows)
    print("totals per kind (us):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"  {k:28s} {v:9.1f}")
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps({"step_ms": total_ms, "totals_us": agg, "first": rows}, indent=1))


if __name__ == "__main__":
    main()
