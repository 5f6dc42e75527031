"""Cross-kernel timeline of the CUDA-graph decode step (debug build with -DTL_TRACE=1):

  TL_DEFINES="-DTL_TRACE=1" TL_LIB_SUFFIX=_trace python tiny-llm_b200/csrc/build.py
  TL_LIB=.../libtiny_llm_b200_trace.so python tools/graph_timeline.py [--layers 36] [--context 128]

Thread 0 of CTA 0 of every projection / attention launch appends (tag, %globaltimer) events to one
device buffer; the script replays the engine's self-advancing decode graph, sorts the events and
prints, per kernel kind, the median duration of every phase and the gaps between launches.
Tags: 10 entry, 11 weights requested, 12 dependency wait over, 13 activations staged, 14 last unit
consumed, 16 block barrier, 15 stored (projection); 20 entry, 21 K/V requested, 22 dependency wait
over, 23 rows + q ready, 29 done (attention).
"""
from __future__ import annotations

import argparse
import collections
import ctypes
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]

from extensions_b200 import tiny_llm_ext_b200 as ext  # noqa: E402
from tiny_llm_b200 import Qwen3ModelWeek3  # noqa: E402
from tiny_llm_b200.engine import DecodeEngine  # noqa: E402
from tiny_llm_b200.synthetic import synthetic_qwen3  # noqa: E402

PHASES = {(10, 11): "request-weights", (11, 12): "dep-wait", (12, 13): "stage", (13, 14): "consume", (14, 16): "block-barrier",
          (16, 15): "reduce+store", (20, 21): "request-kv", (21, 22): "dep-wait", (22, 23): "q-path+rows-landed", (23, 29): "scores+softmax+PV"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=36)
    ap.add_argument("--context", type=int, default=128)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = ctypes.CDLL(str(ext.current_library_path()))
    ns = synthetic_qwen3("qwen3-4b", seed=0, device=dev, num_hidden_layers=args.layers)
    model = Qwen3ModelWeek3(ns, page_size=128)
    engine = DecodeEngine(model, 1, args.context + 256, dev)
    engine.reserve_pools()
    cache = model.create_kv_cache()
    for layer_cache in cache:
        for _ in range(args.context):
            layer_cache.append_token_slot()
    engine.decode_on_device([1000], [args.context], cache, 4)  # warm-up replays
    torch.cuda.synchronize()
    cap = 1 << 16
    events = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.tl_debug_trace(ctypes.c_void_p(events.data_ptr()), ctypes.c_void_p(count.data_ptr()), ctypes.c_uint(cap))
    engine.decode_on_device([1000], [args.context + 4], cache, 1)
    torch.cuda.synchronize()
    n = int(count[0])
    lib.tl_debug_trace(None, None, ctypes.c_uint(0))
    ev = events[: 2 * n].cpu().reshape(-1, 2).tolist()
    ev.sort(key=lambda e: e[1])
    # group events into launches: a launch starts at tag 10 or 20; events of overlapping launches are
    # told apart by their tag families (1x projection, 2x attention), and projections never overlap
    # each other past the entry stamp, so "latest open launch of the family" is exact enough
    launches, open_p, open_a = [], [], []
    for tag, t in ev:
        fam = open_p if tag < 20 else open_a
        if tag in (10, 20):
            rec = {"kind": "proj" if tag == 10 else "attn", "t": {tag: t}}
            launches.append(rec)
            fam.append(rec)
        else:
            for rec in fam:
                if tag not in rec["t"]:
                    rec["t"][tag] = t
                    break
            if tag in (15, 29):
                fam[:] = [r for r in fam if tag not in r["t"]]
    per_layer = 5
    names = ["rms+qkv", "attention", "o+res", "rms+gate|up+swiglu", "down+res"]
    stats = collections.defaultdict(lambda: collections.defaultdict(list))
    order = [rec for rec in launches]
    for i, rec in enumerate(order[: args.layers * per_layer]):
        name = names[i % per_layer]
        for (a, b), label in PHASES.items():
            if a in rec["t"] and b in rec["t"]:
                stats[name][label].append((rec["t"][b] - rec["t"][a]) / 1e3)
        first, last = min(rec["t"].values()), max(rec["t"].values())
        stats[name]["total"].append((last - first) / 1e3)
        if i > 0:
            prev = order[i - 1]
            stats[name]["entry-after-prev-end"].append((first - max(prev["t"].values())) / 1e3)
            stats[name]["entry-after-prev-entry"].append((first - min(prev["t"].values())) / 1e3)
    print(f"{n} events, {len(launches)} launches; medians in us (CTA 0 of each launch)")
    for name in names:
        print(f"  {name:20s}", {k: round(statistics.median(v), 2) for k, v in stats[name].items()})
    span = (ev[-1][1] - ev[0][1]) / 1e3
    print(f"span first->last event: {span:.1f} us")


if __name__ == "__main__":
    main()
