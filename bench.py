"""Headline benchmark: Qwen3-4B W4A16 decode tokens/s on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config[1] of BASELINE.json, the one the metric is quoted on):
Qwen3-4B-shaped random W4A16 weights, ONE request per GPU, 128-token prompt
(chunk-prefilled through the paged path), then greedy decode with paged-KV GQA
attention + dequant matvec.  A "step" is one decode step (one token per GPU).
At N > 1 every rank serves its own request (request i -> rank i mod N, weak
scaling); weights are drawn on rank 0 and broadcast once over NCCL; there is no
data-path collective.

Numbers on the JSON line:
  value     decode tok/s, whole job, device-resident: K replays of the captured
            decode step with token feedback on the device (CUDA events, max over ranks)
  e2e       same metric through the public model call, per step: pinned-host token
            -> device copy, model(...), device -> host read of the sampled token
  roofline  the W4A16 weight-streaming kernel: the 145 projection launches of one token
            exactly as the decode graph issues them (2.137 GB of packed weights, > L2),
            replayed from a CUDA graph, CUDA-event timed; achieved = algorithmic
            bytes / time, traffic = ncu DRAM bytes of the same launches
  cpu_baseline  the reference's CPU path (oracle.model: dense bf16 weights, readable
            operators) on the host cores, bounded sample
"""

from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]

import torch  # noqa: E402

METRIC = "Qwen3-4B W4A16 decode tok/s"
UNIT = "tok/s"
MODEL = "qwen3-4b"
PROMPT_LEN = 128
PAGE_SIZE = 128


def measured_peaks() -> dict:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        data = json.loads(path.read_text())
        return {"hbm_gbs": float(data["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def synthetic_prompt(seed: int, length: int, vocab: int) -> list[int]:
    """Token ids as benches/bench.py:190-225 draws them: uniform in [256, V-1]."""
    rng = random.Random(seed)
    return [rng.randint(256, vocab - 1) for _ in range(length)]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cuda_time_ms(fn, stream=None) -> float:
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    fn()
    end.record(stream)
    end.synchronize()
    return start.elapsed_time(end)


# --------------------------------------------------------------------------- ours
def run_ours(args) -> None:
    from extensions_b200 import tiny_llm_ext_b200 as ext
    from tiny_llm_b200 import Qwen3ModelWeek3
    from tiny_llm_b200.batch import greedy_tokens
    from tiny_llm_b200.parallel import barrier, init_distributed, max_over_ranks, replicated_model
    from tiny_llm_b200.synthetic import weight_stream_bytes

    rank, world, device = init_distributed("cuda")
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    pdl = os.environ.get("TL_PDL", "1") != "0"
    ext.set_pdl(pdl)
    t_load = time.perf_counter()
    model_ns, broadcast_bytes = replicated_model(MODEL, seed=0, rank=rank, device=device)
    model = Qwen3ModelWeek3(model_ns, page_size=PAGE_SIZE)
    load_s = time.perf_counter() - t_load
    margs = model_ns.args
    steps, warmup = args.steps, max(args.warmup, 3)
    max_seq = PROMPT_LEN + 2 * (steps + warmup) + 64
    model.decode_graph_max_seq_len = ((max_seq + PAGE_SIZE - 1) // PAGE_SIZE) * PAGE_SIZE
    engine = model.decode_engine(1)

    prompt = synthetic_prompt(1000 + rank, PROMPT_LEN, margs.vocab_size)
    cache = model.create_kv_cache()
    t0 = time.perf_counter()
    first = model(torch.tensor([prompt], dtype=torch.int32, device=device), 0, cache, logits_to_keep=1)
    token = int(greedy_tokens(first[:, -1, :])[0])
    torch.cuda.synchronize()
    prefill_s = time.perf_counter() - t0

    # ---- value: device-resident decode (graph replays, token feedback on device)
    offset = PROMPT_LEN
    engine.decode_on_device([token], [offset], cache, min(warmup, engine.log_capacity))
    torch.cuda.synchronize()
    offset += warmup
    token = int(engine.next_tokens[0])
    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
    launches0 = ext.launch_count()
    replays0 = engine.graph_replays
    barrier(device)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    done = 0
    while done < steps:  # one call unless --steps exceeds the engine's on-device token log (4096)
        take = min(steps - done, engine.log_capacity)
        out_tokens = engine.decode_on_device([token], [offset + done], cache, take)
        done += take
        if done < steps:
            token = int(out_tokens[-1, 0])
    end.record()
    torch.cuda.synchronize()
    barrier(device)
    ms = max_over_ranks(start.elapsed_time(end), device)
    clocks = sampler.stop() if rank == 0 else None
    # graph replays do not pass through the C ABI; the launches recorded when the step was captured
    # are what each replay executes (plus whatever went through the ABI directly in the region)
    gpu_launches = engine.kernels_per_step * (engine.graph_replays - replays0) + (ext.launch_count() - launches0)
    offset += steps
    token = int(out_tokens[-1, 0])
    value = world * steps / (ms / 1e3)

    # ---- e2e: public API per step, pinned host token in, sampled token out
    e2e_steps = min(steps, 64)
    pinned_in = torch.empty(1, 1, dtype=torch.int32, pin_memory=True)
    pinned_out = torch.empty(1, dtype=torch.int32, pin_memory=True)
    for phase, count in (("warm", 3), ("timed", e2e_steps)):
        barrier(device)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(count):
            pinned_in[0, 0] = token
            tok_dev = pinned_in.to(device, non_blocking=True)
            logits = model(tok_dev, offset, cache, logits_to_keep=1)  # the call a user makes
            pinned_out.copy_(greedy_tokens(logits[:, -1, :]), non_blocking=True)
            torch.cuda.synchronize()
            token = int(pinned_out[0])
            offset += 1
        e2e_s = time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_s, device)
    e2e_value = world * e2e_steps / e2e_s

    # ---- roofline of the weight-streaming kernel (rank 0 only, N == 1 semantics)
    roofline = None
    if rank == 0:
        roofline = matvec_roofline(model, engine, ext, device)

    for c in cache:
        c.release()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(sample_steps=3)

    if rank == 0:
        stream_bytes = weight_stream_bytes(margs)
        kv_bytes = 147456 * (PROMPT_LEN + warmup + steps // 2)
        peak = measured_peaks()
        line = {
            "metric": METRIC,
            "value": round(value, 2),
            "unit": UNIT,
            "n_gpus": world,
            "steps": steps,
            "warmup": warmup,
            "ms_per_step": round(ms / steps, 4),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "bf16 activations, 4-bit weights (W4A16), fp32 accumulation",
            "data": "synthetic (random Qwen3-4B-shaped W4A16 weights, random prompt)",
            "config": {
                "workload": "Qwen3-4B W4A16 single-request decode, batch=1 per GPU, paged-KV GQA + dequant matvec",
                "prompt_len": PROMPT_LEN, "page_size": PAGE_SIZE, "requests_per_gpu": 1, "parallelism": f"dp{world}",
                "l2_policy": "inputs larger than L2: each step streams 2.14 GB of packed weights (L2 = 126 MB)",
                "decode_graph": "cuda-graph replay, fused=%s, pdl=%s" % (engine.fused, int(pdl)),
            },
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": 4 + engine._meta_len * 4,
                    "d2h_bytes_per_step": 4, "steps": e2e_steps},
            "gpu_launches": int(gpu_launches),
            "kernels_per_step": int(engine.kernels_per_step),
            "clocks": clocks,
            "roofline": roofline,
            "token_roofline": {
                "bytes_per_token": stream_bytes + kv_bytes,
                "achieved_gbs": round((stream_bytes + kv_bytes) / (ms / steps / 1e3) / 1e9, 1),
                "frac": round((stream_bytes + kv_bytes) / (ms / steps / 1e3) / 1e9 / peak["hbm_gbs"], 4),
                "peak_gbs": peak["hbm_gbs"], "peak_source": peak["source"],
            },
            "cpu_baseline": cpu_baseline,
            "setup": {"weights_load_s": round(load_s, 2), "broadcast_bytes": broadcast_bytes, "prefill_128_s": round(prefill_s, 3)},
        }
        print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of the w4a16_stream5_kernel launches of one decode
# token, averaged over its 145 launches (profiles/r01_launches_decode_final.csv; algorithmic: 14.76 MB).
NCU_TRAFFIC_PER_LAUNCH = 14826718


def matvec_roofline(model, engine, ext, device) -> dict:
    """The projection launches of one decode token exactly as the engine issues them (per layer:
    rms_norm+q|k|v, o+residual, rms_norm+gate|up+swiglu, down+residual; then rms_norm+head), M = 1,
    from one captured graph: 2.137 GB of distinct packed weights per replay, so every launch
    streams from HBM (L2 = 126 MB)."""
    from tiny_llm_b200.synthetic import weight_stream_bytes

    H = model.hidden_size
    bf = torch.bfloat16
    x = torch.randn(1, H, device=device).to(bf)
    y = torch.randn(1, engine.Hq * engine.D, device=device).to(bf)
    res = torch.randn(1, H, device=device).to(bf)
    norm_w = torch.ones(H, device=device, dtype=bf)
    inter = model.layers_inner[0].mlp.hidden_dim
    act = torch.randn(1, inter, device=device).to(bf)
    head = model.w_lm_head if model.w_lm_head is not None else model.embedding.weight
    launches = 4 * len(model.layers_inner) + 1

    def body():
        for block, pk in zip(model.layers_inner, engine._packed):
            wo, wd = block.self_attn.wo, block.mlp.w_down
            ext.quantized_matmul_fused(pk.qkv.scales, pk.qkv.biases, pk.qkv.weight, x, norm_w, prologue=ext.PRO_RMSNORM, eps=1e-6)
            ext.quantized_matmul_fused(wo.scales, wo.biases, wo.weight, y, residual=res, epilogue=ext.EPI_RESIDUAL)
            ext.quantized_matmul_fused(pk.gate_up.scales, pk.gate_up.biases, pk.gate_up.weight, x, norm_w, prologue=ext.PRO_RMSNORM,
                                       eps=1e-6, epilogue=ext.EPI_SWIGLU_PAIRS)
            ext.quantized_matmul_fused(wd.scales, wd.biases, wd.weight, act, residual=res, epilogue=ext.EPI_RESIDUAL)
        ext.quantized_matmul_fused(head.scales, head.biases, head.weight, x, norm_w, prologue=ext.PRO_RMSNORM, eps=1e-6)

    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        body()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            body()
        for _ in range(3):
            graph.replay()
        stream.synchronize()
        times = [cuda_time_ms(graph.replay, stream) for _ in range(10)]
    ms = statistics.median(times)
    # algorithmic bytes: every packed weight, scale and bias once (0.53125 B/weight) + activations in/out
    margs = model.mlx_model.args
    io = 2 * len(model.layers_inner) * (H + (engine.Hq + 2 * engine.Hkv) * engine.D + engine.Hq * engine.D + 2 * H + H + inter + inter + 2 * H)
    algorithmic = weight_stream_bytes(margs) + io + 2 * (H + margs.vocab_size)
    peak = measured_peaks()
    achieved = algorithmic / (ms / 1e3) / 1e9
    return {
        "kernel": "w4a16_stream5_kernel<bf16, M=1> (W4A16 dequant matvec with fused rms_norm / residual / SwiGLU epilogue)",
        "bound": "hbm", "achieved": round(achieved, 1), "peak": peak["hbm_gbs"], "peak_source": peak["source"], "unit": "GB/s",
        "frac": round(achieved / peak["hbm_gbs"], 4), "traffic": NCU_TRAFFIC_PER_LAUNCH, "launches": launches,
        "avg_launch_us": round(ms * 1e3 / launches, 3), "algorithmic_bytes_per_launch": round(algorithmic / launches),
        "timing": f"cuda events on the launching stream around a graph replay of one token's {launches} projection launches, median of 10",
    }


# ------------------------------------------------------------- CPU reference arm
CPU_SAMPLE_LAYERS = 4


def run_cpu_baseline(sample_steps: int, warmup_steps: int = 1) -> dict:
    """tiny_llm_ref's CPU-capable path (oracle.model) on the host cores, bounded: the same synthetic
    Qwen3-4B shapes with CPU_SAMPLE_LAYERS of the 36 transformer blocks (+ embedding and tied head),
    an 8-token prompt and a few decode steps; the per-token time is scaled to the full depth by
    weight bytes (a decode step on the CPU is one pass over every dense weight)."""
    from oracle.model import ReferenceCpuModel, greedy_decode
    from tiny_llm_b200.synthetic import CONFIGS, synthetic_qwen3

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = CONFIGS[MODEL]
    ns = synthetic_qwen3(MODEL, seed=0, device="cpu", num_hidden_layers=CPU_SAMPLE_LAYERS)
    model = ReferenceCpuModel(ns)
    del ns
    prompt = synthetic_prompt(1000, 8, model.args.vocab_size)
    timings: dict = {}
    greedy_decode(model, prompt, 1 + warmup_steps + sample_steps, timings=timings)
    per_step = timings["decode_s"][warmup_steps:]
    H, inter = full["hidden_size"], full["intermediate_size"]
    q_w, kv_w = full["num_attention_heads"] * full["head_dim"], full["num_key_value_heads"] * full["head_dim"]
    layer_w = H * (q_w + 2 * kv_w) + q_w * H + 3 * H * inter
    head_w = full["vocab_size"] * H
    scale = (full["num_hidden_layers"] * layer_w + head_w) / (CPU_SAMPLE_LAYERS * layer_w + head_w)
    sample_s = statistics.median(per_step)
    value = 1.0 / (sample_s * scale)
    return {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"oracle.model (reference CPU path: dense bf16 weights, readable ops), {CPU_SAMPLE_LAYERS} of "
                       f"{full['num_hidden_layers']} Qwen3-4B blocks + tied head, 8-token prompt, median of {len(per_step)} decode steps "
                       f"({1e3 * sample_s:.0f} ms each), scaled x{scale:.2f} by weight bytes to the full depth"),
            "ms_per_step": round(1e3 * sample_s * scale, 1)}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, max(args.warmup, 1)
    sample = min(steps, 4)  # bounded: ~5 s per CPU decode step of the 4B model on the box's host cores
    base = run_cpu_baseline(sample_steps=sample, warmup_steps=1)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": sample,
        "warmup": 1,
        "ms_per_step": base["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16 weights/activations (dequantised W4), fp32 attention",
        "data": "synthetic (same random Qwen3-4B-shaped weights as the GPU arm)",
        "config": {"workload": "Qwen3-4B single-request decode, batch=1, reference CPU path on host cores", "prompt_len": 8,
                   "note": "MLX cannot be installed here and the reference's native ops are GPU-only; this is the oracle port of tiny_llm_ref's CPU-capable path"},
        "cpu_baseline": {"kind": base["kind"], "cores": base["cores"], "sample": base["sample"], "value": base["value"], "unit": UNIT},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
