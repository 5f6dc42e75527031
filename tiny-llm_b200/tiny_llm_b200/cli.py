"""Command-line front doors: the equivalents of the reference's ``main.py`` (one prompt,
``/root/reference/main.py:1-190``) and ``batch-main.py`` (continuous batching,
``/root/reference/batch-main.py:1-102``) over a checkpoint DIRECTORY (MLX 4-bit safetensors layout,
see checkpoint.py; there is no hub download here).

    python -m tiny_llm_b200.cli generate --model /path/to/Qwen3-4B-MLX-4bit --prompt "..." [--loader week3]
    python -m tiny_llm_b200.cli batch    --model /path/to/ckpt --prompts-file prompts.txt --batch-size 5
    python -m tiny_llm_b200.cli generate --synthetic tiny-d128 --prompt-ids 5,17,3 --max-new-tokens 8   (no files needed)
"""

from __future__ import annotations

import argparse
import sys

import torch


def _load(args, device):
    from .checkpoint import load_checkpoint, load_tokenizer
    from .synthetic import synthetic_qwen3

    if args.synthetic:
        return synthetic_qwen3(args.synthetic, seed=0, device=device, realistic=args.synthetic.startswith("tiny")), None, "Qwen/Qwen3-synthetic"
    ns = load_checkpoint(args.model, device=device)
    tokenizer = None
    try:
        tokenizer = load_tokenizer(args.model)
    except Exception as exc:  # token-id mode still works without tokenizer files
        print(f"(no tokenizer loaded from {args.model}: {exc})", file=sys.stderr)
    return ns, tokenizer, "Qwen/Qwen3-" + str(args.model)


def _model(args, ns, name):
    from .models import dispatch_model

    if args.loader == "week2":
        return dispatch_model(name, ns, week=2)
    return dispatch_model(name, ns, week=3, enable_paged_attention=not args.disable_paged_attention)


def _prompt_ids(args, tokenizer, prompt: str | None):
    if args.prompt_ids:
        return [int(t) for t in args.prompt_ids.split(",")]
    if tokenizer is None:
        raise SystemExit("a text prompt needs the checkpoint's tokenizer files; use --prompt-ids")
    messages = [{"role": "system", "content": "You are a helpful assistant."}, {"role": "user", "content": prompt}]
    text = tokenizer.apply_chat_template(messages, tokenize=False, add_generation_prompt=True, enable_thinking=args.enable_thinking)
    return tokenizer.encode(text, add_special_tokens=False)


def cmd_generate(args) -> int:
    from .generate import greedy_generate_ids
    from .sampler import make_sampler

    device = torch.device(args.device)
    ns, tokenizer, name = _load(args, device)
    model = _model(args, ns, name)
    ids = _prompt_ids(args, tokenizer, args.prompt)
    sampler = None if args.sampler_temp == 0 else make_sampler(args.sampler_temp, top_p=args.sampler_top_p, top_k=args.sampler_top_k)
    if tokenizer is not None:
        tokenizer.detokenizer.reset()

    def emit(token: int) -> None:
        if tokenizer is None:
            print(token, end=" ", flush=True)
        else:
            tokenizer.detokenizer.add_token(token)
            print(tokenizer.detokenizer.last_segment, end="", flush=True)

    produced = greedy_generate_ids(model, ids, args.max_new_tokens, eos_token_id=getattr(tokenizer, "eos_token_id", None), device=device,
                                   on_token=emit, sampler=sampler)
    print()
    return 0 if produced is not None else 1


def cmd_batch(args) -> int:
    from .batch import batch_generate

    device = torch.device(args.device)
    ns, tokenizer, name = _load(args, device)
    model = _model(args, ns, name)
    if args.prompts_file:
        prompts = [line.strip() for line in open(args.prompts_file) if line.strip()]
    else:
        prompts = [args.prompt or "Give me a short introduction to large language models."]
    if args.prompt_ids or tokenizer is None:
        queue = [[int(t) for t in p.split(",")] for p in (args.prompt_ids.split(";") if args.prompt_ids else prompts)]
    else:
        queue = [tokenizer.apply_chat_template([{"role": "user", "content": p}], tokenize=False, add_generation_prompt=True,
                                               enable_thinking=args.enable_thinking) for p in prompts]
    results = batch_generate(model, tokenizer, queue, max_seq_len=args.max_seq_len, batch_size=args.batch_size, prefill_step=args.prefill_step,
                             verbose=not args.quiet, device=device, max_new_tokens=[args.max_new_tokens] * len(queue) if args.max_new_tokens else None)
    for idx, text in sorted(results):
        print(f"--- request {idx}\n{text}")
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="tiny_llm_b200.cli")
    sub = ap.add_subparsers(dest="command", required=True)
    for name, fn in (("generate", cmd_generate), ("batch", cmd_batch)):
        p = sub.add_parser(name)
        p.set_defaults(fn=fn)
        p.add_argument("--model", default=None, help="checkpoint directory (config.json + *.safetensors [+ tokenizer files])")
        p.add_argument("--synthetic", default=None, help="random weights of a named shape (tiny, tiny-d128, qwen3-0.6b, qwen3-4b) instead of --model")
        p.add_argument("--prompt", default="Give me a short introduction to large language models.")
        p.add_argument("--prompt-ids", default=None, help="comma-separated token ids (';' between requests for `batch`)")
        p.add_argument("--loader", choices=["week2", "week3"], default="week3")
        p.add_argument("--device", default="cuda:0")
        p.add_argument("--disable-paged-attention", action="store_true")
        p.add_argument("--enable-thinking", action="store_true")
        p.add_argument("--max-new-tokens", type=int, default=128)
    sub.choices["generate"].add_argument("--sampler-temp", type=float, default=0.0)
    sub.choices["generate"].add_argument("--sampler-top-p", type=float, default=None)
    sub.choices["generate"].add_argument("--sampler-top-k", type=int, default=None)
    sub.choices["batch"].add_argument("--prompts-file", default=None)
    sub.choices["batch"].add_argument("--batch-size", type=int, default=5)
    sub.choices["batch"].add_argument("--prefill-step", type=int, default=128)
    sub.choices["batch"].add_argument("--max-seq-len", type=int, default=512)
    sub.choices["batch"].add_argument("--quiet", action="store_true")
    args = ap.parse_args(argv)
    if not args.model and not args.synthetic:
        ap.error("--model or --synthetic is required")
    return args.fn(args)


if __name__ == "__main__":
    raise SystemExit(main())
