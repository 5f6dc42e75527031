"""Rank body of tests/test_data_parallel_gloo.py (run as a script, one process per rank)."""

import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path[:0] = [str(ROOT), str(ROOT / "tiny-llm_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from extensions_b200 import tiny_llm_ext_b200
    from oracle import ext_cpu
    from tiny_llm_b200 import ContinuousBatcher, Qwen3ModelWeek3
    from tiny_llm_b200.parallel import init_distributed, max_over_ranks, replicated_model, shard, shard_indices, sum_over_ranks
    from tiny_llm_b200.synthetic import named_tensors

    ext_cpu.install(tiny_llm_ext_b200)  # CPU stand-in for the kernels (test only)
    r, w, device = init_distributed("cpu")
    assert (r, w) == (rank, world)
    model_ns, nbytes = replicated_model("tiny", seed=5, rank=r, device=device, max_position_embeddings=128)
    # every rank must hold identical bytes after the broadcast
    digest = sum(int(t.view(torch.int32).to(torch.int64).sum()) if t.dtype == torch.uint32 else float(t.float().sum()) for _, t in named_tensors(model_ns))
    all_requests = [[3 + i] * (4 + i) for i in range(7)]
    mine = shard(all_requests, r, w)
    assert [len(p) - 4 for p in mine] == shard_indices(7, r, w)
    model = Qwen3ModelWeek3(model_ns, page_size=4)
    results = ContinuousBatcher(model, None, mine, max_seq_len=64, batch_size=2, prefill_step=4, verbose=False, max_new_tokens=[2] * len(mine)).run()
    tokens = sum(len(text.split()) for _, text in results)
    total = sum_over_ranks(float(tokens), device)
    slowest = max_over_ranks(float(rank + 1), device)
    torch.save(dict(rank=r, digest=digest, nbytes=nbytes, tokens=tokens, total=total, slowest=slowest,
                    texts=[text for _, text in sorted(results)], free=all(p.used_page_ids == set() for p in model.page_pools)),
               os.path.join(out_dir, f"rank{r}.pt"))
    dist.destroy_process_group()



if __name__ == "__main__":
    _worker(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
