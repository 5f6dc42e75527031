"""Data-parallel request sharding + one-time weight broadcast, world_size 2 on
CPU with the gloo backend (the N>1 path of bench.py / the serving launcher)."""

import os
import subprocess
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent


def test_two_rank_data_parallel_serving(tmp_path):
    world, port = 2, 29531 + (os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, str(HERE / "dp_worker.py"), str(r), str(world), str(port), str(tmp_path)]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    assert r0["digest"] == r1["digest"], "weights differ after broadcast"
    assert r0["nbytes"] == r1["nbytes"] > 0
    assert (r0["tokens"], r1["tokens"]) == (8, 6)  # 4 and 3 requests, 2 tokens each
    assert r0["total"] == r1["total"] == 14.0
    assert r0["slowest"] == r1["slowest"] == 2.0
    assert r0["free"] and r1["free"]
