"""Front-door pieces on the host: the MLX-4bit safetensors loader / writer, the
sampler and the CLIs, driven through the CPU stand-in of the extension (no GPU)."""

import json

import pytest
import torch

from tiny_llm_b200 import load_checkpoint, make_sampler, save_checkpoint
from tiny_llm_b200.cli import main as cli_main
from tiny_llm_b200.synthetic import named_tensors, synthetic_qwen3


def _bits(t):
    return t.view(torch.int32) if t.dtype == torch.uint32 else t


def test_checkpoint_round_trip_is_bit_exact_and_keeps_the_duck_type(tmp_path):
    ns = synthetic_qwen3("tiny", seed=4, tie_word_embeddings=False)
    save_checkpoint(ns, tmp_path)
    config = json.loads((tmp_path / "config.json").read_text())
    assert config["quantization"] == {"group_size": 128, "bits": 4} and config["num_hidden_layers"] == 2
    back = load_checkpoint(tmp_path)
    assert vars(back.args) == vars(ns.args)
    a, b = dict(named_tensors(ns)), dict(named_tensors(back))
    assert a.keys() == b.keys()
    assert all(torch.equal(_bits(a[k]), _bits(b[k])) and a[k].dtype == b[k].dtype for k in a)
    layer = back.model.layers[1].self_attn.q_proj
    assert layer.weight.dtype == torch.uint32 and layer.group_size == 128 and layer.bits == 4
    assert hasattr(back, "lm_head")


def test_loader_reports_missing_tensors_and_wrong_dtypes(tmp_path):
    ns = synthetic_qwen3("tiny", seed=4)
    save_checkpoint(ns, tmp_path)
    from safetensors.torch import load_file, save_file

    tensors = {k: v.clone() for k, v in load_file(str(tmp_path / "model.safetensors")).items()}  # the file is mmapped: copy before rewriting it
    broken = dict(tensors)
    del broken["model.layers.0.mlp.up_proj.scales"]
    save_file(broken, str(tmp_path / "model.safetensors"))
    with pytest.raises(KeyError, match="up_proj.scales"):
        load_checkpoint(tmp_path)
    dense = dict(tensors)
    dense["model.layers.0.mlp.up_proj.weight"] = torch.zeros(4, 4, dtype=torch.bfloat16)
    save_file(dense, str(tmp_path / "model.safetensors"))
    with pytest.raises(ValueError, match="packed uint32"):
        load_checkpoint(tmp_path)


def reference_sampler_keep_set(logprobs, top_p, top_k):
    """sampler.py:9-21 restated with explicit loops (one row): the ids that stay finite."""
    order = sorted(range(len(logprobs)), key=lambda i: -logprobs[i])
    keep = set(order[:top_k]) if top_k else set(order)
    if top_p:
        mass, kept = 0.0, set()
        for i in order:
            if i in keep and mass < top_p:
                kept.add(i)
            if i in keep:
                mass += float(torch.exp(torch.tensor(logprobs[i])))
        keep = kept
    return keep


@pytest.mark.parametrize("top_p,top_k", [(None, None), (None, 3), (0.6, None), (0.35, 5), (0.999, 2)])
def test_sampler_draws_only_from_the_reference_keep_set(top_p, top_k):
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(1, 40, generator=g) * 2
    logprobs = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    keep = reference_sampler_keep_set(logprobs[0].tolist(), top_p, top_k)
    sample = make_sampler(0.8, top_p=top_p, top_k=top_k, generator=torch.Generator().manual_seed(6))
    drawn = {int(sample(logprobs)[0]) for _ in range(300)}
    assert drawn <= keep
    if len(keep) <= 5:
        assert drawn == keep, "300 draws at temperature 0.8 reach every kept token of a small set"
    assert int(make_sampler(0, top_p, top_k)(logprobs)[0]) == int(torch.argmax(logprobs))


def test_sampler_is_reproducible_and_batched():
    lp = torch.log_softmax(torch.randn(4, 100, generator=torch.Generator().manual_seed(1)), dim=-1)
    a = make_sampler(1.0, 0.9, 20, generator=torch.Generator().manual_seed(2))(lp)
    b = make_sampler(1.0, 0.9, 20, generator=torch.Generator().manual_seed(2))(lp)
    assert a.shape == (4,) and torch.equal(a, b)


def test_cli_generate_and_batch_on_a_synthetic_model(cpu_ext, capsys):
    assert cli_main(["generate", "--synthetic", "tiny-d128", "--prompt-ids", "5,17,3,250", "--max-new-tokens", "4", "--device", "cpu"]) == 0
    out = capsys.readouterr().out.split()
    assert len(out) == 4 and all(t.isdigit() for t in out)
    assert cli_main(["generate", "--synthetic", "tiny-d128", "--prompt-ids", "5,17,3,250", "--max-new-tokens", "4", "--device", "cpu",
                     "--sampler-temp", "0.7", "--sampler-top-k", "5"]) == 0
    capsys.readouterr()
    assert cli_main(["batch", "--synthetic", "tiny-d128", "--prompt-ids", "5,17,3;9,2,4,6,8", "--max-new-tokens", "3", "--device", "cpu",
                     "--batch-size", "2", "--max-seq-len", "64", "--prefill-step", "4", "--quiet"]) == 0
    text = capsys.readouterr().out
    assert "--- request 0" in text and "--- request 1" in text
