#!/bin/bash
# round-2 call 21: cost-model split count for the tcgen05 decode attention; one-CTA-per-row q/k norm + RoPE + append
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c21_pytest.log; tail -4 gpurun_out/c21_pytest.log | cut -c1-220
timeout 200 python tools/kbench.py --attention-only --out gpurun_out/c21_kbench_att.json 2>&1 | tail -9
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
env timeout 200 python tools/decode_ab.py --tag b32_ctx4096 --batch 32 --context 4096 --steps 32 2>&1 | tail -1
for i in 1 2; do
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c21_bench_serve_$i.json 2> gpurun_out/c21_bench_serve_$i.err; echo "bench serve rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c21_bench_serve_$i.json'))['serving'];print(d['output_tok_s'], 'decode p50', d['decode_step_ms_p50'], 'chunk p50', d.get('prefill_chunk_ms_p50'), 'prefill s', d['time_in_prefill_s'], 'decode s', d['time_in_decode_s'])"
done
timeout 900 python bench.py --workload serve8k --no-cpu-baseline > gpurun_out/c21_bench_serve8k.json 2> gpurun_out/c21_bench_serve8k.err; echo "bench serve8k rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c21_bench_serve8k.json'));print(d['value'], d['serving'])" | cut -c1-600
