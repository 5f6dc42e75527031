#!/bin/bash
# round-2 call 28: q|k|v projection planes feeding q/k norm + RoPE + append directly
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c28_pytest.log; tail -8 gpurun_out/c28_pytest.log | cut -c1-220
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
timeout 600 python bench.py --workload serve --no-extra --no-cpu-baseline > gpurun_out/c28_bench_serve.json 2> gpurun_out/c28_bench_serve.err; echo "bench serve rc=$?"; tail -c 200 gpurun_out/c28_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c28_bench_serve.json'));s=d['serving'];print(d['value'], 'decode p50', s['decode_step_ms_p50'], 'chunk p50', s.get('prefill_chunk_ms_p50'), 'max', s.get('prefill_chunk_ms_max'), 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'])"
