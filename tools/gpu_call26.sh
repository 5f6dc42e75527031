#!/bin/bash
# round-2 call 26: q/k norm + RoPE kernel with the trig shared per CTA
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c26_pytest.log; tail -4 gpurun_out/c26_pytest.log | cut -c1-220
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_decode_b64.csv python tools/launch_list.py --mode decode --batch 64 --context 1024 > gpurun_out/final_ll_b64.log 2>&1; tail -1 gpurun_out/final_ll_b64.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_chunk128.csv python tools/launch_list.py --mode chunk --chunk 128 --context 512 > gpurun_out/final_ll_chunk.log 2>&1; tail -1 gpurun_out/final_ll_chunk.log
grep -h "qk_norm" gpurun_out/final_launches_decode_b64.csv | head -3 | cut -c1-200
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/final_bench_serve.json 2> gpurun_out/final_bench_serve.err; echo "bench serve rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/final_bench_serve.json'));s=d['serving'];print(d['value'], 'decode p50', s['decode_step_ms_p50'], 'chunk p50', s.get('prefill_chunk_ms_p50'), 'max', s.get('prefill_chunk_ms_max'), 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'], d.get('extra'))"
