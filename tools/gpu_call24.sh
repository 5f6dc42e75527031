#!/bin/bash
# round-2 call 24: row-variant step graphs (16 / 32 rows of a 64-slot engine)
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c24_pytest.log; tail -6 gpurun_out/c24_pytest.log | cut -c1-220
for i in 1 2; do
timeout 600 python bench.py --workload serve --no-extra --no-cpu-baseline > gpurun_out/c24_bench_serve_$i.json 2> gpurun_out/c24_bench_serve_$i.err; echo "bench serve rc=$?"; tail -c 200 gpurun_out/c24_bench_serve_$i.err
python -c "
import json;d=json.load(open('gpurun_out/c24_bench_serve_$i.json'));s=d['serving'];print(d['value'], 'decode p50', s['decode_step_ms_p50'], 'p95', s['decode_step_ms_p95'], 'chunk p50', s.get('prefill_chunk_ms_p50'), 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'], s.get('graph_captures'))"
done
TL_ROW_VARIANTS=0 timeout 600 python bench.py --workload serve --no-extra --no-cpu-baseline > gpurun_out/c24_bench_serve_full.json 2> gpurun_out/c24_bench_serve_full.err
python -c "
import json;d=json.load(open('gpurun_out/c24_bench_serve_full.json'));s=d['serving'];print('full rows', d['value'], 'decode p50', s['decode_step_ms_p50'], 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'])"
timeout 900 python bench.py --workload serve8k --no-cpu-baseline > gpurun_out/c24_bench_serve8k.json 2> gpurun_out/c24_bench_serve8k.err; echo "bench serve8k rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c24_bench_serve8k.json'));s=d['serving'];print(d['value'], 'decode p50', s['decode_step_ms_p50'], 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'])"
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
