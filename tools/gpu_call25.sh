#!/bin/bash
# round-2 call 25: tail chunks of 2..8 tokens through the chunk graph; serve spread
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.log; tail -6 gpurun_out/c25_pytest.log | cut -c1-220
for i in 1 2 3; do
timeout 600 python bench.py --workload serve --no-extra --no-cpu-baseline > gpurun_out/c25_bench_serve_$i.json 2> gpurun_out/c25_bench_serve_$i.err; echo "bench serve rc=$?"; tail -c 200 gpurun_out/c25_bench_serve_$i.err
python -c "
import json;d=json.load(open('gpurun_out/c25_bench_serve_$i.json'));s=d['serving'];print(d['value'], 'decode p50', s['decode_step_ms_p50'], 'chunk p50', s.get('prefill_chunk_ms_p50'), 'max', s.get('prefill_chunk_ms_max'), 'over2x', s.get('prefill_chunks_over_2x_p50'), 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'], s.get('row_variant_replays'))"
done
