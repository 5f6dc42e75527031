#!/bin/bash
# round-2 call 23: host-side trims of the per-step path (slab versions, replay on the caller's stream)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c23_pytest.log; tail -4 gpurun_out/c23_pytest.log | cut -c1-220
timeout 600 python bench.py --no-extra > gpurun_out/c23_bench_decode.json 2> gpurun_out/c23_bench_decode.err; echo "bench decode rc=$?"; tail -c 300 gpurun_out/c23_bench_decode.err
python -c "
import json;d=json.load(open('gpurun_out/c23_bench_decode.json'));print(d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --workload serve --no-extra --no-cpu-baseline > gpurun_out/c23_bench_serve.json 2> gpurun_out/c23_bench_serve.err; echo "bench serve rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c23_bench_serve.json'));s=d['serving'];print(d['value'], 'decode p50', s['decode_step_ms_p50'], 'chunk p50', s.get('prefill_chunk_ms_p50'), 'prefill s', s['time_in_prefill_s'], 'decode s', s['time_in_decode_s'], s.get('prefill_chunk_ms_max'), s.get('prefill_chunks_over_2x_p50'), s.get('graph_captures'))"
