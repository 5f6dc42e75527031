#!/bin/bash
# round-2 call 6: skinny GEMM after the race fix / deeper activation ring / batched finisher; chunk-graph prefill; serving
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 600 python tools/skinny_stress.py 60 > gpurun_out/c6_stress.log 2>&1; tail -12 gpurun_out/c6_stress.log
timeout 1500 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
tail -25 gpurun_out/c6_pytest.log | cut -c1-250
timeout 600 python tools/kbench.py --out gpurun_out/c6_kbench.json --batches 16,32,64,128 --only q,kv,o,gate_up,down,lm_head 2>&1 | tail -26
for B in 16 64; do timeout 300 python tools/decode_ab.py --tag "b$B" --batch $B --steps 48 2>&1 | tail -1; done
timeout 300 python tools/decode_ab.py --tag "b64_ctx1024" --batch 64 --context 1024 --steps 32 2>&1 | tail -1
timeout 300 python tools/decode_ab.py --tag "b1" --steps 96 2>&1 | tail -1
timeout 900 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c6_bench_serve.json 2> gpurun_out/c6_bench_serve.err; echo "bench serve rc=$?"; tail -c 600 gpurun_out/c6_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c6_bench_serve.json'));print(d['value'], d['serving'])"
timeout 600 python bench.py --workload prefill --no-cpu-baseline --steps 4 > gpurun_out/c6_bench_prefill.json 2> gpurun_out/c6_bench_prefill.err; echo "bench prefill rc=$?"; tail -c 300 gpurun_out/c6_bench_prefill.err
python -c "
import json;d=json.load(open('gpurun_out/c6_bench_prefill.json'));print('prefill', d['value'], d['roofline']['achieved'], d['extra']['attention_roofline']['achieved'], d['extra'].get('chunked'))"
