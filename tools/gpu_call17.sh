#!/bin/bash
# round-2 call 17: CTA-pair GEMM (cta_group::2), residual projection + next RMSNorm in the reduction kernel
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 120 python tools/gemm_bench.py 4096 > gpurun_out/c17_gemm2.txt 2>&1; tail -6 gpurun_out/c17_gemm2.txt | cut -c1-200
TL_GEMM2=0 timeout 120 python tools/gemm_bench.py 4096 > gpurun_out/c17_gemm1.txt 2>&1; tail -5 gpurun_out/c17_gemm1.txt | head -4
timeout 120 python tools/gemm_bench.py 512 2>&1 | head -4
TL_GEMM2=0 timeout 120 python tools/gemm_bench.py 512 2>&1 | head -4
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log; tail -12 gpurun_out/c17_pytest.log | cut -c1-220
timeout 400 python bench.py --workload prefill --no-cpu-baseline --steps 4 > gpurun_out/c17_bench_prefill.json 2> gpurun_out/c17_bench_prefill.err; echo "bench prefill rc=$?"; tail -c 300 gpurun_out/c17_bench_prefill.err
python -c "
import json;d=json.load(open('gpurun_out/c17_bench_prefill.json'));print('prefill', d['value'], d['roofline']['achieved'], d['extra']['attention_roofline']['achieved'], d['extra'].get('chunked'))"
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c17_bench_serve.json 2> gpurun_out/c17_bench_serve.err; echo "bench serve rc=$?"; tail -c 300 gpurun_out/c17_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c17_bench_serve.json'));print(d['value'], d['serving'])"
