#!/bin/bash
# round-2 call 15: skinny GEMM on 128-wide group blocks (one barrier round per quantisation group)
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest.log; tail -6 gpurun_out/c15_pytest.log | cut -c1-220
timeout 300 python tools/skinny_stress.py 30 > gpurun_out/c15_stress.log 2>&1; grep -v "^  run" gpurun_out/c15_stress.log | tail -4
timeout 300 python tools/kbench.py --out gpurun_out/c15_kbench.json --batches 16,64,128 --only q,o,gate_up,down,lm_head 2>&1 | tail -16
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 200 python tools/skinny_blocks.py > gpurun_out/c15_blocks_head.txt 2>&1; sed -n 1,3p gpurun_out/c15_blocks_head.txt; sed -n 8,24p gpurun_out/c15_blocks_head.txt
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c15_bench_serve.json 2> gpurun_out/c15_bench_serve.err; echo "bench serve rc=$?"; tail -c 300 gpurun_out/c15_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c15_bench_serve.json'));print(d['value'], d['serving'])"
