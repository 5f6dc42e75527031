#!/bin/bash
# round-2 call 11: skinny GEMM with two alternating dequantiser groups (thread = row), scales from L2; dynamic split-KV dealing
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log; tail -6 gpurun_out/c11_pytest.log | cut -c1-220
timeout 300 python tools/skinny_stress.py 30 > gpurun_out/c11_stress.log 2>&1; grep -v "^  run" gpurun_out/c11_stress.log | tail -4
timeout 300 python tools/kbench.py --out gpurun_out/c11_kbench.json --batches 16,64,128 --only q,o,gate_up,down,lm_head 2>&1 | tail -16
timeout 200 python tools/kbench.py --attention-only --out gpurun_out/c11_kbench_att.json 2>&1 | tail -12
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 200 python tools/skinny_timeline.py 2>&1 | tail -6 | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_skinny_kernel -s 3 -c 1 -f -o gpurun_out/c11_skinny_head python tools/ncu_skinny.py > gpurun_out/c11_ncu.log 2>&1; tail -2 gpurun_out/c11_ncu.log
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c11_bench_serve.json 2> gpurun_out/c11_bench_serve.err; echo "bench serve rc=$?"; tail -c 300 gpurun_out/c11_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c11_bench_serve.json'));print(d['value'], d['serving'])"
