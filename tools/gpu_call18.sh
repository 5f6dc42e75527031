#!/bin/bash
# round-2 call 18: one "full" barrier per stage in both prefill GEMMs; batched loads in the reduce + RMSNorm kernel
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 120 python tools/gemm_bench.py 4096 > gpurun_out/c18_gemm2.txt 2>&1; tail -5 gpurun_out/c18_gemm2.txt | head -4
TL_GEMM2=0 timeout 120 python tools/gemm_bench.py 4096 > gpurun_out/c18_gemm1.txt 2>&1; tail -5 gpurun_out/c18_gemm1.txt | head -4
TL_GEMM2=0 TL_GEMM_MT=2 timeout 120 python tools/gemm_bench.py 4096 2>&1 | head -4
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_GEMM2=0 TL_LIB=$T timeout 100 python tools/gemm_blocks.py 4096 2560 19456 > gpurun_out/c18_gemm_blocks_gate_up.txt 2>&1; sed -n 1,2p gpurun_out/c18_gemm_blocks_gate_up.txt; sed -n 14,22p gpurun_out/c18_gemm_blocks_gate_up.txt
TL_GEMM2=0 TL_LIB=$T timeout 100 python tools/gemm_blocks.py 4096 9728 2560 > gpurun_out/c18_gemm_blocks_down.txt 2>&1; sed -n 1,2p gpurun_out/c18_gemm_blocks_down.txt; sed -n 14,22p gpurun_out/c18_gemm_blocks_down.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest.log; tail -5 gpurun_out/c18_pytest.log | cut -c1-220
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c18_bench_serve.json 2> gpurun_out/c18_bench_serve.err; echo "bench serve rc=$?"; tail -c 300 gpurun_out/c18_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c18_bench_serve.json'));print(d['value'], d['serving'])"
