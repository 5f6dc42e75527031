#!/bin/bash
# round-2 call 19: per-block timeline of the CTA-pair GEMM; serve repeat (run-to-run spread)
mkdir -p gpurun_out
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 100 python tools/gemm_blocks.py 4096 2560 19456 > gpurun_out/c19_gemm2_blocks_gate_up.txt 2>&1; sed -n 1,2p gpurun_out/c19_gemm2_blocks_gate_up.txt; sed -n 10,30p gpurun_out/c19_gemm2_blocks_gate_up.txt
TL_LIB=$T timeout 100 python tools/gemm_blocks.py 4096 9728 2560 > gpurun_out/c19_gemm2_blocks_down.txt 2>&1; sed -n 1,2p gpurun_out/c19_gemm2_blocks_down.txt; sed -n 14,24p gpurun_out/c19_gemm2_blocks_down.txt
for i in 1 2; do
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c19_bench_serve_$i.json 2> gpurun_out/c19_bench_serve_$i.err; echo "bench serve rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c19_bench_serve_$i.json'))['serving'];print(d['output_tok_s'], 'decode p50', d['decode_step_ms_p50'], 'chunk p50', d.get('prefill_chunk_ms_p50'), 'prefill s', d['time_in_prefill_s'], 'decode s', d['time_in_decode_s'])"
done
