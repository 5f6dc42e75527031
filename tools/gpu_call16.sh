#!/bin/bash
# round-2 call 16: per-block timeline of the tiled GEMM (four Qwen3-4B projection shapes at 4096 rows)
mkdir -p gpurun_out
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
for shape in "4096 2560 19456" "4096 9728 2560" "4096 4096 2560" "4096 2560 6144"; do
  TL_LIB=$T timeout 200 python tools/gemm_blocks.py $shape > "gpurun_out/c16_gemm_blocks_$(echo $shape | tr ' ' 'x').txt" 2>&1
  head -2 "gpurun_out/c16_gemm_blocks_$(echo $shape | tr ' ' 'x').txt"; sed -n 12,30p "gpurun_out/c16_gemm_blocks_$(echo $shape | tr ' ' 'x').txt"
done
