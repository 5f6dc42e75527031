#!/bin/bash
# round-2 call 13: per-block timeline of the skinny GEMM; launch lists of the 64-slot decode step and of a 128-token prefill chunk; merge kernel
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 200 python tools/skinny_blocks.py > gpurun_out/c13_blocks_head.txt 2>&1; tail -45 gpurun_out/c13_blocks_head.txt
TL_LIB=$T TL_SKINNY_TMEM_A=0 timeout 200 python tools/skinny_blocks.py > gpurun_out/c13_blocks_head_sa.txt 2>&1; tail -12 gpurun_out/c13_blocks_head_sa.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "attention or paged" > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.log; tail -3 gpurun_out/c13_pytest.log | cut -c1-220
timeout 200 python tools/kbench.py --attention-only --out gpurun_out/c13_kbench_att.json 2>&1 | tail -9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c13_launches_decode_b64.csv python tools/launch_list.py --mode decode --batch 64 --context 1024 > gpurun_out/c13_ll_decode.log 2>&1; tail -1 gpurun_out/c13_ll_decode.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c13_launches_chunk128.csv python tools/launch_list.py --mode chunk --chunk 128 --context 512 > gpurun_out/c13_ll_chunk.log 2>&1; tail -1 gpurun_out/c13_ll_chunk.log
