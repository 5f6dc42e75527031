#!/bin/bash
# round-2 call 1: sanity tests + decode launch-geometry A/B + timeline
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python tools/decode_ab.py --tag "$tag" 2>&1 | tail -1; }
run full16 TL_S5_HALF=0
run half_default TL_S5_HALF=1
run half_148 TL_S5_HALF=1 TL_S5_GRID=2560:148,6144:148,19456:148
run half_qkv192 TL_S5_HALF=1 TL_S5_GRID=6144:192
run half_gu296 TL_S5_HALF=1 TL_S5_GRID=19456:304,6144:192
run half_nopdl TL_S5_HALF=1 TL_PDL=0
env TL_S5_HALF=1 timeout 300 python tools/decode_ab.py --tag half_b2 --batch 2 2>&1 | tail -1
env TL_S5_HALF=0 timeout 300 python tools/decode_ab.py --tag full_b2 --batch 2 2>&1 | tail -1
env TL_S5_HALF=1 timeout 300 python tools/decode_ab.py --tag half_b4 --batch 4 2>&1 | tail -1
env TL_S5_HALF=0 timeout 300 python tools/decode_ab.py --tag full_b4 --batch 4 2>&1 | tail -1
env TL_S5_HALF=1 timeout 300 python tools/decode_ab.py --tag half_ctx4k --context 4096 --steps 64 2>&1 | tail -1
TL_LIB=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so TL_S5_HALF=1 timeout 300 python tools/graph_timeline.py > gpurun_out/c1_timeline_half.txt 2>&1
TL_LIB=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so TL_S5_HALF=0 timeout 300 python tools/graph_timeline.py > gpurun_out/c1_timeline_full.txt 2>&1
cat gpurun_out/c1_timeline_half.txt | tail -8
