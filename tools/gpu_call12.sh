#!/bin/bash
# round-2 call 12: skinny GEMM with the dequantised tile in tensor memory (A operand from TMEM); L2 bulk prefetch of the streaming kernel's CTA slice
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.log; tail -6 gpurun_out/c12_pytest.log | cut -c1-220
timeout 300 python tools/skinny_stress.py 30 > gpurun_out/c12_stress.log 2>&1; grep -v "^  run" gpurun_out/c12_stress.log | tail -4
timeout 300 python tools/kbench.py --out gpurun_out/c12_kbench_ta.json --batches 16,64,128 --only q,o,gate_up,down,lm_head 2>&1 | tail -16
TL_SKINNY_TMEM_A=0 timeout 300 python tools/kbench.py --out gpurun_out/c12_kbench_sa.json --batches 64 --only q,o,gate_up,down,lm_head 2>&1 | tail -6
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64_ta
ab b64_sa TL_SKINNY_TMEM_A=0
env timeout 200 python tools/decode_ab.py --tag b16 --batch 16 --context 1024 --steps 32 2>&1 | tail -1
ab1() { tag=$1; shift; env "$@" timeout 150 python tools/decode_ab.py --tag "$tag" --steps 96 2>&1 | tail -1; }
ab1 b1_l2pf
ab1 b1_nol2pf TL_S5_L2PREFETCH=0
ab1 b1_l2pf_res0 TL_S5_RESERVE=0
env timeout 150 python tools/decode_ab.py --tag b1_ctx4096 --context 4096 --steps 64 2>&1 | tail -1
env TL_S5_L2PREFETCH=0 timeout 150 python tools/decode_ab.py --tag b1_ctx4096_nol2pf --context 4096 --steps 64 2>&1 | tail -1
env timeout 150 python tools/decode_ab.py --tag b4 --batch 4 --steps 64 2>&1 | tail -1
env TL_S5_L2PREFETCH=0 timeout 150 python tools/decode_ab.py --tag b4_nol2pf --batch 4 --steps 64 2>&1 | tail -1
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 150 python tools/graph_timeline.py > gpurun_out/c12_timeline.txt 2>&1; tail -7 gpurun_out/c12_timeline.txt | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_skinny_kernel -s 3 -c 1 -f -o gpurun_out/c12_skinny_head python tools/ncu_skinny.py > gpurun_out/c12_ncu.log 2>&1; tail -2 gpurun_out/c12_ncu.log
