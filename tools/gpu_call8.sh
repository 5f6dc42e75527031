#!/bin/bash
# round-2 call 8: generation-6 streaming kernel (TMA rings, two CTAs per SM): parity subset, decode A/B, timeline; ncu of the round-2 kernels
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
TL_STREAM6=1 timeout 900 python -m pytest tests -m gpu -q -x -k "quantized_matmul_matches_oracle or fused_projection or swiglu_pairs or lm_head_one_hot or engine_step or device_resident or split_kv_attention or identity_activations" > gpurun_out/c8_pytest_s6.log 2>&1; echo "pytest s6 rc=$?" >> gpurun_out/c8_pytest_s6.log; tail -15 gpurun_out/c8_pytest_s6.log | cut -c1-220
ab() { tag=$1; shift; env "$@" timeout 300 python tools/decode_ab.py --tag "$tag" --steps 96 2>&1 | tail -1; }
ab base
ab stream6 TL_STREAM6=1
ab stream6_res0 TL_STREAM6=1 TL_S5_RESERVE=0
env TL_STREAM6=1 timeout 300 python tools/decode_ab.py --tag s6_b2 --batch 2 --steps 64 2>&1 | tail -1
env TL_STREAM6=1 timeout 300 python tools/decode_ab.py --tag s6_b4 --batch 4 --steps 64 2>&1 | tail -1
env TL_STREAM6=1 timeout 300 python tools/decode_ab.py --tag s6_b8 --batch 8 --steps 64 2>&1 | tail -1
timeout 300 python tools/decode_ab.py --tag base_b8 --batch 8 --steps 64 2>&1 | tail -1
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T TL_STREAM6=1 timeout 300 python tools/graph_timeline.py > gpurun_out/c8_timeline_s6.txt 2>&1; tail -7 gpurun_out/c8_timeline_s6.txt
TL_LIB=$T timeout 300 python tools/graph_timeline.py > gpurun_out/c8_timeline_base.txt 2>&1; tail -7 gpurun_out/c8_timeline_base.txt
TL_STREAM6=1 timeout 300 python tools/kbench.py --quick --only q,o,gate_up,down,lm_head --out gpurun_out/c8_kbench_s6.json 2>&1 | tail -11
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"paged_prefill_tc|w4a16_skinny_kernel|w4a16_stream5" -s 18 -c 6 -o gpurun_out/r02_kernels python tools/ncu_round2.py > gpurun_out/c8_ncu_kernels.log 2>&1; echo "ncu kernels rc=$?"; tail -2 gpurun_out/c8_ncu_kernels.log
ls -la gpurun_out/*.ncu-rep
