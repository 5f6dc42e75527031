#!/bin/bash
# round-2 call 3: calibrate the decode regression against the round-1 tree on the SAME box
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,power.limit,temperature.gpu --format=csv > gpurun_out/c3_smi.txt
(cd _r1 && timeout 600 python bench.py --no-cpu-baseline > ../gpurun_out/c3_r1_bench.json 2> ../gpurun_out/c3_r1_bench.err); echo "r1 bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/c3_r1_bench.json'));print('r1 code:',d['ms_per_step'],d['roofline']['avg_launch_us'],d['clocks'])"
ab() { tag=$1; shift; env "$@" timeout 300 python tools/decode_ab.py --tag "$tag" --steps 96 2>&1 | tail -1; }
ab full_nochain TL_S5_HALF=0 TL_CHAIN=0
ab full_nochain_nocarve_all TL_S5_HALF=0 TL_CHAIN=0 TL_S5_CARVEOUT=0
ab full_nochain_att1 TL_S5_HALF=0 TL_CHAIN=0 TL_LIB=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_att1.so
ab full_nochain_att1_nocarve TL_S5_HALF=0 TL_CHAIN=0 TL_S5_CARVEOUT=0 TL_LIB=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_att1.so
ab half_chain_nocarve TL_S5_HALF=1 TL_CHAIN=1 TL_S5_CARVEOUT=0
ab full_nochain_unfusedattn TL_S5_HALF=0 TL_CHAIN=0 TL_ATTENTION_FUSED=0
(cd _r1 && timeout 300 python tools/kbench.py --quick --out ../gpurun_out/c3_r1_kbench.json > ../gpurun_out/c3_r1_kbench.log 2>&1); tail -16 gpurun_out/c3_r1_kbench.log
timeout 300 python tools/kbench.py --quick --out gpurun_out/c3_kbench.json > gpurun_out/c3_kbench.log 2>&1; tail -16 gpurun_out/c3_kbench.log
TL_S5_HALF=0 timeout 300 python tools/kbench.py --quick --only q,o,gate_up,down,lm_head --out gpurun_out/c3_kbench_full.json 2>&1 | tail -12
