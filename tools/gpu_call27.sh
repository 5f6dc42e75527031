#!/bin/bash
# round-2 call 27: split-count policy of the swap-AB GEMM (cap / fixed-cost sweeps) at 64 and 128 rows, in isolation and in the step
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
kb() { tag=$1; shift; echo "== $tag"; env "$@" timeout 200 python tools/kbench.py --out gpurun_out/c27_kbench_$tag.json --batches 64,128 --only q,o,gate_up,down 2>&1 | tail -8; }
kb default
kb cap8 TL_SKINNY_MAX_SPLITS=8
kb cap4 TL_SKINNY_MAX_SPLITS=4
kb fixed4 TL_SKINNY_FIXED_COST=4
kb fixed20 TL_SKINNY_FIXED_COST=20
ab() { tag=$1; shift; env "$@" timeout 200 python tools/decode_ab.py --tag "$tag" --batch 64 --context 1024 --steps 32 2>&1 | tail -1; }
ab b64_default
ab b64_cap8 TL_SKINNY_MAX_SPLITS=8
ab b64_cap4 TL_SKINNY_MAX_SPLITS=4
ab b64_fixed4 TL_SKINNY_FIXED_COST=4
ab b64_fixed20 TL_SKINNY_FIXED_COST=20
