#!/bin/bash
# round-2 call 4: full GPU tests (skinny GEMM, TC decode attention, parity r2), decode variants without the carveout, kernel sweeps, serving
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -30 gpurun_out/c4_pytest.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -k "qwen3_4b_full_depth or config1_golden" -s > gpurun_out/c4_pytest_slow.log 2>&1; echo "pytest slow rc=$?" >> gpurun_out/c4_pytest_slow.log
tail -6 gpurun_out/c4_pytest_slow.log | cut -c1-250
ab() { tag=$1; shift; env "$@" timeout 300 python tools/decode_ab.py --tag "$tag" --steps 96 2>&1 | tail -1; }
ab full_nochain TL_S5_HALF=0 TL_CHAIN=0
ab full_chain TL_S5_HALF=0 TL_CHAIN=1
ab full_chain_res4 TL_S5_HALF=0 TL_CHAIN=1 TL_S5_RESERVE=4
ab full_nochain_res16 TL_S5_HALF=0 TL_CHAIN=0 TL_S5_RESERVE=16
ab half_chain TL_S5_HALF=1 TL_CHAIN=1
ab half_chain_carve50 TL_S5_HALF=1 TL_CHAIN=1 TL_S5_CARVEOUT=50
ab full_chain_carve25 TL_S5_HALF=0 TL_CHAIN=1 TL_S5_CARVEOUT=25
for B in 2 4 8 16 32 64; do env TL_S5_HALF=0 timeout 300 python tools/decode_ab.py --tag "b$B" --batch $B --steps 48 2>&1 | tail -1; done
env TL_S5_HALF=0 timeout 300 python tools/decode_ab.py --tag "b64_ctx1024" --batch 64 --context 1024 --steps 32 2>&1 | tail -1
env TL_S5_HALF=0 TL_ATTENTION_FUSED=1 timeout 300 python tools/decode_ab.py --tag "b64_ctx1024_fusedattn" --batch 64 --context 1024 --steps 32 2>&1 | tail -1
TL_S5_HALF=0 timeout 600 python tools/kbench.py --out gpurun_out/c4_kbench.json --batches 1,8,16,32,64,128 > gpurun_out/c4_kbench.log 2>&1; cat gpurun_out/c4_kbench.log | tail -50
TL_S5_HALF=0 TL_DECODE_TC=0 timeout 300 python tools/kbench.py --attention-only --out gpurun_out/c4_kbench_att_old.json 2>&1 | grep attention
TL_S5_HALF=0 TL_DECODE_TC=1 timeout 300 python tools/kbench.py --attention-only --out gpurun_out/c4_kbench_att_tc.json 2>&1 | grep attention
TL_S5_HALF=0 timeout 900 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c4_bench_serve.json 2> gpurun_out/c4_bench_serve.err; echo "bench serve rc=$?"; tail -c 400 gpurun_out/c4_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c4_bench_serve.json'));print(d['value'], d['serving'])"
