#!/bin/bash
# round-2 call 2: full GPU tests, decode hand-off A/B (flag chain x half CTAs), timelines, first runs of the new bench workloads
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x -k "not (paged_attention_matches_oracle or prefill)" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -4 gpurun_out/c2_pytest.log
timeout 900 python -m pytest tests -m gpu -q -k "paged_attention_matches_oracle or prefill" > gpurun_out/c2_pytest_tc.log 2>&1; echo "pytest tc rc=$?" >> gpurun_out/c2_pytest_tc.log
tail -25 gpurun_out/c2_pytest_tc.log | cut -c1-300
ab() { tag=$1; shift; env "$@" timeout 300 python tools/decode_ab.py --tag "$tag" --steps 96 2>&1 | tail -1; }
ab full_nochain TL_S5_HALF=0 TL_CHAIN=0
ab full_nochain_nocarve TL_S5_HALF=0 TL_CHAIN=0 TL_S5_CARVEOUT=0
ab full_chain TL_S5_HALF=0 TL_CHAIN=1
ab full_chain_res0 TL_S5_HALF=0 TL_CHAIN=1 TL_S5_RESERVE=0
ab half_nochain TL_S5_HALF=1 TL_CHAIN=0
ab half_chain TL_S5_HALF=1 TL_CHAIN=1
ab half_chain_148 TL_S5_HALF=1 TL_CHAIN=1 TL_S5_GRID=2560:148,6144:148,19456:148
ab half_chain_gu296 TL_S5_HALF=1 TL_CHAIN=1 TL_S5_GRID=19456:304,6144:192
env TL_S5_HALF=1 TL_CHAIN=1 timeout 300 python tools/decode_ab.py --tag half_chain_b4 --batch 4 --steps 96 2>&1 | tail -1
env TL_S5_HALF=0 TL_CHAIN=1 timeout 300 python tools/decode_ab.py --tag full_chain_b4 --batch 4 --steps 96 2>&1 | tail -1
env TL_S5_HALF=0 TL_CHAIN=1 timeout 300 python tools/decode_ab.py --tag full_chain_b32 --batch 32 --steps 48 2>&1 | tail -1
env TL_S5_HALF=0 TL_CHAIN=0 timeout 300 python tools/decode_ab.py --tag full_nochain_b32 --batch 32 --steps 48 2>&1 | tail -1
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T TL_S5_HALF=1 TL_CHAIN=1 timeout 300 python tools/graph_timeline.py > gpurun_out/c2_timeline_half_chain.txt 2>&1
TL_LIB=$T TL_S5_HALF=0 TL_CHAIN=1 timeout 300 python tools/graph_timeline.py > gpurun_out/c2_timeline_full_chain.txt 2>&1
TL_LIB=$T TL_S5_HALF=0 TL_CHAIN=0 timeout 300 python tools/graph_timeline.py > gpurun_out/c2_timeline_full_nochain.txt 2>&1
tail -7 gpurun_out/c2_timeline_half_chain.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c2_bench_decode.json 2> gpurun_out/c2_bench_decode.err; echo "bench decode rc=$?"; tail -c 600 gpurun_out/c2_bench_decode.err
timeout 600 python bench.py --workload prefill --no-cpu-baseline --steps 4 > gpurun_out/c2_bench_prefill.json 2> gpurun_out/c2_bench_prefill.err; echo "bench prefill rc=$?"; tail -c 600 gpurun_out/c2_bench_prefill.err
timeout 900 python bench.py --workload serve > gpurun_out/c2_bench_serve.json 2> gpurun_out/c2_bench_serve.err; echo "bench serve rc=$?"; tail -c 600 gpurun_out/c2_bench_serve.err
