#!/bin/bash
# round-2 call 9: per-warp barrier arrivals (skinny / GEMM / attention), ex2 softmax, coalesced GEMM feed; generation-6 streaming kernel retest
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log; tail -6 gpurun_out/c9_pytest.log | cut -c1-220
timeout 300 python tools/skinny_stress.py 30 > gpurun_out/c9_stress.log 2>&1; grep -v "^  run" gpurun_out/c9_stress.log | tail -4
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 200 python tools/skinny_timeline.py 2>&1 | tail -6 | cut -c1-200
timeout 400 python tools/kbench.py --out gpurun_out/c9_kbench.json --batches 16,64,128 --only q,kv,o,gate_up,down,lm_head 2>&1 | tail -19
timeout 400 python bench.py --workload prefill --no-cpu-baseline --steps 4 > gpurun_out/c9_bench_prefill.json 2> gpurun_out/c9_bench_prefill.err; echo "bench prefill rc=$?"; tail -c 300 gpurun_out/c9_bench_prefill.err
python -c "
import json;d=json.load(open('gpurun_out/c9_bench_prefill.json'));print('prefill', d['value'], d['roofline']['achieved'], d['extra']['attention_roofline']['achieved'], d['extra'].get('chunked'))"
timeout 200 python tools/decode_ab.py --tag "b64_ctx1024" --batch 64 --context 1024 --steps 32 2>&1 | tail -1
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c9_bench_serve.json 2> gpurun_out/c9_bench_serve.err; echo "bench serve rc=$?"; tail -c 300 gpurun_out/c9_bench_serve.err
python -c "
import json;d=json.load(open('gpurun_out/c9_bench_serve.json'));print(d['value'], d['serving'])"
# ---- generation 6 (opt-in)
TL_STREAM6=1 timeout 400 python -m pytest tests -m gpu -q -x -k "quantized_matmul_matches_oracle or fused_projection or swiglu_pairs or lm_head_one_hot or engine_step or device_resident or split_kv_attention or identity_activations" > gpurun_out/c9_pytest_s6.log 2>&1; echo "pytest s6 rc=$?" >> gpurun_out/c9_pytest_s6.log; tail -8 gpurun_out/c9_pytest_s6.log | cut -c1-220
ab() { tag=$1; shift; env "$@" timeout 150 python tools/decode_ab.py --tag "$tag" --steps 96 2>&1 | tail -1; }
ab base
ab stream6 TL_STREAM6=1
ab stream6_res0 TL_STREAM6=1 TL_S5_RESERVE=0
env TL_STREAM6=1 timeout 150 python tools/decode_ab.py --tag s6_b4 --batch 4 --steps 64 2>&1 | tail -1
TL_LIB=$T TL_STREAM6=1 timeout 150 python tools/graph_timeline.py > gpurun_out/c9_timeline_s6.txt 2>&1; tail -7 gpurun_out/c9_timeline_s6.txt
TL_STREAM6=1 timeout 150 python tools/kbench.py --quick --only q,o,gate_up,down,lm_head --out gpurun_out/c9_kbench_s6.json 2>&1 | tail -11
