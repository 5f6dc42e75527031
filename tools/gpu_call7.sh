#!/bin/bash
# round-2 call 7: skinny GEMM timeline + stress after the reduce-kernel rewrite; serve A/B (chunk graph on/off); ncu captures
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
T=$PWD/tiny-llm_b200/extensions_b200/tiny_llm_ext_b200/libtiny_llm_b200_trace.so
TL_LIB=$T timeout 300 python tools/skinny_timeline.py 2>&1 | tail -8 | tee gpurun_out/c7_skinny_timeline.txt
timeout 600 python tools/skinny_stress.py 60 > gpurun_out/c7_stress.log 2>&1; grep -v "^  run" gpurun_out/c7_stress.log | tail -12
timeout 900 python -m pytest tests -m gpu -q -x -k "quantized_matmul or fused_projection or swiglu_pairs or engine_batch or prefill_chunk" > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log; tail -5 gpurun_out/c7_pytest.log | cut -c1-200
timeout 600 python tools/kbench.py --out gpurun_out/c7_kbench.json --batches 16,64,128 --only q,kv,o,gate_up,down,lm_head 2>&1 | tail -19
for B in 16 64; do timeout 300 python tools/decode_ab.py --tag "b$B" --batch $B --steps 48 2>&1 | tail -1; done
timeout 300 python tools/decode_ab.py --tag "b64_ctx1024" --batch 64 --context 1024 --steps 32 2>&1 | tail -1
for PG in 128 0; do
TL_PREFILL_GRAPH=$PG timeout 900 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c7_bench_serve_pg$PG.json 2> gpurun_out/c7_bench_serve_pg$PG.err; echo "bench serve pg=$PG rc=$?"; tail -c 300 gpurun_out/c7_bench_serve_pg$PG.err
python -c "
import json;d=json.load(open('gpurun_out/c7_bench_serve_pg$PG.json'));print(d['value'], d['serving'])"
done
# ---- ncu: launch list + DRAM traffic of one decode token, then --set full of the round-2 kernels
timeout 900 ncu --nvtx --nvtx-include "decode/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_decode.csv python tools/profile_decode.py --layers 36 --steps 1 --pdl > gpurun_out/c7_ncu_decode.log 2>&1; echo "ncu decode rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"paged_prefill_tc|w4a16_skinny_kernel|w4a16_stream5" -s 18 -c 6 -o gpurun_out/r02_kernels python tools/ncu_round2.py > gpurun_out/c7_ncu_kernels.log 2>&1; echo "ncu kernels rc=$?"; tail -3 gpurun_out/c7_ncu_kernels.log
ls -la gpurun_out/*.ncu-rep
