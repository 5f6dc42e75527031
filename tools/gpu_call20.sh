#!/bin/bash
# round-2 call 20: two CTAs per SM for the 128-column swap-AB GEMM; CPU reference arm pinned; batch-1 launch list
mkdir -p gpurun_out; rm -f gpurun_out/decode_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.log; tail -4 gpurun_out/c20_pytest.log | cut -c1-220
timeout 300 python tools/skinny_stress.py 30 > gpurun_out/c20_stress.log 2>&1; grep -v "^  run" gpurun_out/c20_stress.log | tail -4
timeout 300 python tools/kbench.py --out gpurun_out/c20_kbench.json --batches 96,128 --only q,o,gate_up,down,lm_head 2>&1 | tail -11
TL_SKINNY_WIDE1=1 timeout 300 python tools/kbench.py --out gpurun_out/c20_kbench_wide1.json --batches 96,128 --only q,o,gate_up,down,lm_head 2>&1 | tail -11
for i in 1 2; do
timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c20_bench_serve_$i.json 2> gpurun_out/c20_bench_serve_$i.err; echo "bench serve rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c20_bench_serve_$i.json'))['serving'];print(d['output_tok_s'], 'decode p50', d['decode_step_ms_p50'], 'chunk p50', d.get('prefill_chunk_ms_p50'), 'prefill s', d['time_in_prefill_s'], 'decode s', d['time_in_decode_s'])"
done
TL_SKINNY_WIDE1=1 timeout 600 python bench.py --workload serve --no-cpu-baseline > gpurun_out/c20_bench_serve_w1.json 2> gpurun_out/c20_bench_serve_w1.err
python -c "
import json;d=json.load(open('gpurun_out/c20_bench_serve_w1.json'))['serving'];print('wide1', d['output_tok_s'], 'decode p50', d['decode_step_ms_p50'], 'chunk p50', d.get('prefill_chunk_ms_p50'), 'prefill s', d['time_in_prefill_s'], 'decode s', d['time_in_decode_s'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_decode_b1.csv python tools/launch_list.py --mode decode --batch 1 --context 128 > gpurun_out/final_ll_b1.log 2>&1; tail -1 gpurun_out/final_ll_b1.log
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "bench reference rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference_2.json 2> gpurun_out/final_bench_reference_2.err
python -c "
import json
for f in ['final_bench_reference.json','final_bench_reference_2.json']:
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['cpu_baseline']['sample'][-160:])"
