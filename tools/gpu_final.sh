#!/bin/bash
# round-2 final evidence run (one GPU): tests, the bench lines of every workload, ncu captures.  Outputs under gpurun_out/final_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/final_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -4 gpurun_out/final_pytest.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/final_bench_decode.json 2> gpurun_out/final_bench_decode.err; echo "bench decode rc=$?"; tail -c 200 gpurun_out/final_bench_decode.err
timeout 600 python bench.py --workload prefill > gpurun_out/final_bench_prefill.json 2> gpurun_out/final_bench_prefill.err; echo "bench prefill rc=$?"
timeout 600 python bench.py --workload serve > gpurun_out/final_bench_serve.json 2> gpurun_out/final_bench_serve.err; echo "bench serve rc=$?"
timeout 900 python bench.py --workload serve8k > gpurun_out/final_bench_serve8k.json 2> gpurun_out/final_bench_serve8k.err; echo "bench serve8k rc=$?"; tail -c 300 gpurun_out/final_bench_serve8k.err
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "bench reference rc=$?"
python - <<'PY'
import json
for w in ["decode", "prefill", "serve", "serve8k", "reference"]:
    try:
        d = json.load(open(f"gpurun_out/final_bench_{w}.json"))
        print(w, d.get("value"), d.get("unit"), "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(w, "unreadable", e)
PY
# launch lists (per-kernel durations; cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_decode_b1.csv python tools/launch_list.py --mode decode --batch 1 --context 128 > gpurun_out/final_ll_b1.log 2>&1; tail -1 gpurun_out/final_ll_b1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_decode_b64.csv python tools/launch_list.py --mode decode --batch 64 --context 1024 > gpurun_out/final_ll_b64.log 2>&1; tail -1 gpurun_out/final_ll_b64.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_chunk128.csv python tools/launch_list.py --mode chunk --chunk 128 --context 512 > gpurun_out/final_ll_chunk.log 2>&1; tail -1 gpurun_out/final_ll_chunk.log
# full captures of the seven kernels (fourth round of tools/ncu_round2.py)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"paged_prefill_tc|w4a16_skinny_kernel|w4a16_stream5|w4a16_gemm2" -s 21 -c 7 -f -o gpurun_out/final_kernels python tools/ncu_round2.py > gpurun_out/final_ncu.log 2>&1; tail -2 gpurun_out/final_ncu.log
timeout 300 python tools/kbench.py --out gpurun_out/final_kbench.json --batches 1,8,16,64,128 --only q,kv,o,gate_up,down,lm_head 2>&1 | tail -32
timeout 200 python tools/kbench.py --attention-only --out gpurun_out/final_kbench_att.json 2>&1 | tail -9
timeout 120 python tools/gemm_bench.py 4096 > gpurun_out/final_gemm_bench.txt 2>&1; tail -5 gpurun_out/final_gemm_bench.txt | head -4
