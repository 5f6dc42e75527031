#!/bin/bash
# round-2 call 30: flash attention with P in tensor memory, arrival fixed; refresh of the artefacts it changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/final_pytest.log; tail -4 gpurun_out/final_pytest.log | cut -c1-200
for i in 1 2 3 4 5 6; do timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "full_size_prefill_attention or paged_attention" 2>&1 | tail -1; done
if [ $rc -ne 0 ]; then echo "tests failed: artefacts not refreshed"; exit 1; fi
timeout 900 python bench.py > gpurun_out/final_bench_decode.json 2> gpurun_out/final_bench_decode.err; echo "bench decode rc=$?"
timeout 600 python bench.py --workload prefill > gpurun_out/final_bench_prefill.json 2> gpurun_out/final_bench_prefill.err; echo "bench prefill rc=$?"
timeout 600 python bench.py --workload serve > gpurun_out/final_bench_serve.json 2> gpurun_out/final_bench_serve.err; echo "bench serve rc=$?"
timeout 900 python bench.py --workload serve8k > gpurun_out/final_bench_serve8k.json 2> gpurun_out/final_bench_serve8k.err; echo "bench serve8k rc=$?"
python - <<'PY'
import json
for w in ["decode", "prefill", "serve", "serve8k"]:
    d = json.load(open(f"gpurun_out/final_bench_{w}.json"))
    print(w, d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"), (d.get("extra") or {}).get("prefill", {}).get("attention_roofline", {}).get("achieved") if w == "decode" else (d.get("extra") or {}).get("attention_roofline", {}).get("achieved"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"paged_prefill_tc|w4a16_skinny_kernel|w4a16_stream5|w4a16_gemm2" -s 21 -c 7 -f -o gpurun_out/final_kernels python tools/ncu_round2.py > gpurun_out/final_ncu.log 2>&1; tail -1 gpurun_out/final_ncu.log
timeout 200 python tools/kbench.py --attention-only --out gpurun_out/final_kbench_att.json 2>&1 | tail -9 | head -8
