#!/bin/bash
# round-2 call 29: flash attention with P in tensor memory (TS-form P V), A/B against P through shared memory
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not (qwen3_4b_full_depth or config1_golden)" > gpurun_out/c29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c29_pytest.log; tail -6 gpurun_out/c29_pytest.log | cut -c1-220
for v in 1 0; do
TL_FA_P_TMEM=$v timeout 400 python bench.py --workload prefill --no-cpu-baseline --no-extra --steps 4 > gpurun_out/c29_bench_prefill_$v.json 2> gpurun_out/c29_bench_prefill_$v.err; echo "bench prefill P_TMEM=$v rc=$?"; tail -c 200 gpurun_out/c29_bench_prefill_$v.err
python -c "
import json;d=json.load(open('gpurun_out/c29_bench_prefill_$v.json'));print('prefill', d['value'], d['roofline']['achieved'], d['extra']['attention_roofline']['achieved'])"
TL_FA_P_TMEM=$v timeout 200 python tools/kbench.py --attention-only --out gpurun_out/c29_kbench_att_$v.json 2>&1 | tail -9 | head -8
done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:paged_prefill_tc -s 6 -c 1 --csv --log-file gpurun_out/c29_ncu_fa.csv python tools/ncu_round2.py > gpurun_out/c29_ncu.log 2>&1; grep -v "^==" gpurun_out/c29_ncu_fa.csv | cut -c1-30,200-400 | tail -5
