#!/bin/bash
# last sanity run of the final tree: every GPU test, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/last_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/last_pytest.log; tail -3 gpurun_out/last_pytest.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench_decode.json 2> gpurun_out/final_bench_decode.err; echo "bench decode rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/final_bench_decode.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
