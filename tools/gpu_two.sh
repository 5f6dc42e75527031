#!/bin/bash
# two GPUs, one process each (NCCL weight broadcast, i mod N request sharding): the bench lines the driver runs at N > 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
run() { out=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; tail -c 400 gpurun_out/$out.err | tail -3; }
run two_bench_decode --no-extra
run two_bench_serve8k --workload serve8k
run two_bench_serve --workload serve
run two_bench_reference --impl reference
python - <<'PY'
import json
for w in ["decode", "serve8k", "serve", "reference"]:
    try:
        d = json.loads(open(f"gpurun_out/two_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, d.get("value"), d.get("unit"), "n_gpus", d.get("n_gpus"), "ms/step", d.get("ms_per_step"), "launches", d.get("gpu_launches"), (d.get("config") or {}).get("parallelism"))
    except Exception as e:
        print(w, "unreadable", e)
PY
