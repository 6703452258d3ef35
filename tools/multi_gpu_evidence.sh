#!/bin/bash
# Under gpurun --gpus N: batch and slide bench lines at N GPUs + the N-rank bit-identity check of the strip-sharded slide path.
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/${TAG}_bench_batch_${N}gpu.json 2> gpurun_out/batch$N.err; tail -c 300 gpurun_out/batch$N.err
$TR bench.py --gpus $N --workload wsi --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/${TAG}_bench_wsi_${N}gpu.json 2> gpurun_out/wsi$N.err; tail -c 300 gpurun_out/wsi$N.err
$TR tools/wsi_check.py 2>&1 | tail -3 | tee gpurun_out/${TAG}_wsi_check_${N}gpu.txt
python - <<PY
import json
for w in ("batch","wsi"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s_${N}gpu.json"%w).read().strip().splitlines()[-1])
        print(w, "N=$N", d["value"], d["unit"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
    except Exception as e: print(w, "failed", e)
PY
