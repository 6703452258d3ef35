#!/bin/bash
# Under gpurun (1 GPU): the round's evidence in one call — GPU test tier, smoke, bench lines (batch / sweep / wsi / metric), then the
# profile pass of tools/gpu_profile_step.sh.  Everything lands in gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_batch_1gpu.json 2> gpurun_out/bench_batch.err; tail -c 400 gpurun_out/bench_batch.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/bench_batch.err
python bench.py --workload sweep --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/${TAG}_bench_sweep_1gpu.json 2> gpurun_out/bench_sweep.err; tail -c 300 gpurun_out/bench_sweep.err
python bench.py --workload wsi --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/${TAG}_bench_wsi_1gpu.json 2> gpurun_out/bench_wsi.err; tail -c 300 gpurun_out/bench_wsi.err
python tools/metric_bench.py 4096 --cpu > gpurun_out/${TAG}_metric_config4.json 2> gpurun_out/metric.err; tail -c 300 gpurun_out/metric.err
python - <<PY
import json
for f in ("batch_1gpu","sweep_1gpu","wsi_1gpu"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d.get("roofline",{}).get("kernel"), d.get("roofline",{}).get("frac"))
    except Exception as e: print(f, "failed", e)
print(open("gpurun_out/${TAG}_metric_config4.json").read()[-600:])
PY
bash tools/gpu_profile_step.sh $TAG > gpurun_out/profile_step.log 2>&1; tail -5 gpurun_out/profile_step.log
