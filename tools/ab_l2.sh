#!/bin/bash
# Under gpurun: A/B of the traversal order of the statistics / apply pair (CELLSEG_L2_ORDER) on the bench step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for o in 0 5 6 4 7; do
  CELLSEG_L2_ORDER=$o python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/ab_l2_$o.json 2> gpurun_out/ab_l2_$o.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_l2_$o.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("order $o: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], {n:round(k[n]["ms_per_step"],3) for n in k if n.startswith("k_nhwc") or n.startswith("k_stem")})
PY
done
