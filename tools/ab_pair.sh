#!/bin/bash
# Under gpurun: A/B of the fused statistics + apply launch (k_nhwc_pair) against the two-launch pair on the bench step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet.py -m gpu -x -q 2>&1 | tail -5
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$tag.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("$tag: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], {n:round(k[n]["ms_per_step"],3) for n in k if n.startswith("k_nhwc")})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/ab_$tag.err").read()[-1500:])
PY
}
for spec in "$@"; do
  tag=$(echo "$spec" | tr ' =' '__')
  run "$tag" $spec
done
