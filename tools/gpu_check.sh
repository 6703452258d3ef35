#!/bin/bash
# Run on the GPU box (via gpurun): smoke, GPU test tier, quick timings, one racecheck pass.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/pp_bench.py 2>&1 | tail -12
if [ "$1" == "race" ]; then
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python - <<'PY' 2>&1 | tail -15
import sys; sys.path.insert(0, ".")
import torch, numpy as np
from cellseg_b200 import postprocess, synth
m = torch.from_numpy(synth.noise_mask(200, 300, 0.55, 1)).cuda()
postprocess.postprocess_device(m); torch.cuda.synchronize(); print("race run done")
PY
fi
