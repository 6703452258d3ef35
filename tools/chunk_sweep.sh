#!/bin/bash
# Under gpurun: bench.py at several windows-per-forward values (L2 residency of the activations vs grid size).
for c in "$@"; do
python bench.py --chunk $c --no_cpu_baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('chunk', $c, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'glue', d['stages']['predictor_glue_ms_per_step'])"
done
