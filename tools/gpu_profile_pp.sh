#!/bin/bash
# Under gpurun: launch list + one full ncu capture of the tile kernels of the post-processing chain and the blend kernel.
mkdir -p gpurun_out
W=${1:-batch}
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_pp_$W.csv python tools/pp_once.py $W 3 > gpurun_out/ncu_pp_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile|k_final|k_label_merge' -s 5 -c 5 -f -o gpurun_out/prof_pp_$W python tools/pp_once.py $W 3 >> gpurun_out/ncu_pp_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_blend' -s 2 -c 1 -f -o gpurun_out/prof_blend python tools/blend_once.py 64 4 >> gpurun_out/ncu_pp_$W.log 2>&1
tail -5 gpurun_out/ncu_pp_$W.log
