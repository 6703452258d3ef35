#!/bin/bash
# Under gpurun: launch list + one full ncu capture of the tile kernels of the post-processing chain.
mkdir -p gpurun_out
W=${1:-batch}
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_pp_$W.csv python tools/pp_once.py $W 3 > gpurun_out/ncu_pp_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_filter_local|k_label_local|k_label_final|k_filter_final' -s 6 -c 6 -f -o gpurun_out/prof_pp_$W python tools/pp_once.py $W 3 >> gpurun_out/ncu_pp_$W.log 2>&1
tail -5 gpurun_out/ncu_pp_$W.log
