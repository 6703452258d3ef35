#!/bin/bash
# Under gpurun: one full ncu capture of the kernels matching $1 in an eager engine step (64 images, 400 windows per forward);
# $2 = launches to skip, $3 = launches to capture.  Summary + hottest source lines land in gpurun_out/.
K=$1; S=${2:-4}; C=${3:-1}; TAG=${4:-$K}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -f -o gpurun_out/prof_$TAG python tools/engine_once.py 64 2 400 > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep smsp__inst_executed_pipe > gpurun_out/ncu_${TAG}_summary.txt
python tools/ncu_lines.py gpurun_out/prof_$TAG.ncu-rep $K 0 30 > gpurun_out/ncu_${TAG}_lines.txt
head -60 gpurun_out/ncu_${TAG}_summary.txt; head -50 gpurun_out/ncu_${TAG}_lines.txt
