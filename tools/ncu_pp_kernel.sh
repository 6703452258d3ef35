#!/bin/bash
# Under gpurun: one full ncu capture of kernel $1 of the post-processing chain on dense blob masks (64 x 1024^2) + hottest lines.
K=$1; S=${2:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -f -o gpurun_out/prof_pp_$K python tools/pp_once.py batch 2 > gpurun_out/ncu_pp_$K.log 2>&1
tail -1 gpurun_out/ncu_pp_$K.log
python tools/ncu_summary.py gpurun_out/prof_pp_$K.ncu-rep smsp__inst_executed_pipe_lsu l1tex__data_pipe_lsu_wavefronts_mem_shared.sum > gpurun_out/ncu_pp_${K}_summary.txt
python tools/ncu_lines.py gpurun_out/prof_pp_$K.ncu-rep $K 0 40 > gpurun_out/ncu_pp_${K}_lines.txt
cat gpurun_out/ncu_pp_${K}_summary.txt | head -40; cat gpurun_out/ncu_pp_${K}_lines.txt
