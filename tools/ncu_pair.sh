#!/bin/bash
# Under gpurun: full ncu captures of the fused statistics + apply launches of one eager forward (400 windows).
mkdir -p gpurun_out
CELLSEG_NORM_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:"k_nhwc_pair" -s 15 -c 8 -f -o gpurun_out/prof_pair \
    python tools/engine_once.py 64 2 400 > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
python tools/ncu_summary.py gpurun_out/prof_pair.ncu-rep lts__t_sectors_srcunit_tex_op_read lts__t_sector_op_read_hit_rate > gpurun_out/ncu_pair_summary.txt
cat gpurun_out/ncu_pair_summary.txt | head -150
