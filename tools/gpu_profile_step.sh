#!/bin/bash
# Under gpurun: evidence for profiles/ — (1) launch list of the bench's timed region (graph replays included: ncu serialises the
# graph's kernel nodes), (2) full ncu captures of every libcellseg kernel of one eager step at the bench's shapes,
# (3) the post-processing chain alone on cell-like masks.   Usage: tools/gpu_profile_step.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
CS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench_batch_graph.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
python tools/launch_shares.py gpurun_out/${TAG}_launches_bench_batch_graph.csv 40 > gpurun_out/${TAG}_launch_shares_bench_batch.md
# one eager call on 64 images = 4 forwards of 400 windows; the first forward's glue kernels of the 2nd call
ncu --set full --clock-control none --import-source on -k regex:"k_stem|k_nhwc|k_head" -s 152 -c 38 -f -o gpurun_out/prof_step_glue \
    python tools/engine_once.py 64 2 400 > gpurun_out/ncu_glue2.log 2>&1
tail -1 gpurun_out/ncu_glue2.log
python tools/ncu_summary.py gpurun_out/prof_step_glue.ncu-rep > gpurun_out/${TAG}_ncu_step_glue.txt; rm -f gpurun_out/prof_step_glue.ncu-rep
# blend + the whole post chain of the 2nd call (predictor masks)
ncu --set full --clock-control none --import-source on -k regex:"k_blend|k_certify|k_tile|k_final|k_u8_max|k_filter|k_label|k_scan" -s 14 -c 14 -f -o gpurun_out/prof_step_post \
    python tools/engine_once.py 64 2 400 > gpurun_out/ncu_post2.log 2>&1
tail -1 gpurun_out/ncu_post2.log
python tools/ncu_summary.py gpurun_out/prof_step_post.ncu-rep > gpurun_out/${TAG}_ncu_step_post.txt
for k in k_blend k_tile k_certify k_final; do python tools/ncu_lines.py gpurun_out/prof_step_post.ncu-rep $k 0 25 > gpurun_out/${TAG}_ncu_lines_$k.txt; done
rm -f gpurun_out/prof_step_post.ncu-rep
# the chain alone on cell-like masks (dense blobs)
ncu --set full --clock-control none --import-source on -k regex:"k_certify|k_tile|k_final|k_filter|k_label|k_scan" -s 12 -c 12 -f -o gpurun_out/prof_pp_blob \
    python tools/pp_once.py batch 2 > gpurun_out/ncu_pp_blob.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_pp_blob.ncu-rep > gpurun_out/${TAG}_ncu_postprocess_blob_masks.txt; rm -f gpurun_out/prof_pp_blob.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_postprocess_blob_masks.csv python tools/pp_once.py batch 3 > /dev/null 2>&1
ls -la gpurun_out | tail -20
