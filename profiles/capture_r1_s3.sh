#!/bin/bash
# Round 1, session 3 captures (run on the GPU box through gpurun, from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/capture_r1_s3.sh'
# 1. launch list (device time per launch) of one short bench run, eager launches so every kernel is listed by name
# 2. `--set full` counters of the kernels that changed this session (no source import: the reports must stay small)
set -x
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r1_s3_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-microbench --no-cpu-baseline --no-graph > gpurun_out/r1_s3_bench_under_ncu.log 2>&1
$NCU --set full -k regex:scan_lanes -s 2 -c 1 -o gpurun_out/r1_s3_lanes_fp16 python profiles/scan_once.py n1_fp16 > /dev/null 2>&1
XP_LANES_DT_C16=0 $NCU --set full -k regex:scan_lanes -s 2 -c 1 -o gpurun_out/r1_s3_lanes_fused_dt python profiles/scan_dt_once.py > /dev/null 2>&1
$NCU --set full -k regex:nms_ -c 4 -o gpurun_out/r1_s3_nms python profiles/nms_once.py > /dev/null 2>&1
$NCU --set full -k regex:"patch_embed|dt_proj|encoder_tail|l2_normalize_cl|detector_post_cl" -c 7 -o gpurun_out/r1_s3_glue python profiles/step_profile.py E fp16 64 > /dev/null 2>&1
# gpurun copies at most 64 MiB back: keep the raw-counter CSV of every report (what profiles/summarize_ncu.py reads), drop the reports
for r in gpurun_out/r1_s3_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv; rm $r; done
ls -la gpurun_out
