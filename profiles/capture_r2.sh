#!/bin/bash
# Round 2 captures.   gpurun --timeout 1500 -- 'bash profiles/capture_r2.sh'
set -x
mkdir -p gpurun_out
export PYTHONPATH=.
# launch list of the bench step (eager, so that every kernel is a separate launch record)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
    python bench.py --steps 2 --warmup 1 --no-microbench --no-extras --no-cpu-baseline --no-graph > gpurun_out/r2_launches_bench.log 2>&1
NCU="ncu --clock-control none --set full"
$NCU -k regex:scan_lanes -c 1 -o gpurun_out/r2_lanes_fp16 -f python profiles/scan_once.py n1_fp16 > /dev/null 2>&1
$NCU -k regex:linear_res_ln -c 1 -o gpurun_out/r2_linear_res_ln -f python profiles/step_profile.py E fp16 64 > /dev/null 2>&1
$NCU -k regex:ss2d_merge_norm -c 1 -o gpurun_out/r2_merge_norm -f python profiles/step_profile.py E fp16 64 > /dev/null 2>&1
$NCU -k regex:mlp_res_ln -s 1 -c 1 -o gpurun_out/r2_mlp_res_ln -f python profiles/mlp_once.py > /dev/null 2>&1
$NCU -k regex:dwconv_pack -s 1 -c 1 -o gpurun_out/r2_dwconv_pack -f python profiles/ss2d_once.py > /dev/null 2>&1
for r in gpurun_out/r2_lanes_fp16 gpurun_out/r2_linear_res_ln gpurun_out/r2_merge_norm gpurun_out/r2_mlp_res_ln gpurun_out/r2_dwconv_pack; do
    ncu -i $r.ncu-rep --page raw --csv > $r.raw.csv
done
ls -la gpurun_out | tail -12
