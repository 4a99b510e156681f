#!/bin/bash
# Round 1, session 3 (second batch): `--set full` counters of the matcher after the 3xFP16 change and of the stage-0 launches
# of the encoder's remaining big kernels.   gpurun --timeout 1200 -- 'bash profiles/capture_r1_s3b.sh'
set -x
NCU="ncu --clock-control none --set full"
$NCU -k regex:nn_argmin -c 1 -o gpurun_out/r1_s3b_match python profiles/tc_once.py > /dev/null 2>&1
for k in add_layer_norm ss2d_merge_norm ss2d_dwconv_pack linear_act_tc; do
  $NCU -k regex:$k -c 1 -o gpurun_out/r1_s3b_$k python profiles/step_profile.py E fp16 64 > /dev/null 2>&1
done
for r in gpurun_out/r1_s3b_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv; rm $r; done
ls -la gpurun_out
