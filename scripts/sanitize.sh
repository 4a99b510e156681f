#!/bin/bash
# compute-sanitizer over a subset of the GPU suite.  Run on the GPU box:
#   gpurun --timeout 2400 -- 'bash scripts/sanitize.sh'
# memcheck: out-of-bounds / misaligned accesses of every kernel family; racecheck: shared-memory hazards of the
# warp-specialised mbarrier pipelines (scan_lanes / scan_rows / ss2d_core / linear_act_tc / nn_argmin_tc).
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9 --print-limit 20"
SEL_MEM='ss2d_core or (test_gpu_scan and (golden or n16 or n1)) or test_gpu_tail or test_metrics or xpoint_tiny or ss2d_block_golden'
SEL_RACE='(ss2d_core and 16-20) or (ss2d_core and 8-8) or scan_n16_k4 or scan_n1_k4 or get_matches_any or ss2d_block_golden'
timeout 1100 $SAN --tool memcheck python -m pytest tests -m gpu -q -x -k "$SEL_MEM" -p no:cacheprovider > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
timeout 900 $SAN --tool racecheck python -m pytest tests -m gpu -q -x -k "$SEL_RACE" -p no:cacheprovider > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_racecheck.log
tail -n 5 gpurun_out/r2_memcheck.log; tail -n 5 gpurun_out/r2_racecheck.log
