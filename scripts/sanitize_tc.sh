#!/bin/bash
# compute-sanitizer over the tcgen05 projection kernels added in round 2 (linear_res_ln_tc, mlp_res_ln_tc) and the kernels
# touched late in the round (dwconv_pack, stem).   gpurun --timeout 1500 -- 'bash scripts/sanitize_tc.sh'
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9 --print-limit 20"
SEL='(mlp_res_ln and (1000 or 777 or 129 or 1-96)) or (linear_res_ln and (1000 or 777 or 513)) or dwconv or stem'
timeout 600 $SAN --tool memcheck python -m pytest tests/test_gpu_cross.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/r2_memcheck_tc.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_memcheck_tc.log
timeout 600 $SAN --tool racecheck python -m pytest tests/test_gpu_cross.py -m gpu -q -x -k "(mlp_res_ln and (1000 or 129)) or (linear_res_ln and 1000)" -p no:cacheprovider > gpurun_out/r2_racecheck_tc.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_racecheck_tc.log
tail -n 6 gpurun_out/r2_memcheck_tc.log; tail -n 6 gpurun_out/r2_racecheck_tc.log
