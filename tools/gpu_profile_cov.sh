#!/bin/bash
# usage (under gpurun): tools/gpu_profile_cov.sh <tag>   -- ncu --set full of the config-5 kernels (one covariance call, 128 points)
tag=${1:-r2}
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum
CMD="python bench.py --workload covariance --steps 1 --warmup 1 --points 128 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${tag}_cov_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --metrics $M \
    -k regex:'cov_|tri_' -s 9 -c 9 -f \
    -o gpurun_out/${tag}_cov_prof $CMD > gpurun_out/${tag}_cov_ncu.log 2>&1
echo "cov profile rc=$?"
