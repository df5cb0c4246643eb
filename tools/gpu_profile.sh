#!/bin/bash
# usage (under gpurun): tools/gpu_profile.sh <tag> [points]
# plain run first (must exit 0), then one `ncu --set full` capture of all six kernels of one timed
# step, with the executed-FP64 instruction counters SURVEY 8(d) asks for.  Read the report back with
# tools/ncu_kernels.py (writes profiles/ncu_<tag>_kernels.{json,txt}).
tag=${1:-r2}
pts=${2:-4096}
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum
CMD="python bench.py --steps 2 --warmup 3 --points $pts --no-cpu-baseline --no-e2e --min-seconds 0"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --metrics $M \
    -k regex:'limber_tables|mass_tables|nu_nodes|halo_sums|halo_splines|wtheta' -s 18 -c 6 -f \
    -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "profile rc=$?"
