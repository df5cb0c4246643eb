#!/bin/bash
# usage (under gpurun --gpus N): tools/gpu_strong.sh N <tag>  -- strong scaling of the headline workload, 4096 points in all
N=$1; tag=$2
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --points 4096 --min-seconds 1 > gpurun_out/${tag}_strong_n$N.json 2> gpurun_out/${tag}_strong_n$N.err
echo "strong rc=$?"; tail -c 600 gpurun_out/${tag}_strong_n$N.json | cut -c1-600
