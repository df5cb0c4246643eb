#!/bin/bash
# usage (under gpurun --gpus N): tools/gpu_scaling.sh N <tag>  -- headline bench weak + strong, config 5 strong, NCCL tests
N=$1; tag=$2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -3
run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_weak_n$N.json 2> gpurun_out/${tag}_weak_n$N.err; echo "weak rc=$?"
run bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --points 4096 > gpurun_out/${tag}_strong_n$N.json 2> gpurun_out/${tag}_strong_n$N.err; echo "strong rc=$?"
run bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --points 32768 --min-seconds 0 > gpurun_out/${tag}_strong32k_n$N.json 2> gpurun_out/${tag}_strong32k_n$N.err; echo "strong 32k rc=$?"
run bench.py --gpus $N --workload covariance --scaling strong --points 4096 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_cov_strong_n$N.json 2> gpurun_out/${tag}_cov_strong_n$N.err; echo "cov strong rc=$?"
python - <<PY
import json
for name in ("weak", "strong", "strong32k", "cov_strong"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s_n$N.json" % name).read().strip().splitlines()[-1])
        print(name, "n_gpus", d["n_gpus"], "points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"].get("points_per_gpu"), d["config"].get("collective"))
    except Exception as e:
        print(name, "unreadable", e)
PY
