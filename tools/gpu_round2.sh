#!/bin/bash
# usage (under gpurun): tools/gpu_round2.sh <tag>  -- GPU tests, headline bench, config-5 bench, config-5 ncu capture
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/${tag}_pytest.log | tail -12
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --workload covariance > gpurun_out/${tag}_bench_cov.json 2> gpurun_out/${tag}_bench_cov.err; echo "bench cov rc=$?"; tail -5 gpurun_out/${tag}_bench_cov.err
timeout 900 bash tools/gpu_profile_cov.sh $tag
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["step_ms"])
print({k: round(v, 3) for k, v in d["roofline"]["all_kernels_ms"].items()})
try:
    d = json.loads(open("gpurun_out/${tag}_bench_cov.json").read().strip().splitlines()[-1])
    print("cov points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    print({k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
    print(d.get("cpu_baseline"))
except Exception as e:
    print("cov bench unreadable", e)
PY
