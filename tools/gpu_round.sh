#!/bin/bash
# usage (under gpurun): tools/gpu_round.sh <tag>   -- GPU tests, bench line, launch list, full ncu capture
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0 > gpurun_out/${tag}_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 bash tools/gpu_profile.sh $tag 4096
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print({k: round(v, 3) for k, v in d["roofline"]["all_kernels_ms"].items()})
PY
