#!/bin/bash
# usage (under gpurun): tools/gpu_round3.sh <tag> [variants...]  -- GPU tests, headline bench, variant benches
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/${tag}_pytest.log | tail -8
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: v for k, v in d["step_ms"].items() if k != "all"})
print({k: round(v, 3) for k, v in d["roofline"]["all_kernels_ms"].items()}, "sustained", d.get("sustained", {}).get("ms_per_step"))
PY
bash tools/gpu_variants.sh "$@"
