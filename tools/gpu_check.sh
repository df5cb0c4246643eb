#!/bin/bash
# quick GPU check: parity tests (bounded) then a bench line without the CPU arm
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 200 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print({k: round(v, 3) for k, v in d["roofline"]["all_kernels_ms"].items()})
PY
