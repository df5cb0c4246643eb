#!/bin/bash
# time each variant library with the bench (no CPU arm); parity tests on the default build only
for v in "$@"; do
  export CHOMP_B200_LIB=/root/repo/tools/variants/$v.so
  timeout 200 python bench.py --no-cpu-baseline --min-seconds 0 --steps 20 --warmup 5 2>gpurun_out/bench_$v.err | tail -1 > gpurun_out/bench_$v.json
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_%s.json" % v))
    print(v, "points/s %.0f  ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(x, 3) for k, x in d["roofline"]["all_kernels_ms"].items()})
except Exception as e:
    print(v, "FAILED", e)
PY
done
