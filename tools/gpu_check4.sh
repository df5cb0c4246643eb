#!/bin/bash
tag=$1
timeout 300 python -m pytest tests/test_gpu_covariance.py tests/test_gpu_covariance_named.py tests/test_gpu_cross_covariance.py -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log | cut -c1-300
timeout 100 python bench.py --workload covariance --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_cov.json 2> gpurun_out/${tag}_bench_cov.err; echo "cov rc=$?"
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err | cut -c1-300
python - <<PY
import json
def last(p): return json.loads(open(p).read().strip().splitlines()[-1])
d = last("gpurun_out/${tag}_bench.json")
print("points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["step_ms"]["min"], d["step_ms"]["max"])
d = last("gpurun_out/${tag}_bench_cov.json")
print("cov points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items() if v["ms"] > 0.5})
PY
