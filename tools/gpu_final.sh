#!/bin/bash
# usage (under gpurun): tools/gpu_final.sh <tag>  -- the round's evidence in one call: GPU tests, headline and config-5
# bench lines (with the reference CPU arm), fast/slow split, batch sweep, launch list, full ncu captures, sanitizer
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
timeout 300 python bench.py --workload covariance --steps 5 --warmup 2 > gpurun_out/${tag}_bench_cov.json 2> gpurun_out/${tag}_bench_cov.err; echo "bench cov rc=$?"
timeout 120 python bench.py --grouped 64 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_grouped.json 2> gpurun_out/${tag}_bench_grouped.err; echo "grouped rc=$?"
timeout 200 python bench.py --sweep --sweep-min 8 --sweep-max 18 > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; echo "sweep rc=$?"
timeout 200 bash tools/gpu_launchlist.sh $tag > gpurun_out/${tag}_launchlist.txt 2>&1; echo "launch list rc=$?"
timeout 400 bash tools/gpu_profile.sh $tag 4096
timeout 400 bash tools/gpu_profile_cov.sh $tag
( timeout 280 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15;
  timeout 280 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 ) > gpurun_out/${tag}_sanitizer.txt 2>&1; echo "sanitizer rc=$?"
tail -4 gpurun_out/${tag}_sanitizer.txt
python - <<PY
import json
def last(p):
    return json.loads(open(p).read().strip().splitlines()[-1])
try:
    d = last("gpurun_out/${tag}_bench.json")
    print("points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["cpu_baseline"]["value"])
    print({k: round(v, 3) for k, v in d["roofline"]["all_kernels_ms"].items()})
    d = last("gpurun_out/${tag}_bench_cov.json")
    print("cov points/s %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("cpu_baseline", {}).get("value"))
    d = last("gpurun_out/${tag}_bench_grouped.json")
    print("grouped %.0f vs ungrouped %.0f  identical %s" % (d["value"], d["ungrouped_same_batch"]["value"], d["bit_identical_to_ungrouped"]))
    d = last("gpurun_out/${tag}_sweep.json")
    print([(r["points"], round(r["points_per_s"])) for r in d["sweep"]])
except Exception as e:
    print("summary failed", e)
PY
