#!/bin/bash
# usage (under gpurun): tools/gpu_launchlist.sh <tag> [lib]  -- per-launch durations of one bench run (serialised, cold)
tag=$1
[ -n "$2" ] && export CHOMP_B200_LIB=$2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --min-seconds 0 > gpurun_out/${tag}_launch.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${tag}_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr+1:]:
    if len(r) > mv:
        try: agg[r[kn].split("(")[0]].append(float(r[mv].replace(",", "")))
        except ValueError: pass
for k, v in agg.items():
    print("%-40s n=%3d  last %.1f us  median %.1f us" % (k[:40], len(v), v[-1]/1e3, sorted(v)[len(v)//2]/1e3))
PY
