"""Group SASS instructions by execution count classes: share of instructions and of stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
for hi, r in enumerate(rows):
    if "Instructions Executed" in r: break
hdr = rows[hi]
ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
cls = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = tots = 0
for r in rows[hi+1:]:
    if len(r) <= ie: continue
    n = float(r[ie] or 0); s = float(r[ns] or 0)
    key = float("%.2g" % n)
    cls[key][0] += 1; cls[key][1] += n; cls[key][2] += s
    tot += n; tots += s
for k in sorted(cls, key=lambda k: -cls[k][1])[:25]:
    c = cls[k]
    print("exec ~%-9.3g  %5d instr  inst %5.1f%%  samples %5.1f%%" % (k, c[0], 100*c[1]/tot, 100*c[2]/tots))
