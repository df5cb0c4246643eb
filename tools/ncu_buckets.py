"""Bucket an ncu source page by SASS address ranges of the functions in the cubin (a kernel and the
__noinline__ device functions it calls) and by source file: where do the warp instructions and the stall
samples go?   usage: ncu_buckets.py <src.csv> <kernel substring> [cubin]"""
import csv, re, subprocess, sys, collections
src_csv, kname = sys.argv[1], sys.argv[2]
cubin = sys.argv[3] if len(sys.argv) > 3 else "/tmp/cub/chomp_b200.sm_100a.cubin"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2 = {}
inside = False; cur = None; func = None
for ln in dis:
    if ln.startswith("//-----") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside: continue
    m = re.match(r'\s*(\S+):\s*$', ln)
    if m and not m.group(1).startswith('.L'): func = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)), (m.group(3) or "").split("/")[-1], int(m.group(4) or 0)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: addr2[int(m.group(1), 16)] = (cur, m.group(2).strip(), func)
rows = list(csv.reader(open(src_csv)))
for hi, r in enumerate(rows):
    if "Instructions Executed" in r: break
hdr = rows[hi]
ia, ie, ns = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
by_line = collections.defaultdict(lambda: [0.0, 0.0]); tot = [0.0, 0.0]
stalls = collections.defaultdict(collections.Counter)
for r in rows[hi+1:]:
    if r and r[0] == "Kernel Name": break
    if len(r) <= ie: continue
    a = int(r[ia], 16)
    if base is None: base = a
    cur, sass, func = addr2.get(a - base, (None, "?", None))
    n = float(r[ie] or 0); s = float(r[ns] or 0)
    key = cur[:2] if cur else ("?", 0)
    by_line[key][0] += n; by_line[key][1] += s; tot[0] += n; tot[1] += s
    for i, h in stall_cols:
        stalls[key][h] += float(r[i] or 0)
ranges = [tuple(x.split(":")) for x in sys.argv[4:]]   # name:file:lo:hi
print("total warp inst %.4g  samples %d" % (tot[0], tot[1]))
if ranges:
    agg = collections.defaultdict(lambda: [0.0, 0.0]); st = collections.defaultdict(collections.Counter)
    for (f, l), (n, s) in by_line.items():
        name = "other"
        for rn, rf, lo, hi_ in ranges:
            if f == rf and int(lo) <= l <= int(hi_): name = rn; break
        agg[name][0] += n; agg[name][1] += s
        st[name].update(stalls[(f, l)])
    for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        top = ", ".join("%s %.0f%%" % (h[6:], 100*v/max(s, 1)) for h, v in st[name].most_common(4))
        print("%-18s inst %5.1f%%  samples %5.1f%%   %s" % (name, 100*n/tot[0], 100*s/tot[1], top))
