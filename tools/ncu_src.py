#!/usr/bin/env python
"""Per-source-line view of one kernel from `ncu -i rep --page source --csv` (several kernels per file).
usage: ncu_src.py <src.csv> <kernel substring> [top N] [cubin]
Stall samples and executed warp instructions are aggregated per CUDA source line (nvdisasm -g line info
of the cubin extracted with `cuobjdump -xelf all`), with the top stall reasons per line."""
import collections
import csv
import re
import subprocess
import sys

csv.field_size_limit(1 << 30)
src_csv, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cubin = sys.argv[4] if len(sys.argv) > 4 else "/tmp/cub/chomp_b200.sm_100a.cubin"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line, inside, cur = {}, False, None
for ln in dis:
    if ln.startswith("//-----") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
# sections: "Kernel Name" row, header row, data rows; take the first section of this kernel whose Source column is SASS
sec, i = None, 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and kname in rows[i][1]:
        hdr = rows[i + 1]
        j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            j += 1
        body = rows[i + 2:j]
        if body and re.match(r"^\s*(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", body[0][hdr.index("Source")] or ""):
            sec = (hdr, body)
            break
        i = j
    else:
        i += 1
if sec is None:
    sys.exit("no SASS section for " + kname)
hdr, body = sec
ia, ie, ns = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(k, h[6:]) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
per, samp, ops = collections.Counter(), collections.Counter(), collections.Counter()
stalls = collections.defaultdict(collections.Counter)
tot = tots = 0
base = None
kstall = collections.Counter()
for r in body:
    if len(r) <= ie:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    off = a - base
    n = float(r[ie] or 0)
    s = float(r[ns] or 0)
    line, sass = addr2line.get(off, (("?", 0), "?"))
    per[line] += n
    samp[line] += s
    tot += n
    tots += s
    ops[sass.split()[1] if sass.startswith("@") and len(sass.split()) > 1 else sass.split()[0]] += n
    for k, nm in stall_cols:
        v = float(r[k] or 0)
        if v:
            stalls[line][nm] += v
            kstall[nm] += v
print("%s: warp instructions %.4g, samples %d" % (kname, tot, tots))
if top < 0:     # negative N: rank the lines by executed instructions instead of samples
    for line, n in per.most_common(-top):
        print("%-20s:%4d  inst %5.1f%%  samples %5.1f%%" % (line[0], line[1], 100*n/tot, 100*samp[line]/max(tots, 1)))
    top = 0
print("kernel stalls:", ", ".join("%s %.1f%%" % (k, 100*v/max(tots, 1)) for k, v in kstall.most_common(8)))
for line, s in samp.most_common(top):
    st = ", ".join("%s %.0f%%" % (k, 100*v/max(s, 1)) for k, v in stalls[line].most_common(3))
    print("%-20s:%4d  samples %5.1f%%  inst %5.1f%%   %s" % (line[0], line[1], 100*s/max(tots, 1), 100*per[line]/tot, st))
if len(sys.argv) > 5:      # lines that execute the given opcode most
    want = sys.argv[5]
    byline = collections.Counter()
    for r in body:
        if len(r) <= ie:
            continue
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        line, sass = addr2line.get(a - base, (("?", 0), "?"))
        op = sass.split()[1] if sass.startswith("@") and len(sass.split()) > 1 else sass.split()[0]
        if op.startswith(want):
            byline[line] += float(r[ie] or 0)
    print("lines executing %s:" % want, ", ".join("%s:%d %.1f%%" % (l[0], l[1], 100*n/tot) for l, n in byline.most_common(14)))
print("opcodes:", ", ".join("%s %.1f%%" % (o, 100*n/tot) for o, n in ops.most_common(16)))
