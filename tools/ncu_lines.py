"""Aggregate an ncu SASS source page per CUDA source line, using nvdisasm -g line info.
usage: ncu_lines.py <src.csv from ncu --page source --csv> <kernel mangled-name substring> [cubin]"""
import csv, re, subprocess, sys, collections
src_csv, kname = sys.argv[1], sys.argv[2]
cubin = sys.argv[3] if len(sys.argv) > 3 else "/tmp/cub/chomp_b200.sm_100a.cubin"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line = {}
inside = False; cur = None
for ln in dis:
    if ln.startswith("//-----") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
for hi, r in enumerate(rows):
    if "Instructions Executed" in r: break
hdr = rows[hi]
ia, ie, ns = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
per = collections.Counter(); samp = collections.Counter(); ops = collections.Counter()
tot = 0; tots = 0
for r in rows[hi+1:]:
    if r and r[0] == "Kernel Name": break          # a second capture of the same kernel: first one only
    if len(r) <= ie: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    off = a - base
    n = float(r[ie] or 0); s = float(r[ns] or 0)
    line, sass = addr2line.get(off, (("?", 0), "?"))
    per[line] += n; samp[line] += s; tot += n; tots += s
    ops[sass.split()[0] if not sass.startswith("@") else sass.split()[1]] += n
print("total warp instructions %.4g, samples %d" % (tot, tots))
for line, n in per.most_common(45):
    print("%-22s:%4d  inst %5.1f%%  samples %5.1f%%" % (line[0], line[1], 100*n/tot, 100*samp[line]/max(tots, 1)))
print("opcodes:", ", ".join("%s %.1f%%" % (o, 100*n/tot) for o, n in ops.most_common(14)))
