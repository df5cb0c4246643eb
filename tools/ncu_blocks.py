"""Basic-block style view: consecutive SASS instructions with the same executed count."""
import csv, re, subprocess, sys
src_csv, kname = sys.argv[1], sys.argv[2]
cubin = "/tmp/cub/chomp_b200.sm_100a.cubin"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line = {}; inside = False; cur = None
for ln in dis:
    if ln.startswith("//-----") and ".text." in ln: inside = kname in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = "%s:%s" % (m.group(1).split("/")[-1].replace(".cuh", "").replace("_intrinsics.hpp", ""), m.group(2)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
for hi, r in enumerate(rows):
    if "Instructions Executed" in r: break
hdr = rows[hi]
ia, ie, ns = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; blocks = []
for r in rows[hi+1:]:
    if len(r) <= ie: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    off = a - base; n = float(r[ie] or 0); s = float(r[ns] or 0)
    line, sass = addr2line.get(off, ("?", "?"))
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    if blocks and blocks[-1]["n"] == n: b = blocks[-1]
    else: b = dict(n=n, cnt=0, samp=0, lines=[], ops={}, off=off); blocks.append(b)
    b["cnt"] += 1; b["samp"] += s
    if not b["lines"] or b["lines"][-1] != line: b["lines"].append(line)
    b["ops"][op.split(".")[0]] = b["ops"].get(op.split(".")[0], 0) + 1
tot = sum(b["n"]*b["cnt"] for b in blocks); tots = sum(b["samp"] for b in blocks)
print("total %.4g" % tot)
for b in blocks:
    share = 100*b["n"]*b["cnt"]/tot
    if share < float(sys.argv[3] if len(sys.argv) > 3 else 0.4): continue
    ops = sorted(b["ops"].items(), key=lambda kv: -kv[1])[:5]
    print("off %5x  exec %.3g x %3d instr = %5.1f%% (samples %4.1f%%)  %s | %s" % (b["off"], b["n"], b["cnt"], share, 100*b["samp"]/tots, " ".join("%s%d" % kv for kv in ops), ",".join(b["lines"][:6])))
