"""Per source line: share of one stall reason's samples (column name substring)."""
import csv, re, subprocess, sys, collections
src_csv, kname, col = sys.argv[1], sys.argv[2], sys.argv[3]
dis = subprocess.run(["nvdisasm", "-g", "-c", "/tmp/cub/chomp_b200.sm_100a.cubin"], capture_output=True, text=True).stdout.splitlines()
addr2line = {}; inside = False; cur = None
for ln in dis:
    if ln.startswith("//-----") and ".text." in ln: inside = kname in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
for hi, r in enumerate(rows):
    if "Instructions Executed" in r: break
hdr = rows[hi]
ia = hdr.index("Address"); ic = [i for i, h in enumerate(hdr) if h == col][0]; ns = hdr.index("# Samples")
base = None; per = collections.Counter(); tot = 0; alls = 0
for r in rows[hi+1:]:
    if len(r) <= ic: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    v = float(r[ic] or 0); alls += float(r[ns] or 0)
    line, sass = addr2line.get(a - base, ("?", "?"))
    per[(line, sass[:50])] += v; tot += v
print(col, "total", tot, "of", alls, "samples")
for k, v in per.most_common(25): print("%5.1f%%  %-24s %s" % (100*v/tot, k[0], k[1]))
