"""Static SASS instruction count per source line for one kernel (nvdisasm -g)."""
import re, subprocess, sys, collections
kname = sys.argv[1]
dis = subprocess.run(["nvdisasm", "-g", "-c", "/tmp/cub/chomp_b200.sm_100a.cubin"], capture_output=True, text=True).stdout.splitlines()
inside = False; cur = None; cnt = collections.Counter(); files = collections.Counter()
for ln in dis:
    if ln.startswith("//-----") and ".text." in ln: inside = kname in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln): cnt[cur] += 1; files[cur[0] if cur else "?"] += 1
tot = sum(cnt.values())
print("total", tot, "instr =", tot*16/1024, "KB")
print(files.most_common(8))
for k, v in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%-24s %5d  %5.1f KB" % ("%s:%d" % k, v, v*16/1024))
