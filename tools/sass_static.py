"""Static SASS size per source line of one kernel: nvdisasm -g output parser.
usage: python tools/sass_static.py <cubin> <kernel-substring> [top]"""
import collections, re, subprocess, sys
cubin, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
cur_fn = None; cur_line = None; cnt = collections.Counter(); files = {}
infn = False
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"//-+ \.text\.(\S+)", ln)
    if m:
        infn = kern in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln) and cur_line:
        cnt[cur_line] += 1
tot = sum(cnt.values())
byf = collections.Counter()
for (f, l), v in cnt.items(): byf[f] += v
print("total", tot, "instr =", tot * 16 // 1024, "KB", dict(byf))
for (f, l), v in cnt.most_common(top): print("%5d  %s:%d" % (v, f, l))
