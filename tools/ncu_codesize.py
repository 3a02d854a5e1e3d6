"""Static SASS size per source line / file from an ncu cuda,sass source dump (first kernel instance)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur = None; hdr = None; key = None; seen = set(); cnt = collections.Counter(); src = {}
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] != "":
        k = (cur, r[0])
        key = None if k in seen else k
        if key: seen.add(k); src[k] = r[1].strip()[:80]
        continue
    if key and r[2].startswith("0x"): cnt[key] += 1
tot = sum(cnt.values())
byfile = collections.Counter()
for (f, l), v in cnt.items(): byfile[f] += v
print("total static SASS:", tot, "=", tot * 16 // 1024, "KB;", dict(byfile))
for (f, l), v in cnt.most_common(top): print("%5d  %s:%s  %s" % (v, f, l, src[(f, l)]))
