"""Summarise an ncu `--page source --csv --print-source cuda,sass` dump: per source line executed warp-instructions
and stall samples (first kernel instance only).  usage: python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys
import collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
lines = collections.OrderedDict()
seen_first_kernel = 0
kernel = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path" or r[0] == "File Name":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if kernel is None:
            kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] == "":
        continue      # sass rows
    iE = hdr.index("Instructions Executed")
    iS = hdr.index("# Samples")
    iB = hdr.index("stall_barrier")
    try:
        key = (cur_file, int(r[0]))
        val = (int(r[iE]), int(r[iS]), r[1].strip()[:90])
    except ValueError:
        continue
    if key in lines:
        continue      # later kernel instances repeat the same keys
    lines[key] = val
tot = sum(v[0] for v in lines.values())
totS = sum(v[1] for v in lines.values())
print("kernel:", kernel, " total warp-inst (first instance):", tot, " samples:", totS)
byfile = collections.Counter()
for (f, l), v in lines.items():
    byfile[f] += v[0]
print("by file:", dict(byfile))
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * v[0] / tot, 100.0 * v[1] / max(1, totS), f, l, v[2]))
