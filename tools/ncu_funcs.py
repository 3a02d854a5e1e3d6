"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by enclosing function (device functions are
inlined, so this is the per-phase instruction / stall-sample budget of a kernel).  usage: ncu_funcs.py dump.csv"""
import collections
import csv
import os
import re
import sys

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_drone2d_activeperception_b200", "csrc")
_starts = {}


def func_of(fname, ln):
    if fname not in _starts:
        st = []
        path = os.path.join(CSRC, fname)
        if os.path.isfile(path):
            for i, l in enumerate(open(path), 1):
                m = re.match(r"^(?:template.*\n)?(?:static |D2D_HD |__device__ |__global__ |__host__ )+.*?(\w+)\s*\(", l)
                if m and not l.startswith(" "):
                    st.append((i, m.group(1)))
        _starts[fname] = st
    name = fname
    for i, n in _starts[fname]:
        if i <= ln:
            name = fname + ":" + n
        else:
            break
    return name


rows = list(csv.reader(open(sys.argv[1])))
cur = None
hdr = None
seen = set()
agg = collections.Counter()
smp = collections.Counter()
kernels = 0
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    if (cur, ln) in seen:
        continue
    seen.add((cur, ln))
    e = int(r[hdr.index("Instructions Executed")])
    s = int(r[hdr.index("# Samples")])
    k = func_of(cur, ln)
    agg[k] += e
    smp[k] += s
tot = sum(agg.values())
ts = sum(smp.values())
for k, v in agg.most_common(40):
    print("%-52s %9d  %5.1f%% inst  %5.1f%% samples" % (k, v, 100.0 * v / tot, 100.0 * smp[k] / max(ts, 1)))
print("total", tot)
