"""Aggregate an ncu cuda,sass source dump by phase (line ranges of d2d_step.cuh) + d2d_math.cuh."""
import csv, sys, collections
path = sys.argv[1]
ranges = [("scalars load/store", 100, 133), ("bulk/reset", 134, 170), ("agents", 171, 219), ("mark", 220, 226),
          ("cast_ray (march+setup)", 227, 311), ("rays loop", 312, 324), ("tracker", 325, 438), ("leader", 439, 533),
          ("obs(block)", 534, 578), ("fused block kernel", 579, 638), ("obs(warp)", 639, 666), ("fused warp kernel", 667, 760)]
rows = list(csv.reader(open(path)))
cur = None; hdr = None; seen = set(); agg = collections.Counter(); smp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    if (cur, ln) in seen: continue
    seen.add((cur, ln))
    e = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
    if cur == "d2d_step.cuh":
        name = next((n for n, a, b in ranges if a <= ln <= b), "other")
    else:
        name = cur
    agg[name] += e; smp[name] += s
tot = sum(agg.values()); ts = sum(smp.values())
for k, v in agg.most_common():
    print("%-28s %9d  %5.1f%% inst  %5.1f%% samples" % (k, v, 100.0 * v / tot, 100.0 * smp[k] / max(ts, 1)))
print("total", tot)
