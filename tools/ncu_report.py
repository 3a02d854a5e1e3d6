"""One text summary of an .ncu-rep for profiles/: key raw metrics, stall reasons (share of warp-state samples), the
per-function instruction budget (device functions are inlined) and the hottest source lines.
usage: python tools/ncu_report.py rep.ncu-rep "title line" > profiles/xxx.txt"""
import csv
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
print("# " + title)
print(subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), rep], stdout=subprocess.PIPE, text=True).stdout.rstrip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2]
st = {}
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(data[i].replace(",", ""))
        except ValueError:
            pass
tot = sum(st.values()) or 1.0
print("\n# stall reasons (share of warp-state samples)")
print(", ".join("%s %.1f" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:12]))
for name in ("smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
             "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum"):
    if name in hdr:
        print(name, data[hdr.index(name)])
with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
    f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                           text=True).stdout)
    path = f.name
print("\n# warp-instructions by function (device functions are inlined: per-phase budget)")
print("\n".join(subprocess.run([sys.executable, os.path.join(HERE, "ncu_funcs.py"), path], stdout=subprocess.PIPE,
                               text=True).stdout.splitlines()[:36]))
print("\n# hottest source lines")
print("\n".join(subprocess.run([sys.executable, os.path.join(HERE, "ncu_lines.py"), path, "28"], stdout=subprocess.PIPE,
                               text=True).stdout.splitlines()[:32]))
os.unlink(path)
