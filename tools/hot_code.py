#!/usr/bin/env python
"""Instruction-cache view of a kernel: joins the per-SASS-instruction execution counts of an .ncu-rep (source page) with
nvdisasm's line info and reports the STATIC size of the code that is executed often, by enclosing source function, plus
where the warp-state samples of the `no_instruction` stall fall.  (The L1.5 instruction cache of an SM holds ~32 KB; a
kernel whose de-phased warps walk more than that per step starves on instruction fetch.)

    python tools/hot_code.py rep.ncu-rep kernel_mangled_substring units   # units = env-steps the launch processed
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)


def func_table():
    csrc = os.path.join(ROOT, "gym_drone2d_activeperception_b200", "csrc")
    tab = {}
    for f in os.listdir(csrc):
        st = []
        for i, l in enumerate(open(os.path.join(csrc, f), errors="replace"), 1):
            m = re.match(r"^(?:static |D2D_HD |__device__ |__global__ |__host__ |template )+.*?(\w+)\s*\(", l)
            if m and not l.startswith(" "):
                st.append((i, m.group(1)))
        tab[f] = st
    return tab


def func_of(tab, fname, ln):
    name = fname
    for i, n in tab.get(fname, []):
        if i <= ln:
            name = n
        else:
            break
    return name


def main():
    rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
    so = os.path.join(ROOT, "gym_drone2d_activeperception_b200", "libdrone2d.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    # instruction offset -> (file, line) for the kernel's section
    loc = {}
    inside = False
    cur = ("?", 0)
    for l in dis:
        if l.startswith("//-") and ".text." in l:
            inside = kern in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ia, ie, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ino = [i for i, h in enumerate(hdr) if "no_inst" in h.lower() or "No Instruction" in h]
    data = rows[2:]
    base = int(data[0][ia], 16)
    tab = func_table()
    stat_hot, stat_all, dyn = collections.Counter(), collections.Counter(), collections.Counter()
    smp = collections.Counter()
    n_hot = n_exec = 0
    for r in data:
        off = int(r[ia], 16) - base
        e = int(r[ie])
        (f, ln), _ = loc.get(off, (("?", 0), ""))
        fn = func_of(tab, f, ln)
        stat_all[fn] += 1
        dyn[fn] += e
        smp[fn] += int(r[ismp])
        if e > 0:
            n_exec += 1
        if e > 0.25 * units:
            stat_hot[fn] += 1
            n_hot += 1
    print("kernel %s: %d SASS instructions (%.1f KB); executed at all: %d (%.1f KB); executed > 0.25 x per unit: %d (%.1f KB)" %
          (kern, len(data), len(data) / 64.0, n_exec, n_exec / 64.0, n_hot, n_hot / 64.0))
    print("%-34s %10s %10s %12s %9s" % ("function", "hot instr", "hot KB", "dyn / unit", "samples %"))
    ts = sum(smp.values()) or 1
    for fn, n in stat_hot.most_common(40):
        print("%-34s %10d %10.2f %12.1f %9.1f" % (fn, n, n / 64.0, dyn[fn] / units, 100.0 * smp[fn] / ts))


if __name__ == "__main__":
    main()
