"""Generates csrc/d2d_tan_table.inc and csrc/d2d_sincos_table.inc (double-double tan/sin/cos of j/32, j=0..25)
with 80-digit mpmath arithmetic.  Build-time tool; the generated .inc files are committed."""
import os
import mpmath as mp

mp.mp.dps = 80
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym_drone2d_activeperception_b200", "csrc")


def dd(x):
    h = float(x)
    l = float(x - mp.mpf(h))
    return h, l


def fmt(v):
    return "%.17e" % v


with open(os.path.join(OUT, "d2d_tan_table.inc"), "w") as f:
    for j in range(26):
        h, l = dd(mp.tan(mp.mpf(j) / 32))
        f.write("    {%s, %s},\n" % (fmt(h), fmt(l)))
with open(os.path.join(OUT, "d2d_sincos_table.inc"), "w") as f:
    for j in range(26):
        sh, sl = dd(mp.sin(mp.mpf(j) / 32))
        ch, cl = dd(mp.cos(mp.mpf(j) / 32))
        f.write("    {%s, %s, %s, %s},\n" % (fmt(sh), fmt(sl), fmt(ch), fmt(cl)))
# constants used in d2d_math.cuh, printed for review
for name, val in [("1/3", mp.mpf(1) / 3), ("1/6", mp.mpf(1) / 6), ("1/24", mp.mpf(1) / 24), ("pi/2", mp.pi / 2)]:
    print(name, ["%.20e" % v for v in dd(val)])
p1 = 1.57079632673412561417e+00
p2 = 6.07710050630396597660e-11
p3 = 2.02226624871116645580e-21
print("pio2 tail", "%.20e" % float(mp.pi / 2 - mp.mpf(p1) - mp.mpf(p2) - mp.mpf(p3)))
for name, val in [("2/15", mp.mpf(2) / 15), ("17/315", mp.mpf(17) / 315), ("62/2835", mp.mpf(62) / 2835),
                  ("1382/155925", mp.mpf(1382) / 155925), ("21844/6081075", mp.mpf(21844) / 6081075),
                  ("929569/638512875", mp.mpf(929569) / 638512875)]:
    print(name, "%.20e" % float(val))
