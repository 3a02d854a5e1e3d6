// d2d_math.cuh -- scalar fp64 helpers of the drone2d hot path (device code; also host-compilable so the
// `-m "not gpu"` tests can check the exact same arithmetic against glibc / CPython on the CPU).
//
// Everything here is built from IEEE-754 basic operations (+ - * / sqrt fma) only, so the results are
// identical on sm_100a and on the host.  The translation unit MUST be compiled with FMA contraction off
// (nvcc -fmad=false, gcc -ffp-contract=off): the reference (CPython / NumPy element-wise code) rounds every
// multiply and add separately; fused operations appear only where written (`D2D_FMA`).
//
// What each helper restates (file:line under the reference tree):
//   d2d_cell()        int(x // 10)                  CPython float_divmod      utils.py:655-656, 548, 781
//   d2d_pymod()       x % w (w > 0)                 CPython float_rem         utils.py:614, 743
//   d2d_norm2()       np.linalg.norm([x, y])        sqrt(fma(y,y,x*x))        utils.py:476,756,774 ...
//   d2d_tan()         math.tan                      glibc tan (<1 ulp, not correctly rounded); here: double-double
//                                                   evaluation rounded to nearest                 utils.py:640
//   d2d_sincos()      math.cos / math.sin           same approach             yaw_planner.py:72
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define D2D_HD __host__ __device__ __forceinline__
#define D2D_HD_NOINLINE __host__ __device__ __noinline__     // one shared copy: keeps the kernels inside the I-cache
#else
#define D2D_HD static inline
#define D2D_HD_NOINLINE static
#endif

#if defined(__CUDA_ARCH__)
#define D2D_FMA(a, b, c) __fma_rn((a), (b), (c))
#define D2D_SQRT(a) __dsqrt_rn((a))
#define D2D_RINT(a) rint((a))
#else
#define D2D_FMA(a, b, c) fma((a), (b), (c))
#define D2D_SQRT(a) sqrt((a))
#define D2D_RINT(a) rint((a))
#endif

#define D2D_PI 3.141592653589793
#define D2D_TWO_PI 6.283185307179586
#define D2D_DEG2RAD (D2D_PI / 180.0)

// ------------------------------------------------------------------------------------------------ integer cell
// int(x // scale) for 0 <= x < scale * 2^20.  CPython computes floor of the EXACT quotient here (fmod is exact and
// x - fmod(x, w) is an exact multiple of w), so any exact floor is equivalent; this one needs no fmod.
D2D_HD int d2d_cell(double x, double scale, double inv_scale) {
    int k = (int)(x * inv_scale);
    if (x < scale * (double)k) k -= 1;              // products of small integers are exact
    else if (x >= scale * (double)(k + 1)) k += 1;
    return k;
}

// Python `x % w` for w > 0 (float_rem): fmod then sign fix-up.  fmod is exact, and for |x| < 2w it is |x| or |x| - w
// (an exact subtraction, Sterbenz), so the library fmod is only needed outside that range.
D2D_HD_NOINLINE double d2d_fmod_slow(double ax, double w) { return fmod(ax, w); }
D2D_HD double d2d_fmod_pos(double ax, double w) {   // fmod(ax, w) for ax >= 0, w > 0
    if (ax < w) return ax;
    if (ax < 2.0 * w) return ax - w;
    return d2d_fmod_slow(ax, w);
}
D2D_HD double d2d_pymod(double x, double w) {
    double m = (x < 0) ? -d2d_fmod_pos(-x, w) : d2d_fmod_pos(x, w);
    if (m != 0.0) {
        if (m < 0) m += w;
    } else {
        m = 0.0;  // copysign(0, w), w > 0
    }
    return m;
}

D2D_HD double d2d_sqrt(double v) { return D2D_SQRT(v); }
D2D_HD double d2d_norm2(double x, double y) { return d2d_sqrt(D2D_FMA(y, y, x * x)); }
// `np.linalg.norm([x, y]) <= R` / `< R` for R >= 0, exactly, without the square root unless the squared norm lies within
// 1e-12 (relative) of R^2: sqrt is monotone and correctly rounded, so outside that band the comparison of the squares
// decides (the band is four orders of magnitude wider than the rounding of R*R and of the final sqrt).
D2D_HD bool d2d_norm2_le(double x, double y, double R) {
    const double t = D2D_FMA(y, y, x * x), R2 = R * R;
    if (t <= R2 * (1.0 - 1e-12)) return true;
    if (t >= R2 * (1.0 + 1e-12)) return false;
    return d2d_sqrt(t) <= R;
}
D2D_HD bool d2d_norm2_lt(double x, double y, double R) {
    const double t = D2D_FMA(y, y, x * x), R2 = R * R;
    if (t < R2 * (1.0 - 1e-12)) return true;
    if (t > R2 * (1.0 + 1e-12)) return false;
    return d2d_sqrt(t) < R;
}

// ------------------------------------------------------------------------------------------------ double-double
struct d2d_dd {
    double h, l;
};

D2D_HD d2d_dd dd_two_sum(double a, double b) {
    double s = a + b, bb = s - a;
    d2d_dd r;
    r.h = s;
    r.l = (a - (s - bb)) + (b - bb);
    return r;
}
D2D_HD d2d_dd dd_fast_two_sum(double a, double b) {  // |a| >= |b|
    double s = a + b;
    d2d_dd r;
    r.h = s;
    r.l = b - (s - a);
    return r;
}
D2D_HD d2d_dd dd_two_prod(double a, double b) {
    d2d_dd r;
    r.h = a * b;
    r.l = D2D_FMA(a, b, -r.h);
    return r;
}
D2D_HD d2d_dd dd_add(d2d_dd a, d2d_dd b) {
    d2d_dd s = dd_two_sum(a.h, b.h), t = dd_two_sum(a.l, b.l);
    s.l += t.h;
    s = dd_fast_two_sum(s.h, s.l);
    s.l += t.l;
    return dd_fast_two_sum(s.h, s.l);
}
D2D_HD d2d_dd dd_add_d(d2d_dd a, double b) {
    d2d_dd s = dd_two_sum(a.h, b);
    s.l += a.l;
    return dd_fast_two_sum(s.h, s.l);
}
D2D_HD d2d_dd dd_neg(d2d_dd a) {
    d2d_dd r;
    r.h = -a.h;
    r.l = -a.l;
    return r;
}
D2D_HD d2d_dd dd_mul(d2d_dd a, d2d_dd b) {
    d2d_dd p = dd_two_prod(a.h, b.h);
    p.l += a.h * b.l + a.l * b.h;
    return dd_fast_two_sum(p.h, p.l);
}
D2D_HD d2d_dd dd_mul_d(d2d_dd a, double b) {
    d2d_dd p = dd_two_prod(a.h, b);
    p.l += a.l * b;
    return dd_fast_two_sum(p.h, p.l);
}
// a / b to ~2^-100: one IEEE division for the reciprocal, then long division with multiplications (each correction
// term only needs ~2^-50 relative accuracy because it is 2^-52 of the result).
D2D_HD d2d_dd dd_div(d2d_dd a, d2d_dd b) {
    const double inv = 1.0 / b.h;
    const double q1 = a.h * inv;
    d2d_dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    const double q2 = r.h * inv;
    r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
    const double q3 = r.h * inv;
    d2d_dd q = dd_fast_two_sum(q1, q2);
    return dd_add_d(q, q3);
}

// ------------------------------------------------------------------------------------------------ pi/2 reduction
// r = a - q*(pi/2) as a double-double, q = nearest integer; valid for |a| < ~1e5 (q*PIO2_k exact for q < 2^19).
// Three 33-bit pieces of pi/2 (the classic Cody-Waite split): products with q are exact.
D2D_HD d2d_dd d2d_rem_pio2(double a, int *quadrant) {
    const double INV_PIO2 = 6.36619772367581382433e-01;
    const double P1 = 1.57079632673412561417e+00;   // first 33 bits of pi/2
    const double P2 = 6.07710050630396597660e-11;   // next 33 bits
    const double P3 = 2.02226624871116645580e-21;   // next 33 bits
    const double P3T = 8.47842766036889956997e-32;  // pi/2 - (P1+P2+P3)
    double q = D2D_RINT(a * INV_PIO2);
    *quadrant = ((int)q) & 3;
    double z = a - q * P1;                 // exact (cancellation)
    d2d_dd r = dd_two_sum(z, -(q * P2));   // q*P2 exact
    r = dd_add_d(r, -(q * P3));            // q*P3 exact
    r = dd_add_d(r, -(q * P3T));
    return r;
}

// tan(j/32) for j = 0..25 as double-doubles (generated with 60-digit arithmetic, see tests/test_device_math.py)
#if defined(__CUDA_ARCH__)
#define D2D_TABLE __device__ const
#else
#define D2D_TABLE static const
#endif
D2D_TABLE double D2D_TAN_TAB[26][2] = {
#include "d2d_tan_table.inc"
};
// sin(j/32), cos(j/32) for j = 0..25
D2D_TABLE double D2D_SINCOS_TAB[26][4] = {
#include "d2d_sincos_table.inc"
};

// tan of a small double-double |d| <= 1/64 : d + d*P(d^2), relative error ~1e-24
D2D_HD d2d_dd d2d_tan_small(d2d_dd d) {
    d2d_dd u = dd_two_prod(d.h, d.h);
    u.l += 2.0 * d.h * d.l;
    u = dd_fast_two_sum(u.h, u.l);
    const d2d_dd THIRD = {3.33333333333333314830e-01, 1.85037170770859413132e-17};
    d2d_dd p = dd_mul(u, THIRD);
    const double uh = u.h;
    // 2/15, 17/315, 62/2835, 1382/155925, 21844/6081075, 929569/638512875
    double tail = 1.33333333333333333333e-01 +
                  uh * (5.39682539682539682540e-02 +
                        uh * (2.18694885361552028219e-02 +
                              uh * (8.86323552990219656886e-03 +
                                    uh * (3.59212803657248101693e-03 + uh * 1.45583438705131826825e-03))));
    p = dd_add_d(p, (uh * uh) * tail);
    d2d_dd dp = dd_mul(d, p);
    return dd_add(d, dp);
}

// tan(a) for 0 <= |a| < 1e5, rounded to nearest from a double-double result (rel. error < 2^-90 before rounding).
// Lean double-double arithmetic: every step below carries only the renormalisations its error budget needs
// (tests/test_host.py holds the result to 50-digit mpmath and to the fully renormalised evaluation d2d_tan_ref).
D2D_HD double d2d_tan(double a) {
    // ---- r = a - q*(pi/2), q = nearest integer (Cody-Waite, 33-bit pieces: q*P1, q*P2, q*P3 are exact for q < 2^19)
    const double INV_PIO2 = 6.36619772367581382433e-01;
    const double P1 = 1.57079632673412561417e+00, P2 = 6.07710050630396597660e-11;
    const double P3 = 2.02226624871116645580e-21, P3T = 8.47842766036889956997e-32;
    const double q = D2D_RINT(a * INV_PIO2);
    const bool odd = (((int)q) & 1) != 0;
    const double z = a - q * P1;                       // exact (cancellation)
    d2d_dd r = dd_two_sum(z, -(q * P2));
    r.l += -(q * P3) - (q * P3T);                      // |q*P3| <= 2^-66: its rounding error is < 2^-118
    r = dd_fast_two_sum(r.h, r.l);
    const bool neg = r.h < 0;
    if (neg) { r.h = -r.h; r.l = -r.l; }
    // ---- d = r - j/32, |d| <= 1/64 (+ table end): the subtraction of the heads is exact (Sterbenz / j == 0)
    int j = (int)(r.h * 32.0 + 0.5);
    if (j > 25) j = 25;
    const d2d_dd d = dd_two_sum(r.h - (double)j * 0.03125, r.l);
    // ---- t = tan(d) = d + d*p, p = d^2/3 + 2 d^4/15 + ... (p <= 2^-13.6: it needs ~2^-60 relative accuracy, so only the
    //      1/3 term is carried in double-double; no renormalisation in between, the low words stay far below the heads)
    d2d_dd u = dd_two_prod(d.h, d.h);
    u.l += 2.0 * d.h * d.l;
    const double THIRD_H = 3.33333333333333314830e-01, THIRD_L = 1.85037170770859413132e-17;
    d2d_dd p = dd_two_prod(u.h, THIRD_H);
    const double uh = u.h;
    // 2/15, 17/315, 62/2835, 1382/155925, 21844/6081075, 929569/638512875
    const double tail = 1.33333333333333333333e-01 +
                        uh * (5.39682539682539682540e-02 +
                              uh * (2.18694885361552028219e-02 +
                                    uh * (8.86323552990219656886e-03 +
                                          uh * (3.59212803657248101693e-03 + uh * 1.45583438705131826825e-03))));
    p.l += (uh * THIRD_L + u.l * THIRD_H) + (uh * uh) * tail;
    d2d_dd dp = dd_two_prod(d.h, p.h);
    dp.l += d.h * p.l + d.l * p.h;
    d2d_dd t = dd_fast_two_sum(d.h, dp.h);             // |dp| <= 2^-13 |d|
    t.l += d.l + dp.l;
    // ---- tan(j/32 + d) = (T + t) / (1 - T t); an odd quadrant swaps the operands: tan(r + pi/2) = -1/tan(r).
    //      Branch-free: every lane of a warp runs the same instruction stream (T = 0 makes j == 0 a no-op).
    const double Th = D2D_TAN_TAB[j][0], Tl = D2D_TAN_TAB[j][1];
    d2d_dd num = dd_fast_two_sum(Th, t.h);             // Th >= tan(1/32) > |t| or Th == 0
    num.l += Tl + t.l;
    d2d_dd tt = dd_two_prod(Th, t.h);
    tt.l += Th * t.l + Tl * t.h;
    d2d_dd den = dd_fast_two_sum(1.0, -tt.h);          // |T t| < 1/32
    den.l -= tt.l;
    num = dd_fast_two_sum(num.h, num.l);
    den = dd_fast_two_sum(den.h, den.l);
    const double ah = odd ? den.h : num.h, al = odd ? den.l : num.l;
    const double bh = odd ? num.h : den.h, bl = odd ? num.l : den.l;
    // ---- quotient: q1 = fl(ah / bh); the remainder (ah + al) - q1 (bh + bl) is formed exactly in its leading part
    //      (ah - fl(bh q1) cancels exactly) and in plain double below that, which is 2^-104 of the result
    const double inv = 1.0 / bh;
    const double q1 = ah * inv;
    const d2d_dd pq = dd_two_prod(bh, q1);
    const double rem = (((ah - pq.h) - pq.l) + al) - q1 * bl;
    const double q2 = rem * inv;
    double v = q1 + q2;
    if (odd) v = -v;
    return neg ? -v : v;
}

// the same function with fully renormalised double-double steps (rel. error < 2^-80); kept as the cross-check of the lean
// evaluation above (tests/test_host.py), not used by any kernel
D2D_HD double d2d_tan_ref(double a) {
    int quad;
    d2d_dd r = d2d_rem_pio2(a, &quad);
    const bool neg = r.h < 0;
    if (neg) r = dd_neg(r);
    int j = (int)(r.h * 32.0 + 0.5);
    if (j > 25) j = 25;
    d2d_dd d = dd_add_d(r, -(double)j * 0.03125);
    d2d_dd t = d2d_tan_small(d);
    d2d_dd T;
    T.h = D2D_TAN_TAB[j][0];
    T.l = D2D_TAN_TAB[j][1];
    d2d_dd num = dd_add(T, t);
    d2d_dd den = dd_add_d(dd_neg(dd_mul(T, t)), 1.0);
    const bool odd = (quad & 1) != 0;
    d2d_dd a_, b_;
    a_.h = odd ? den.h : num.h; a_.l = odd ? den.l : num.l;
    b_.h = odd ? num.h : den.h; b_.l = odd ? num.l : den.l;
    d2d_dd res = dd_div(a_, b_);
    if (odd) res = dd_neg(res);
    double v = res.h + res.l;
    return neg ? -v : v;
}

// sin/cos of a small double-double |d| <= 1/64
D2D_HD void d2d_sincos_small(d2d_dd d, d2d_dd *s, d2d_dd *c) {
    d2d_dd u = dd_two_prod(d.h, d.h);
    u.l += 2.0 * d.h * d.l;
    u = dd_fast_two_sum(u.h, u.l);
    const double uh = u.h;
    // sin(d) = d - d^3/6 + d^5/120 - ... ; cos(d) = 1 - d^2/2 + d^4/24 - ...
    const d2d_dd SIXTH = {1.66666666666666657415e-01, 9.25185853854297065662e-18};
    d2d_dd ps = dd_neg(dd_mul(u, SIXTH));
    double ts = 8.33333333333333333333e-03 +
                uh * (-1.98412698412698412698e-04 +
                      uh * (2.75573192239858906526e-06 + uh * (-2.50521083854417187751e-08 + uh * 1.60590438368216145994e-10)));
    ps = dd_add_d(ps, (uh * uh) * ts);
    *s = dd_add(d, dd_mul(d, ps));
    d2d_dd pc = dd_mul_d(u, -0.5);
    const d2d_dd T24 = {4.16666666666666643537e-02, 2.31296463463574266416e-18};
    d2d_dd u2 = dd_mul(u, u);
    pc = dd_add(pc, dd_mul(u2, T24));
    double tc = -1.38888888888888888889e-03 +
                uh * (2.48015873015873015873e-05 + uh * (-2.75573192239858906526e-07 + uh * 2.08767569878680989792e-09));
    pc = dd_add_d(pc, (uh * uh * uh) * tc);
    *c = dd_add_d(pc, 1.0);
}

// sin(a), cos(a) for |a| < 1e5 as double-doubles (relative error ~2^-100)
D2D_HD void d2d_sincos_dd(double a, d2d_dd *sn, d2d_dd *cs) {
    int quad;
    d2d_dd r = d2d_rem_pio2(a, &quad);
    const bool neg = r.h < 0;
    if (neg) r = dd_neg(r);
    int j = (int)(r.h * 32.0 + 0.5);
    if (j > 25) j = 25;
    d2d_dd d = dd_add_d(r, -(double)j * 0.03125);
    d2d_dd sd, cd;
    d2d_sincos_small(d, &sd, &cd);
    d2d_dd S, Cc;
    if (j == 0) {
        S = sd;
        Cc = cd;
    } else {
        d2d_dd Sj = {D2D_SINCOS_TAB[j][0], D2D_SINCOS_TAB[j][1]}, Cj = {D2D_SINCOS_TAB[j][2], D2D_SINCOS_TAB[j][3]};
        S = dd_add(dd_mul(Sj, cd), dd_mul(Cj, sd));
        Cc = dd_add(dd_mul(Cj, cd), dd_neg(dd_mul(Sj, sd)));
    }
    if (neg) S = dd_neg(S);
    switch (quad) {
        case 0: *sn = S; *cs = Cc; break;
        case 1: *sn = Cc; *cs = dd_neg(S); break;
        case 2: *sn = dd_neg(S); *cs = dd_neg(Cc); break;
        default: *sn = dd_neg(Cc); *cs = S; break;
    }
}

// sin(a), cos(a) for |a| < 1e5, each rounded to nearest from a double-double
D2D_HD_NOINLINE void d2d_sincos(double a, double *sn, double *cs) {
    d2d_dd S, C;
    d2d_sincos_dd(a, &S, &C);
    *sn = S.h + S.l;
    *cs = C.h + C.l;
}

// atan2(y, x) rounded to nearest: the library value t0 (CUDA: <= 2 ulp; glibc on the host) plus one Newton step
// tan(theta - t0) = (y cos t0 - x sin t0) / (x cos t0 + y sin t0) with sin / cos of t0 in double-double, so the correction is
// known to ~2^-48 of itself (2^-100 of the result).  The gaze policies compare bearings against bin / wedge boundaries
// (yaw_planner.py:144-189) and a drone on the integer lattice produces bearings that sit exactly on them (atan2(a, a)).
D2D_HD double d2d_atan2_refine(double t0, double y, double x) {
    if (!(fabs(t0) <= 4.0) || t0 == 0.0 || !(fabs(x) <= 1e300) || !(fabs(y) <= 1e300)) return t0;   // NaN, signed zero, infinities
    d2d_dd S, C;
    d2d_sincos_dd(t0, &S, &C);
    const d2d_dd num = dd_add(dd_mul_d(C, y), dd_neg(dd_mul_d(S, x)));
    const d2d_dd den = dd_add(dd_mul_d(C, x), dd_mul_d(S, y));
    return t0 + num.h / den.h;
}
D2D_HD_NOINLINE double d2d_atan2_cr(double y, double x) { return d2d_atan2_refine(atan2(y, x), y, x); }
