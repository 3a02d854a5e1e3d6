// d2d_rvo_math.cuh -- the cone arithmetic of the RVO motion profile (RVO.RVO_update / in_between, utils.py:299-460) as
// host + device functions: included by d2d_rvo.cuh for the kernel and by tests/helpers/mathcheck.cpp, which holds the
// cross-product shortcut to the atan2 evaluation on the CPU (tests/test_host.py).
#pragma once
#include "d2d_math.cuh"

// rx, ry / lx, ly: the vectors atan2 turned into th_right / th_left; mode: which branch of in_between the cone takes
// (0: |th_right - th_left| <= 3.14, 1: left < 0 < right, 2: right < 0 < left, 3: never inside, 4: decide with atan2 only)
struct RvoCone { double tx, ty, th_left, th_right, dist, rad, rx, ry, lx, ly; int mode, pad_; };

// RVO.in_between utils.py:434-460 (None counts as False)
D2D_HD bool d2d_rvo_in_between(double theta_right, double theta_dif, double theta_left) {
    if (fabs(theta_right - theta_left) <= 3.14) return theta_right <= theta_dif && theta_dif <= theta_left;
    if (theta_left < 0 && theta_right > 0) {
        theta_left += 2 * 3.14;
        if (theta_dif < 0) theta_dif += 2 * 3.14;
        return theta_right <= theta_dif && theta_dif <= theta_left;
    }
    if (theta_left > 0 && theta_right < 0) {
        theta_right += 2 * 3.14;
        if (theta_dif < 0) theta_dif += 2 * 3.14;
        return theta_left <= theta_dif && theta_dif <= theta_right;
    }
    return false;
}

// --- in_between without atan2.  theta_dif = atan2(d) is only ever COMPARED with th_right = atan2(r) and th_left = atan2(l), and
// atan2 is monotone in the true angle, so the order of two angles in (-pi, pi] follows from the half-planes of the vectors and
// the sign of their cross product.  Wherever that is not certain by a wide margin (|cross| within 1e-9 relative, a vector
// within 1e-9 of the x axis -- the +-pi seam and the sign of theta_dif --, or one of the cases where the reference's `3.14` /
// `2*3.14` constants bite) the caller falls back to the atan2 evaluation, so the verdicts are those of d2d_rvo_in_between
// on CUDA's atan2 everywhere.  ~25 instructions instead of ~150 for the N * 160 * (N - 1) tests of a step.
#define D2D_RVO_TOL 1e-9
// 1: angle(a) <= angle(b), 0: angle(a) > angle(b), -1: uncertain.  Both vectors strictly off the x axis (checked by the caller).
D2D_HD int d2d_rvo_angle_le(double ax, double ay, double bx, double by) {
    const bool ua = ay > 0.0, ub = by > 0.0;
    if (ua != ub) return ub ? 1 : 0;                      // lower half-plane (-pi, 0) comes before the upper one (0, pi)
    const double p = ax * by, q = ay * bx, cr = p - q;    // same open half-plane: the angles differ by less than pi
    if (fabs(cr) <= D2D_RVO_TOL * (fabs(p) + fabs(q))) return -1;
    return cr > 0.0 ? 1 : 0;
}
// 1 / 0: verdict of in_between(th_right, atan2(dy, dx), th_left); -1: evaluate with atan2
D2D_HD int d2d_rvo_inside_fast(const RvoCone &c, double dx, double dy) {
#ifdef D2D_RVO_NOFAST
    return -1;      // A/B build (tools/parity_campaign.py with D2D_LIB=...): every verdict through atan2, as before round 2
#endif
    if (c.mode == 3) return 0;
    if (c.mode == 4 || !(fabs(dy) > D2D_RVO_TOL * fabs(dx))) return -1;
    if (c.mode == 0) {                                    // th_right <= theta_dif <= th_left
        const int a = d2d_rvo_angle_le(c.rx, c.ry, dx, dy);
        if (a == 0) return 0;
        const int b = d2d_rvo_angle_le(dx, dy, c.lx, c.ly);
        if (b == 0) return 0;
        return (a < 0 || b < 0) ? -1 : 1;
    }
    if (c.mode == 1)                                      // left < 0 < right: theta_dif >= 0 -> right <= dif; < 0 -> dif <= left
        return dy > 0.0 ? d2d_rvo_angle_le(c.rx, c.ry, dx, dy) : d2d_rvo_angle_le(dx, dy, c.lx, c.ly);
    // mode 2, right < 0 < left: theta_dif >= 0 -> left <= dif; < 0 -> dif <= right
    return dy > 0.0 ? d2d_rvo_angle_le(c.lx, c.ly, dx, dy) : d2d_rvo_angle_le(dx, dy, c.rx, c.ry);
}

D2D_HD void d2d_rvo_make_cone(RvoCone &c, double pAx, double pAy, double tx, double ty, double pBx,
                                                  double pBy, double reach) {
    c.tx = tx; c.ty = ty;
    double dist = d2d_norm2(pAx - pBx, pAy - pBy);
    const double theta_BA = atan2(pBy - pAy, pBx - pAx);
    if (reach > dist) dist = reach;
    const double ort = asin(reach / dist);
    const double tl = theta_BA + ort, tr = theta_BA - ort;
    c.dist = dist; c.rad = reach;
    c.rx = cos(tr); c.ry = sin(tr); c.lx = cos(tl); c.ly = sin(tl);
    c.th_right = atan2(c.ry, c.rx);             // atan2(bound_right[1], bound_right[0]), utils.py:372
    c.th_left = atan2(c.ly, c.lx);
    // which branch of in_between (utils.py:434-460) this cone takes, and whether the shortcut is safe for it: in the wrapped
    // branches the reference adds 2 * 3.14 (not 2 pi) to one bound and to a negative theta_dif, which only changes a verdict
    // when a bound lies within 0.0032 of +-pi -- those cones keep the atan2 path
    const double PI_UP = 3.1415926535897936;    // > pi: no atan2 result exceeds it
    int mode;
    if (fabs(c.th_right - c.th_left) <= 3.14) mode = 0;
    else if (c.th_left < 0 && c.th_right > 0) mode = (c.th_left + 2 * 3.14 < PI_UP || c.th_right > 2 * 3.14 - PI_UP) ? 4 : 1;
    else if (c.th_left > 0 && c.th_right < 0) mode = (c.th_right + 2 * 3.14 < PI_UP || c.th_left > 2 * 3.14 - PI_UP) ? 4 : 2;
    else mode = 3;
    if (!(fabs(c.ry) > D2D_RVO_TOL * fabs(c.rx)) || !(fabs(c.ly) > D2D_RVO_TOL * fabs(c.lx))) mode = (mode == 3) ? 3 : 4;
    c.mode = mode; c.pad_ = 0;
}

