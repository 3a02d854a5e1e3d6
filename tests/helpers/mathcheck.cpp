// Host-compiled twin of csrc/d2d_math.cuh for the CPU test-suite (tests/test_device_math.py).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: same arithmetic as nvcc -fmad=false).
#include "../../gym_drone2d_activeperception_b200/csrc/d2d_math.cuh"
extern "C" {
void mc_tan(const double *a, double *out, long n) { for (long i = 0; i < n; i++) out[i] = d2d_tan(a[i]); }
void mc_tan_ref(const double *a, double *out, long n) { for (long i = 0; i < n; i++) out[i] = d2d_tan_ref(a[i]); }
void mc_sincos(const double *a, double *s, double *c, long n) { for (long i = 0; i < n; i++) d2d_sincos(a[i], &s[i], &c[i]); }
void mc_cell(const double *x, int *out, long n, double scale) { for (long i = 0; i < n; i++) out[i] = d2d_cell(x[i], scale, 1.0 / scale); }
void mc_pymod(const double *x, double *out, long n, double w) { for (long i = 0; i < n; i++) out[i] = d2d_pymod(x[i], w); }
void mc_norm2_cmp(const double *x, const double *y, const double *R, int *le, int *lt, long n) { for (long i = 0; i < n; i++) { le[i] = d2d_norm2_le(x[i], y[i], R[i]); lt[i] = d2d_norm2_lt(x[i], y[i], R[i]); } }
// d2d_atan2_cr with the library seed moved by `ulps` units in the last place first (the device seed is CUDA's atan2, <= 2 ulp)
void mc_atan2(const double *y, const double *x, double *out, long n, int ulps) {
    for (long i = 0; i < n; i++) {
        double t0 = atan2(y[i], x[i]);
        for (int k = 0; k < (ulps < 0 ? -ulps : ulps); k++) t0 = nextafter(t0, ulps < 0 ? -10.0 : 10.0);
        out[i] = d2d_atan2_refine(t0, y[i], x[i]);
    }
}
void mc_norm2(const double *x, const double *y, double *out, long n) { for (long i = 0; i < n; i++) out[i] = d2d_norm2(x[i], y[i]); }
}

// RVO cone tests (csrc/d2d_rvo_math.cuh): for cone i (agent A at pA, neighbour at pB, combined radius `reach`) and probe
// vector d: fast[i] = the cross-product shortcut's verdict (-1 = "ask atan2"), slow[i] = in_between on atan2 -- the kernel takes
// fast when it is >= 0 and slow otherwise, so wherever fast >= 0 the two must agree.
#include "../../gym_drone2d_activeperception_b200/csrc/d2d_rvo_math.cuh"
extern "C" void mc_rvo_inside(const double *pA, const double *pB, const double *reach, const double *d, int *fast, int *slow,
                              int *mode, long n) {
    for (long i = 0; i < n; i++) {
        RvoCone c;
        d2d_rvo_make_cone(c, pA[2 * i], pA[2 * i + 1], 0.0, 0.0, pB[2 * i], pB[2 * i + 1], reach[i]);
        fast[i] = d2d_rvo_inside_fast(c, d[2 * i], d[2 * i + 1]);
        slow[i] = d2d_rvo_in_between(c.th_right, atan2(d[2 * i + 1], d[2 * i]), c.th_left) ? 1 : 0;
        mode[i] = c.mode;
    }
}

// integer-cell helpers of the A* kernels (csrc/d2d_plan_math.cuh)
#include "../../gym_drone2d_activeperception_b200/csrc/d2d_plan_math.cuh"
extern "C" void mc_probe5(const uint8_t *bel, const int *x, const int *y, int *out, long n, int sd, int w, int h) {
    for (long i = 0; i < n; i++) out[i] = d2d_probe5_occ(bel, x[i], y[i], sd, w, h);
}
extern "C" void mc_node_keys(const double *pv, unsigned *k32, unsigned *k32i, long n) {
    for (long i = 0; i < n; i++) {
        k32[i] = d2d_node_key32(pv[4 * i], pv[4 * i + 1], pv[4 * i + 2], pv[4 * i + 3]);
        k32i[i] = d2d_node_key32i(pv[4 * i], pv[4 * i + 1], pv[4 * i + 2], pv[4 * i + 3]);
    }
}
