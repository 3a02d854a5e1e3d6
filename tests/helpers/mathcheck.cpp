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
