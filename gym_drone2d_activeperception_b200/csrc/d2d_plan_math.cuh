// d2d_plan_math.cuh -- integer-cell helpers of the A* kernels (belief probes on integer-valued sample coordinates, the packed
// dict keys of Primitive_Node.get_index) as host + device functions: included by d2d_plan.cuh for the kernels and by
// tests/helpers/mathcheck.cpp, which holds them to Python restatements of the reference expressions on the CPU (tests/test_host.py).
#pragma once
#include <stdint.h>
#include "d2d_math.cuh"
#ifndef D2D_GRID
#define D2D_GRID 50
#endif

// the same probe for integer-valued coordinates (A* samples, np.around at traj_planner.py:181): cells by integer arithmetic
D2D_HD int d2d_belief_probe_int(const uint8_t *bel, int x, int y, int w, int h) {
    if (x >= w || x < 0 || y >= h || y < 0) return 1;            // utils.py:546-547
    return bel[(x / 10) * D2D_GRID + (y / 10)];
}

// Primitive_Node.get_index (traj_planner.py:92-93): (round(px)//10, round(py)//10, round(vx), round(vy))
D2D_HD long long d2d_floordiv10(long long a) {
    long long q = a / 10;
    if ((a % 10 != 0) && (a < 0)) q -= 1;
    return q;
}
// packs (round(px)//10, round(py)//10, round(vx), round(vy)) into 28 bits; valid for |v| < 64 and cells in [-32, 95]
D2D_HD uint32_t d2d_node_key32(double px, double py, double vx, double vy) {
    const int a = (int)d2d_floordiv10((long long)rint(px)) + 32, b = (int)d2d_floordiv10((long long)rint(py)) + 32;
    const int c = (int)rint(vx) + 64, d = (int)rint(vy) + 64;
    return ((uint32_t)(a & 127) << 21) | ((uint32_t)(b & 127) << 14) | ((uint32_t)(c & 127) << 7) | (uint32_t)(d & 127);
}

// d2d_node_key32 in 32-bit integer arithmetic (same value wherever that one is valid: |coordinate| < 2^31)
D2D_HD uint32_t d2d_node_key32i(double px, double py, double vx, double vy) {
#ifdef __CUDA_ARCH__
    const int ax = __double2int_rn(px), ay = __double2int_rn(py);
#else
    const int ax = (int)rint(px), ay = (int)rint(py);
#endif
    int a = ax / 10, b = ay / 10;
    if (ax < 0 && a * 10 != ax) a -= 1;              // floor division (Python's //)
    if (ay < 0 && b * 10 != ay) b -= 1;
#ifdef __CUDA_ARCH__
    const int c = __double2int_rn(vx) + 64, d = __double2int_rn(vy) + 64;
#else
    const int c = (int)rint(vx) + 64, d = (int)rint(vy) + 64;
#endif
    return ((uint32_t)((a + 32) & 127) << 21) | ((uint32_t)((b + 32) & 127) << 14) | ((uint32_t)(c & 127) << 7) | (uint32_t)(d & 127);
}

// the five belief probes of Planner.is_free (traj_planner.py:35-47) with all loads in flight: 1 if any of them reads OCCUPIED
// (outside the map counts as occupied, utils.py:546-547)
D2D_HD int d2d_probe5_occ(const uint8_t *bel, int x, int y, int sd, int w, int h) {
    const int xm = x - sd, xp = x + sd, ym = y - sd, yp = y + sd;
    if (xm < 0 || xp < 0 || xm >= w || xp >= w || ym < 0 || yp < 0 || ym >= h || yp >= h)       // a probe may leave the map
        return (d2d_belief_probe_int(bel, xm, y, w, h) == 1) | (d2d_belief_probe_int(bel, x, y, w, h) == 1) |
               (d2d_belief_probe_int(bel, xp, y, w, h) == 1) | (d2d_belief_probe_int(bel, x, ym, w, h) == 1) |
               (d2d_belief_probe_int(bel, x, yp, w, h) == 1);
    const int cx = (x / 10) * D2D_GRID, cy = y / 10;
    const int a = bel[(xm / 10) * D2D_GRID + cy], b = bel[cx + cy], c = bel[(xp / 10) * D2D_GRID + cy];
    const int d = bel[cx + ym / 10], f = bel[cx + yp / 10];
    return (a == 1) | (b == 1) | (c == 1) | (d == 1) | (f == 1);
}

