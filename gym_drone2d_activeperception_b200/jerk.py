"""Host-side tables of the Jerk_Primitive planner (traj_planner.py:403-516).

Everything about a candidate heading that does not depend on the drone's state -- the end-point offset, the primitive's
duration T and its powers, the sample times and their powers -- is evaluated HERE with the reference's own numpy / math
expressions and handed to the library (d2d_set_jerk_tables), so that np.cos / np.sin (numpy's own loops, not libm's), the
numpy-scalar `**` (libm pow), np.arange and np.floor are reproduced by construction.

The heading order: Jerk_Primitive.plan ranks the 72 headings with `cost[:, 0].argsort()` (:478), numpy's default UNSTABLE
sort.  The cost is the squared angular distance to the goal bearing phi_h, so it is symmetric about phi_h and pairs of
headings tie exactly when phi_h % 360 is a multiple of 2.5 -- which is where the canonical scenario starts (start (50, 50),
target (50, 460): bearing exactly 90).  How numpy orders those ties depends on its sorting network (x86-simd-sort: AVX-512 and
AVX2 builds differ), so the order is RECORDED from the host's numpy for each of the 144 such bearings instead of being guessed;
for any other bearing the costs are distinct and ascending order is unique."""
import ctypes as C
from math import radians

import numpy as np
from numpy.linalg import norm

from . import _native


def tie_orders():
    """uint8 [144, 72]: cost[:, 0].argsort() of this host's numpy for phi_h % 360 == 2.5 * m (traj_planner.py:471-478)."""
    theta_range = np.arange(0, 360, 5)
    out = np.zeros((144, _native.JERK_H), dtype=np.uint8)
    for m in range(144):
        phi_h = 2.5 * m
        cost = np.zeros((theta_range.shape[0], 2))
        for i, theta in enumerate(theta_range):
            cost[i, 0] = 1 * (abs(theta % 360 - phi_h % 360) if abs(theta % 360 - phi_h % 360) <= 180
                              else 360 - abs(theta % 360 - phi_h % 360)) ** 2
            cost[i, 1] = theta
        out[m] = cost[:, 0].argsort()
    return out


def make_tables(params, orders=None):
    """d2d_jerk_tables for `params` (drone_max_speed, dt); `orders` overrides the recorded tie orders (tests pin them to the
    ones of the machine that generated the golden fixtures)."""
    t = _native.D2DJerkTables()
    d = 30                                                      # Jerk_Primitive.__init__ :409
    v_max, dt = params.drone_max_speed, params.dt
    theta_col = np.zeros(_native.JERK_H)
    for i, theta in enumerate(np.arange(0, 360, 5)):
        theta_col[i] = theta                                    # cost[i, 1] = theta
    for i in range(_native.JERK_H):
        theta_h = theta_col[i]
        delt_x = d * np.cos(radians(theta_h))                   # generate_primitive :414-415
        delt_y = d * np.sin(radians(theta_h))
        T = 1.2 * norm(np.array([delt_x, delt_y])) / (norm(v_max))
        T = T if T >= 0.5 else 0.5
        times = int(np.floor(T / dt))
        if not 0 < times <= _native.JERK_MAXT:
            raise ValueError("Jerk_Primitive: %d samples per primitive (drone_max_speed too low for the table size)" % times)
        tarr = np.arange(dt, times * dt + dt, dt)
        t.dx[i], t.dy[i], t.T[i], t.times[i] = float(delt_x), float(delt_y), float(T), times
        for k, e in enumerate((2, 3, 4, 5)):
            t.Tp[i][k] = float(T ** e)
        for jj in range(times):
            tt = tarr[jj]
            t.tt[i][jj] = float(tt)
            for k, e in enumerate((2, 3, 4, 5)):
                t.ttp[i][jj][k] = float(tt ** e)
    orders = tie_orders() if orders is None else np.asarray(orders, dtype=np.uint8)
    if orders.shape != (144, _native.JERK_H):
        raise ValueError("jerk tie orders must be [144, 72]")
    C.memmove(C.addressof(t.tie_order), np.ascontiguousarray(orders).ctypes.data, 144 * _native.JERK_H)
    return t
