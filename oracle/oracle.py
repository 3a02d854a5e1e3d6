"""ctypes front-end of the parity oracle (oracle/drone2d_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdrone2d_oracle.so")
MAX_TARGETS = 8
TRAJ_CAP = 2048


class Params(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("map_scale", C.c_double), ("map_w", C.c_double), ("map_h", C.c_double),
        ("agent_radius", C.c_double), ("drone_max_acc", C.c_double), ("drone_radius", C.c_double),
        ("drone_max_yaw_speed", C.c_double), ("view_depth", C.c_double), ("view_range", C.c_double),
        ("max_flight_time", C.c_double), ("var_cam", C.c_double), ("drone_max_speed", C.c_double),
        ("planner", C.c_int32), ("n_rays", C.c_int32), ("gw", C.c_int32), ("gh", C.c_int32), ("local", C.c_int32),
        ("n_u", C.c_int32), ("n_samp", C.c_int32), ("n_way", C.c_int32), ("pad0", C.c_int32),
        ("u_space", C.c_double * 64),
        ("t_samp", C.c_double * 32), ("t_samp2", C.c_double * 32),
        ("t_way", C.c_double * 64), ("t_way2", C.c_double * 64), ("t_way_x2", C.c_double * 64),
        ("n_yaw", C.c_int32), ("pad1", C.c_int32),
        ("v_yaw_space", C.c_double * 16),
        ("ox_cos_thresh", C.c_double),
        ("jerk", C.c_void_p),
    ]


JERK_H, JERK_MAXT = 72, 40


class JerkTables(C.Structure):
    _fields_ = [
        ("dx", C.c_double * JERK_H), ("dy", C.c_double * JERK_H),
        ("T", C.c_double * JERK_H), ("Tp", (C.c_double * 4) * JERK_H),
        ("times", C.c_int32 * JERK_H),
        ("tt", (C.c_double * JERK_MAXT) * JERK_H), ("ttp", ((C.c_double * 4) * JERK_MAXT) * JERK_H),
        ("tie_order", (C.c_uint8 * JERK_H) * 144),
        ("unexpected_ties", C.c_int64),
    ]


def jerk_tie_orders():
    """`cost[:, 0].argsort()` (traj_planner.py:471-478) of THIS host's numpy for the 144 goal bearings that are exact
    multiples of 2.5 degrees, where the cost is symmetric about the bearing and every pair of headings ties: numpy's default
    argsort is unstable (x86-simd-sort networks; AVX-512 and AVX2 builds order ties differently), so the order is recorded
    from the library instead of being guessed.  Returns uint8 [144, 72]."""
    theta_range = np.arange(0, 360, 5)
    out = np.zeros((144, JERK_H), dtype=np.uint8)
    for m in range(144):
        phi_h = 2.5 * m
        cost = np.zeros((theta_range.shape[0], 2))
        for i, theta in enumerate(theta_range):
            a = abs(theta % 360 - phi_h % 360)
            cost[i, 0] = 1 * (a if a <= 180 else 360 - a) ** 2
            cost[i, 1] = theta
        out[m] = cost[:, 0].argsort()
    return out


def jerk_tables(drone_max_speed=40, dt=0.1, d=30, tie_orders=None):
    """Per-heading constants of Jerk_Primitive.generate_primitive (traj_planner.py:413-460), evaluated with the reference's
    own expressions (np.cos / np.sin of math.radians, numpy-scalar `**`, np.arange, np.floor)."""
    from math import radians
    from numpy.linalg import norm
    t = JerkTables()
    theta_range = np.arange(0, 360, 5)
    cost1 = np.zeros(theta_range.shape[0])
    for i, theta in enumerate(theta_range):
        cost1[i] = theta
    v_max = drone_max_speed
    for i in range(JERK_H):
        theta_h = cost1[i]                                     # cost[seq, 1]: the heading as a float64 array element
        delt_x = d * np.cos(radians(theta_h))
        delt_y = d * np.sin(radians(theta_h))
        T = 1.2 * norm(np.array([delt_x, delt_y])) / (norm(v_max))
        T = T if T >= 0.5 else 0.5
        times = int(np.floor(T / dt))
        assert 0 < times <= JERK_MAXT
        tarr = np.arange(dt, times * dt + dt, dt)
        t.dx[i], t.dy[i], t.T[i], t.times[i] = float(delt_x), float(delt_y), float(T), times
        for k, e in enumerate((2, 3, 4, 5)):
            t.Tp[i][k] = float(T ** e)
        for jj in range(times):
            tt = tarr[jj]
            t.tt[i][jj] = float(tt)
            for k, e in enumerate((2, 3, 4, 5)):
                t.ttp[i][jj][k] = float(tt ** e)
    ties = jerk_tie_orders() if tie_orders is None else np.asarray(tie_orders, dtype=np.uint8)
    for m in range(144):
        for i in range(JERK_H):
            t.tie_order[m][i] = int(ties[m, i])
    return t


_P = C.POINTER


class Env(C.Structure):
    _fields_ = [
        ("p", Params), ("n", C.c_int32), ("pad", C.c_int32),
        ("apos", _P(C.c_double)), ("apref", _P(C.c_double)), ("arad", _P(C.c_double)),
        ("gt", _P(C.c_uint8)), ("belief", _P(C.c_uint8)), ("hit", _P(C.c_int8)),
        ("x", C.c_double), ("y", C.c_double), ("yaw", C.c_double), ("vx", C.c_double), ("vy", C.c_double),
        ("steps", C.c_int32), ("state_machine", C.c_int32), ("fail_count", C.c_int32),
        ("target_cursor", C.c_int32), ("n_targets", C.c_int32),
        ("collision", C.c_int32), ("dead_lock", C.c_int32), ("freezing", C.c_int32), ("done", C.c_int32),
        ("replan", C.c_int32), ("plan_ok", C.c_int32), ("planned", C.c_int32), ("newly_tracked", C.c_int32),
        ("targets", (C.c_double * 2) * MAX_TARGETS), ("target", C.c_double * 4),
        ("tracked_agent", C.c_int64),
        ("trk_active", _P(C.c_uint8)), ("trk_mu", _P(C.c_double)), ("trk_sigma", _P(C.c_double)),
        ("trk_radius", _P(C.c_double)), ("trk_ts", _P(C.c_int64)),
        ("buf_count", C.c_int64), ("buf_ts", C.c_int64),
        ("traj_len", C.c_int32), ("traj_head", C.c_int32),
        ("traj_pos", _P(C.c_double)), ("traj_vel", _P(C.c_double)),
        ("local_map", _P(C.c_uint8)), ("yaw_obs", C.c_float), ("pad2", C.c_int32),
        ("ox_last", _P(C.c_double)), ("ox_swep", _P(C.c_double)),
        ("n_samples", C.c_int64),
        ("rng_key", C.c_uint32 * 624), ("rng_pos", C.c_int32), ("rng_has_gauss", C.c_int32), ("rng_gauss", C.c_double),
        ("meas", _P(C.c_double)),
        ("motion_rvo", C.c_int32), ("n_obs", C.c_int32), ("avel", _P(C.c_double)), ("obs", _P(C.c_double)),
        ("rvo_fallbacks", C.c_int64),
        ("owl_U", C.c_double * 36), ("owl_q", C.c_int32), ("pad3", C.c_int32), ("owl_u", C.c_double),
        ("ax", C.c_double), ("ay", C.c_double), ("next_ax", C.c_double), ("next_ay", C.c_double),
    ]


_lib = None
_KEEP = []


def build(force=False):
    if force or not os.path.isfile(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "drone2d_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdrone2d_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.d2do_create.restype = _P(Env)
        L.d2do_create.argtypes = [_P(Params), C.c_int]
        L.d2do_destroy.argtypes = [_P(Env)]
        for name in ("d2do_agents_step", "d2do_trackers_update", "d2do_local_map"):
            getattr(L, name).argtypes = [_P(Env)]
            getattr(L, name).restype = None
        for name in ("d2do_cast_rays", "d2do_replan_check", "d2do_plan", "d2do_is_collide"):
            getattr(L, name).argtypes = [_P(Env)]
            getattr(L, name).restype = C.c_int
        L.d2do_step.argtypes = [_P(Env), C.c_double]
        L.d2do_step.restype = C.c_int
        L.d2do_oxford_plan.argtypes = [_P(Env)]
        L.d2do_oxford_plan.restype = C.c_double
        L.d2do_owl_plan.argtypes = [_P(Env)]
        L.d2do_owl_plan.restype = C.c_double
        L.d2do_policy_plan.argtypes = [_P(Env), C.c_int]
        L.d2do_policy_plan.restype = C.c_double
        L.d2do_run_many.argtypes = [_P(_P(Env)), C.c_int, C.c_int, _P(C.c_double)]
        L.d2do_run_many.restype = C.c_int64
        L.d2do_run_many_oxford.argtypes = [_P(_P(Env)), C.c_int, C.c_int]
        L.d2do_run_many_oxford.restype = C.c_int64
        L.d2do_set_rvo.argtypes = [_P(Env), _P(C.c_double), _P(C.c_double), C.c_int]
        L.d2do_snapshot.argtypes = [_P(Env)]
        L.d2do_snapshot.restype = C.c_void_p
        L.d2do_snapshot_free.argtypes = [C.c_void_p]
        L.d2do_restore.argtypes = [_P(Env), C.c_void_p]
        L.d2do_step_batch.argtypes = [_P(_P(Env)), _P(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.d2do_step_batch.restype = C.c_int64
        L.d2do_gather.argtypes = [_P(_P(Env)), C.c_int] + [C.c_void_p] * 14
        L.d2do_run_batch.argtypes = [_P(_P(Env)), _P(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.d2do_run_batch.restype = C.c_int64
        L.d2do_gather.restype = None
        L.d2do_sizeof_params.restype = C.c_size_t
        L.d2do_sizeof_env.restype = C.c_size_t
        assert L.d2do_sizeof_params() == C.sizeof(Params), (L.d2do_sizeof_params(), C.sizeof(Params))
        L.d2do_sizeof_jerk_tables.restype = C.c_size_t
        assert L.d2do_sizeof_jerk_tables() == C.sizeof(JerkTables), (L.d2do_sizeof_jerk_tables(), C.sizeof(JerkTables))
        assert L.d2do_sizeof_env() == C.sizeof(Env), (L.d2do_sizeof_env(), C.sizeof(Env))
        _lib = L
    return _lib


def oxford_cos_threshold(view_range_deg):
    """c* = min{c : np.arccos(c) <= radians(view_range/2)}, bisection over doubles with the ARRAY
    np.arccos (the SIMD loop Oxford.get_view_map runs, yaw_planner.py:78)."""
    import math
    ang = math.radians(view_range_deg / 2)

    def ok(c):
        return bool(np.arccos(np.array([c, c, c, c, c, c, c, c, c]))[0] <= ang)

    lo, hi = -1.0, 1.0  # ok(hi) True, ok(lo) False unless the wedge is the full circle
    if ok(lo):
        return -1.0
    while True:
        mid = lo + (hi - lo) / 2
        if mid == lo or mid == hi:
            break
        if ok(mid):
            hi = mid
        else:
            lo = mid
    return hi


def make_params(dt=0.1, map_scale=10, map_size=(500, 500), agent_radius=10, drone_max_acceleration=40,
                drone_radius=10, drone_max_yaw_speed=80, drone_view_depth=80, drone_view_range=90,
                max_flight_time=80, var_cam=0, drone_max_speed=40, planner="NoMove", strip_width=10, jerk_tie_orders=None,
                **_ignored):
    """Builds the C params block with the reference's own numpy expressions for the lookup tables
    (traj_planner.py:98-104,180,212; yaw_planner.py:65; utils.py:587)."""
    import math
    p = Params()
    p.dt, p.map_scale, p.map_w, p.map_h = dt, map_scale, map_size[0], map_size[1]
    p.agent_radius, p.drone_max_acc, p.drone_radius = agent_radius, drone_max_acceleration, drone_radius
    p.drone_max_yaw_speed, p.view_depth, p.view_range = drone_max_yaw_speed, drone_view_depth, drone_view_range
    p.max_flight_time, p.var_cam, p.drone_max_speed = max_flight_time, var_cam, drone_max_speed
    p.planner = {"NoMove": 0, "Primitive": 1, "Jerk_Primitive": 2}[planner]
    if p.planner == 2:
        jt = jerk_tables(drone_max_speed, dt, tie_orders=jerk_tie_orders)
        _KEEP.append(jt)                                       # shared by every env created from these params
        p.jerk = C.addressof(jt)
    p.n_rays = math.ceil(map_size[0] / strip_width)
    p.gw, p.gh = map_size[0] // map_scale, map_size[1] // map_scale
    p.local = 4 * (drone_view_depth // map_scale) + 1
    if drone_max_speed <= 40:
        u = np.arange(-drone_max_acceleration, drone_max_acceleration, 0.4 * drone_max_speed - 5)
    else:
        u = np.arange(-drone_max_acceleration, drone_max_acceleration, 4)
    pdt = 2
    sample_num = drone_max_speed * pdt // map_scale
    ts = np.arange(0, pdt, pdt / sample_num)
    tw = np.arange(pdt, 0, -dt)
    assert len(u) <= 64 and len(ts) <= 32 and len(tw) <= 64
    p.n_u, p.n_samp, p.n_way = len(u), len(ts), len(tw)
    for i, v in enumerate(u):
        p.u_space[i] = float(v)
    for i, t in enumerate(ts):
        p.t_samp[i] = float(t)
        p.t_samp2[i] = float(t ** 2)      # numpy-scalar power == libm pow(t, 2.0)
    for i, t in enumerate(tw):
        p.t_way[i] = float(t)
        p.t_way2[i] = float(t ** 2)
        p.t_way_x2[i] = float(2 * t)
    vy = np.arange(-drone_max_yaw_speed, drone_max_yaw_speed, drone_max_yaw_speed / 3)
    p.n_yaw = len(vy)
    for i, v in enumerate(vy):
        p.v_yaw_space[i] = float(v)
    p.ox_cos_thresh = oxford_cos_threshold(drone_view_range)
    return p


class OracleEnv(object):
    """One CPU env.  World state is supplied by the caller (arrays from the reference or from the product's
    host-side world generator); numpy views alias the C arrays."""

    def __init__(self, params, agent_pos, agent_pref, agent_radius, gt_grid, tracker_radius=None,
                 drone=(50.0, 50.0, 270.0), targets=((50, 460),)):
        L = lib()
        self.n = int(len(agent_radius))
        self._params = params
        self._ptr = L.d2do_create(C.byref(params), self.n)
        self.c = self._ptr.contents
        n = max(self.n, 1)
        cells = params.gw * params.gh
        as_arr = np.ctypeslib.as_array
        self.apos = as_arr(self.c.apos, (n, 2))
        self.apref = as_arr(self.c.apref, (n, 2))
        self.arad = as_arr(self.c.arad, (n,))
        self.gt = as_arr(self.c.gt, (params.gw, params.gh))
        self.belief = as_arr(self.c.belief, (params.gw, params.gh))
        self.hit = as_arr(self.c.hit, (n,))
        self.trk_active = as_arr(self.c.trk_active, (n,))
        self.trk_mu = as_arr(self.c.trk_mu, (n, 4))
        self.trk_sigma = as_arr(self.c.trk_sigma, (n, 16))
        self.trk_radius = as_arr(self.c.trk_radius, (n,))
        self.trk_ts = as_arr(self.c.trk_ts, (n,))
        self.traj_pos = as_arr(self.c.traj_pos, (TRAJ_CAP, 2))
        self.traj_vel = as_arr(self.c.traj_vel, (TRAJ_CAP, 2))
        self.local_map = as_arr(self.c.local_map, (params.local, params.local))
        self.ox_last = as_arr(self.c.ox_last, (params.gw, params.gh))
        if self.n:
            self.apos[:self.n] = np.asarray(agent_pos, dtype=np.float64).reshape(self.n, 2)
            self.apref[:self.n] = np.asarray(agent_pref, dtype=np.float64).reshape(self.n, 2)
            self.arad[:self.n] = np.asarray(agent_radius, dtype=np.float64)
            if tracker_radius is not None:
                self.trk_radius[:self.n] = np.asarray(tracker_radius, dtype=np.float64)
        self.gt[:] = np.asarray(gt_grid, dtype=np.uint8)
        self.c.x, self.c.y, self.c.yaw = float(drone[0]), float(drone[1]), float(drone[2])
        self.c.n_targets = len(targets)
        for i, t in enumerate(targets):
            self.c.targets[i][0], self.c.targets[i][1] = float(t[0]), float(t[1])
        # Planner.__init__ traj_planner.py:22: target = [drone.x, drone.y, 0, 0]; Jerk_Primitive.__init__ :406: np.zeros(4)
        if params.planner == 2:
            self.c.target[0], self.c.target[1] = 0.0, 0.0
        else:
            self.c.target[0], self.c.target[1] = float(drone[0]), float(drone[1])

    def set_rng(self, key, pos, has_gauss, gauss):
        """legacy np.random state right after world generation (RandomState.get_state()); needed when var_cam != 0"""
        for i, v in enumerate(np.asarray(key, dtype=np.uint32).tolist()):
            self.c.rng_key[i] = v
        self.c.rng_pos, self.c.rng_has_gauss, self.c.rng_gauss = int(pos), int(has_gauss), float(gauss)

    def set_rvo(self, vel0, obstacles):
        """motion_profile == 'RVO' (drone_v2.py:169-175): initial agent velocities [n,2] and pillars [m,3] (x, y, rad)"""
        v = np.ascontiguousarray(np.asarray(vel0, dtype=np.float64).reshape(max(self.n, 0), 2))
        o = np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 3))
        dp = C.POINTER(C.c_double)
        lib().d2do_set_rvo(self._ptr, v.ctypes.data_as(dp), o.ctypes.data_as(dp), int(o.shape[0]))
        self.avel = np.ctypeslib.as_array(self.c.avel, (max(self.n, 1), 2))

    def step(self, a):
        return bool(lib().d2do_step(self._ptr, float(a)))

    def oxford_plan(self):
        return float(lib().d2do_oxford_plan(self._ptr))

    def owl_plan(self):
        """Owl.plan (yaw_planner.py:191-222) with the class-object-as-instance call pattern of experiment.py:33-34"""
        return float(lib().d2do_owl_plan(self._ptr))

    def policy_plan(self, kind):
        """kind: 0 NoControl, 1 Rotating, 2 LookAhead, 3 LookGoal (yaw_planner.py)"""
        return float(lib().d2do_policy_plan(self._ptr, int(kind)))

    def trajectory(self):
        h, l = self.c.traj_head, self.c.traj_len
        return self.traj_pos[h:h + l].copy(), self.traj_vel[h:h + l].copy()

    def close(self):
        if self._ptr is not None:
            lib().d2do_destroy(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


POLICY = {"scripted": -1, "NoControl": 0, "Rotating": 1, "LookAhead": 2, "LookGoal": 3, "Oxford": 4, "Owl": 5}


class OracleBatch(object):
    """B oracle envs stepped side by side with a CUDA batch (full-batch parity tests, mismatch campaigns).

    `envs` are OracleEnv objects in their INITIAL state (world, pose, rng, rvo already set): a snapshot of each is
    kept so that `step(auto_reset=True)` mirrors the batched env's auto-reset (an env whose previous step returned
    done is re-initialised before it plans / steps).  Slices of the batch run on a thread pool (ctypes drops the GIL);
    `gather` packs the compared outputs into batch arrays in C."""

    INT_FIELDS = ("collision_flag", "dead_lock_flag", "freezing_flag", "done", "state_machine", "fail_count", "steps",
                  "tracker_buffer_count", "tracker_buffer_ts", "traj_len", "replan", "plan_ok")

    def __init__(self, envs, threads=None):
        from concurrent.futures import ThreadPoolExecutor
        self.L = lib()
        self.envs = list(envs)
        self.B = len(self.envs)
        self.n = self.envs[0].n
        self.params = self.envs[0]._params
        self._ptrs = (_P(Env) * self.B)(*[e._ptr for e in self.envs])
        self._snaps = (C.c_void_p * self.B)(*[self.L.d2do_snapshot(e._ptr) for e in self.envs])
        self.threads = max(1, min(threads or len(os.sched_getaffinity(0)), self.B))
        self._pool = ThreadPoolExecutor(self.threads)
        b = np.linspace(0, self.B, self.threads + 1).astype(int)
        self._slices = [(int(b[i]), int(b[i + 1])) for i in range(self.threads) if b[i + 1] > b[i]]

    def _off(self, arr, i0, ctype):
        return C.cast(C.addressof(arr) + i0 * C.sizeof(ctype), C.POINTER(ctype))

    def step(self, actions=None, policy="scripted", auto_reset=True):
        """Steps every env once; returns (applied actions [B] float64, number of envs that reported done)."""
        pol = POLICY[policy] if isinstance(policy, str) else int(policy)
        acts = None
        if pol < 0:
            acts = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.B)
        out = np.empty(self.B, dtype=np.float64)

        def run(sl):
            i0, i1 = sl
            return self.L.d2do_step_batch(self._off(self._ptrs, i0, _P(Env)), self._off(self._snaps, i0, C.c_void_p),
                                          i1 - i0, None if acts is None else acts[i0:].ctypes.data, 1 if auto_reset else 0,
                                          pol, out[i0:].ctypes.data)
        dones = sum(self._pool.map(run, self._slices))
        return out, int(dones)

    def run(self, steps, actions=None, policy="scripted", auto_reset=True):
        """`steps` steps of every env with the step loop inside C (one call per thread): the timed CPU-baseline loop.
        actions: [steps, B] float64 when policy is "scripted".  Returns the number of done reports."""
        pol = POLICY[policy] if isinstance(policy, str) else int(policy)
        parts = []
        if pol < 0:
            acts = np.asarray(actions, dtype=np.float64).reshape(steps, self.B)
            parts = [np.ascontiguousarray(acts[:, i0:i1]) for i0, i1 in self._slices]

        def go(j):
            i0, i1 = self._slices[j]
            return self.L.d2do_run_batch(self._off(self._ptrs, i0, _P(Env)), self._off(self._snaps, i0, C.c_void_p), i1 - i0,
                                         int(steps), parts[j].ctypes.data if parts else None, 1 if auto_reset else 0, pol)
        return int(sum(self._pool.map(go, range(len(self._slices)))))

    def gather(self, trackers=True, rvo=False):
        B, n = self.B, max(self.n, 1)
        p = self.params
        cells, L = p.gw * p.gh, p.local * p.local
        o = {"belief": np.empty((B, p.gw, p.gh), np.uint8), "hit": np.empty((B, n), np.int8),
             "local_map": np.empty((B, p.local, p.local), np.uint8), "ints": np.empty((B, 12), np.int32),
             "state": np.empty((B, 5), np.float64), "yaw_angle": np.empty((B,), np.float32),
             "agent_pos": np.empty((B, n, 2), np.float64), "agent_pref": np.empty((B, n, 2), np.float64)}
        if trackers:
            o.update(tracker_active=np.empty((B, n), np.uint8), tracker_ts=np.empty((B, n), np.int64),
                     tracker_radius=np.empty((B, n), np.float64), tracker_mu=np.empty((B, n, 4), np.float64),
                     tracker_sigma=np.empty((B, n, 16), np.float64))
        if rvo:
            o["agent_vel"] = np.empty((B, n, 2), np.float64)
        order = ["belief", "hit", "local_map", "ints", "state", "yaw_angle", "agent_pos", "agent_pref", "tracker_active",
                 "tracker_ts", "tracker_radius", "tracker_mu", "tracker_sigma", "agent_vel"]

        def run(sl):
            i0, i1 = sl
            args = [(o[k][i0:].ctypes.data if k in o else None) for k in order]
            self.L.d2do_gather(self._off(self._ptrs, i0, _P(Env)), i1 - i0, *args)
        list(self._pool.map(run, self._slices))
        for j, name in enumerate(self.INT_FIELDS):
            o[name] = o["ints"][:, j]
        for j, name in enumerate(("drone_x", "drone_y", "drone_yaw", "drone_vx", "drone_vy")):
            o[name] = o["state"][:, j]
        return o

    def close(self):
        if self._snaps is not None:
            for s in self._snaps:
                self.L.d2do_snapshot_free(s)
            self._snaps = None
            self._pool.shutdown()
            for e in self.envs:
                e.close()
