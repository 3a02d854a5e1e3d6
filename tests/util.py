"""Shared helpers of the test-suite: golden fixture loading, oracle env construction, field comparison."""
import glob
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

PARAM_KEYS = ("dt", "map_scale", "map_size", "agent_radius", "drone_max_acceleration", "drone_radius",
              "drone_max_yaw_speed", "drone_view_depth", "drone_view_range", "max_flight_time", "gaze_method", "planner",
              "var_cam", "drone_max_speed", "motion_profile", "pillar_number", "agent_number", "agent_max_speed", "map_id",
              "static_map")


def golden_files(prefix=""):
    """Episode / step fixtures (tests/golden/metrics_*.npz hold difficulty-metric values and have their own tests)."""
    return sorted(f for f in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")) if not os.path.basename(f).startswith("metrics_"))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["params"] = json.loads(str(d["params_json"]))
    return d


def params_from_golden(g):
    from gym_drone2d_activeperception_b200.params import Params
    P = g["params"]
    kw = {k: P[k] for k in PARAM_KEYS if k in P}
    kw["init_pos"] = P["init_position"]
    kw["target_list"] = P["target_list"]
    return Params(debug=False, **kw)


def oracle_params(p, **kw):
    import oracle
    return oracle.make_params(**kw, dt=p.dt, map_scale=p.map_scale, map_size=p.map_size, agent_radius=p.agent_radius,
                              drone_max_acceleration=p.drone_max_acceleration, drone_radius=p.drone_radius,
                              drone_max_yaw_speed=p.drone_max_yaw_speed, drone_view_depth=p.drone_view_depth,
                              drone_view_range=p.drone_view_range, max_flight_time=p.max_flight_time, var_cam=p.var_cam,
                              drone_max_speed=p.drone_max_speed, planner=p.planner)


_OP_CACHE = {}


def oracle_env_from_world(p, world, i=None, drone=None, jerk_tie_orders=None):
    """world: dict of arrays (one env, or batched with index i)."""
    import oracle
    w = {k: (v[i] if i is not None else v) for k, v in world.items()}
    if p.planner == "Jerk_Primitive":          # the per-heading tables are shared by every env built from one params object
        key = (tuple(sorted((k, repr(v)) for k, v in vars(p).items())), None if jerk_tie_orders is None else jerk_tie_orders.tobytes())
        if key not in _OP_CACHE:
            _OP_CACHE[key] = oracle_params(p, jerk_tie_orders=jerk_tie_orders)
        op = _OP_CACHE[key]
    else:
        op = oracle_params(p)
    e = oracle.OracleEnv(op, w["agent_pos"], w["agent_pref"], w["agent_radius"], w["gt_grid"],
                         w["tracker_radius"], drone=(w["drone_pose"] if drone is None else drone),
                         targets=p.target_list)
    if "rng_key" in w:
        e.set_rng(w["rng_key"], w["rng_pos"], w["rng_has_gauss"], w["rng_gauss"])
    if getattr(p, "motion_profile", "CVM") == "RVO":
        e.set_rvo(w["agent_vel"], w["obstacles"])
    return e


def world_from_golden(g, copies=1):
    w = dict(agent_pos=g["agent_pos0"], agent_pref=g["agent_pref0"], agent_radius=g["agent_radius"],
             tracker_radius=g["tracker_radius"], gt_grid=g["gt_grid"], drone_pose=g["drone0"])
    if "rng_key" in g and g["params"].get("var_cam", 0) != 0:
        w.update(rng_key=g["rng_key"], rng_pos=g["rng_pos"], rng_has_gauss=g["rng_has_gauss"], rng_gauss=g["rng_gauss"])
    if g["params"].get("motion_profile", "CVM") == "RVO":
        w.update(agent_vel=g["agent_vel0"], obstacles=g["obstacles"])
    return {k: np.ascontiguousarray(np.stack([v] * copies)) for k, v in w.items()}


def action_table():
    """The six values Oxford.plan returns (yaw_planner.py:65,127)."""
    return np.arange(-80, 80, 80 / 3) / 80


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


# ------------------------------------------------------------------------------------------ full-batch parity helpers
DISCRETE_FIELDS = ("belief", "hit", "local_map", "collision_flag", "dead_lock_flag", "freezing_flag", "done",
                   "state_machine", "fail_count", "steps", "tracker_buffer_count", "tracker_buffer_ts", "yaw_angle")
STATE_FIELDS = ("drone_x", "drone_y", "drone_yaw", "drone_vx", "drone_vy", "agent_pos", "agent_pref")


def oracle_batch(p, worlds, poses=None, threads=None, jerk_tie_orders=None):
    """oracle.OracleBatch over every env of `worlds` (initial state; optional externally set drone poses [B,3])."""
    import oracle
    B = worlds["drone_pose"].shape[0]
    envs = [oracle_env_from_world(p, worlds, i, drone=None if poses is None else poses[i], jerk_tie_orders=jerk_tie_orders)
            for i in range(B)]
    return oracle.OracleBatch(envs, threads=threads)


def gpu_fields(env, names):
    import torch
    torch.cuda.synchronize()
    return {k: env.buffer(k).cpu().numpy() for k in names}


def batch_mismatch(h, o, n, trackers=True, planner=False):
    """Per-env mismatch masks between the CUDA batch (h: dict of host arrays named like d2d_get_buffer) and the oracle
    batch (o: OracleBatch.gather()).  Returns (discrete: dict name -> bool[B], rel: dict name -> float[B]).
    Discrete fields are compared bit for bit; continuous ones by max |a-b| / max(1, |b|) per env (absolute below 1:
    positions are O(100), velocities O(10); the north-star bar is 1e-9 relative)."""
    B = o["done"].shape[0]
    d, r = {}, {}
    flat = lambda a: np.asarray(a).reshape(B, -1)
    d["belief"] = (flat(h["belief"]) != flat(o["belief"])).any(1)
    d["hit"] = (flat(h["hit"])[:, :n] != flat(o["hit"])[:, :n]).any(1)
    d["local_map"] = (flat(h["local_map"]) != flat(o["local_map"])).any(1)
    d["yaw_angle"] = flat(h["yaw_angle"])[:, 0] != o["yaw_angle"]
    for k in ("collision_flag", "dead_lock_flag", "freezing_flag", "done", "state_machine", "fail_count", "steps"):
        d[k] = h[k].astype(np.int64) != o[k].astype(np.int64)
    if planner:
        d["traj_len"] = (h["traj_nseg"].astype(np.int64) * planner - h["traj_cursor"]) != o["traj_len"]
        d["plan_ok"] = h["plan_ok"].astype(np.int64) != o["plan_ok"]
        d["replan"] = h["replan"].astype(np.int64) != o["replan"]

    def rel(a, b):
        a, b = flat(a).astype(np.float64), flat(b).astype(np.float64)
        return (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max(1) if a.shape[1] else np.zeros(B)
    for k in ("drone_x", "drone_y", "drone_yaw", "drone_vx", "drone_vy"):
        r[k] = rel(h[k], o[k])
    r["agent_pos"] = rel(h["agent_pos"][:, :n], o["agent_pos"][:, :n])
    r["agent_pref"] = rel(h["agent_pref"][:, :n], o["agent_pref"][:, :n])
    if trackers:
        act = o["tracker_active"][:, :n].astype(bool)
        d["tracker_active"] = (h["tracker_active"][:, :n].astype(bool) != act).any(1)
        d["tracker_ts"] = ((h["tracker_ts"][:, :n].astype(np.int64) != o["tracker_ts"][:, :n]) & act).any(1)
        d["tracker_radius"] = (h["tracker_radius"][:, :n] != o["tracker_radius"][:, :n]).any(1)
        d["tracker_buffer"] = (h["tracker_buffer_count"].astype(np.int64) != o["tracker_buffer_count"]) | \
                              (h["tracker_buffer_ts"].astype(np.int64) != o["tracker_buffer_ts"])
        mu_h, mu_o = h["tracker_mu"][:, :n], o["tracker_mu"][:, :n]
        sg_h, sg_o = h["tracker_sigma"][:, :n].reshape(B, n, 16), o["tracker_sigma"][:, :n]
        em = np.abs(mu_h - mu_o) / np.maximum(1.0, np.abs(mu_o))
        es = np.abs(sg_h - sg_o) / np.maximum(1.0, np.abs(sg_o))
        r["tracker_mu"] = np.where(act[:, :, None], em, 0.0).reshape(B, -1).max(1) if n else np.zeros(B)
        r["tracker_sigma"] = np.where(act[:, :, None], es, 0.0).reshape(B, -1).max(1) if n else np.zeros(B)
    return d, r


BATCH_FIELDS = ["belief", "hit", "local_map", "yaw_angle", "collision_flag", "dead_lock_flag", "freezing_flag", "done",
                "state_machine", "fail_count", "steps", "drone_x", "drone_y", "drone_yaw", "drone_vx", "drone_vy", "agent_pos",
                "agent_pref"]
TRACKER_FIELDS = ["tracker_active", "tracker_ts", "tracker_radius", "tracker_mu", "tracker_sigma", "tracker_buffer_count",
                  "tracker_buffer_ts"]
PLANNER_FIELDS = ["traj_nseg", "traj_cursor", "plan_ok", "replan"]
