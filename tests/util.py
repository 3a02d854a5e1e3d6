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


def oracle_params(p):
    import oracle
    return oracle.make_params(dt=p.dt, map_scale=p.map_scale, map_size=p.map_size, agent_radius=p.agent_radius,
                              drone_max_acceleration=p.drone_max_acceleration, drone_radius=p.drone_radius,
                              drone_max_yaw_speed=p.drone_max_yaw_speed, drone_view_depth=p.drone_view_depth,
                              drone_view_range=p.drone_view_range, max_flight_time=p.max_flight_time, var_cam=p.var_cam,
                              drone_max_speed=p.drone_max_speed, planner=p.planner)


def oracle_env_from_world(p, world, i=None, drone=None):
    """world: dict of arrays (one env, or batched with index i)."""
    import oracle
    w = {k: (v[i] if i is not None else v) for k, v in world.items()}
    e = oracle.OracleEnv(oracle_params(p), w["agent_pos"], w["agent_pref"], w["agent_radius"], w["gt_grid"],
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
