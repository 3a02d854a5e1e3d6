"""Run the UNMODIFIED reference (read from /root/reference) headless and record per-step state.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It exists to
(a) generate the golden fixtures under tests/golden/ (see oracle/gen_golden.py) and
(b) let `-m "not gpu"` tests cross-check the C restatement (oracle/drone2d_oracle.c) against the live
    reference when /root/reference is present (it is absent on the GPU box; those tests skip there).

The reference needs gym / pygame / matplotlib / cvxpy at import time only; oracle/shims/ provides empty
stand-ins.  The reference code itself is never copied or modified: instance methods are
wrapped from the outside to observe values (castRays' measurement list, the planner's verdicts).
"""
import os
import sys
import contextlib

import numpy as np

REFERENCE_ROOT = os.environ.get("D2D_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "drone_v2.py"))


@contextlib.contextmanager
def _reference_cwd():
    # the reference loads 'maps/*.npy' relative to its own root (utils.py:72, drone_v2.py:49)
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def import_reference():
    """Returns (utils, drone_v2, traj_planner, yaw_planner) modules of the reference."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import utils as ref_utils  # noqa
    import traj_planner as ref_traj  # noqa
    import yaw_planner as ref_yaw  # noqa
    from envs import drone_v2 as ref_env  # noqa
    return ref_utils, ref_env, ref_traj, ref_yaw


DEFAULT_PARAMS = dict(debug=False, planner="NoMove", gaze_method="Oxford", agent_number=10, agent_radius=15,
                      agent_max_speed=20, drone_max_speed=40, map_id=1, static_map="maps/empty_map.npy")

OXFORD_ACTIONS = None


def oxford_action_set():
    """The six action values Oxford.plan can return (yaw_planner.py:65,127)."""
    global OXFORD_ACTIONS
    if OXFORD_ACTIONS is None:
        OXFORD_ACTIONS = (np.arange(-80, 80, 80 / 3) / 80).tolist()
    return OXFORD_ACTIONS


def make_env(**kw):
    ref_utils, ref_env, _, _ = import_reference()
    pk = dict(DEFAULT_PARAMS)
    pk.update(kw)
    params = ref_utils.Params(**pk)
    with _reference_cwd():
        env = ref_env.Drone2DEnv2(params)
    return env, params


def world_snapshot(env):
    """Initial world as arrays (what A0 produces): agents, per-tracker radius, ground-truth grid."""
    n = len(env.agents)
    return dict(
        agent_pos0=np.array([np.asarray(a.position, dtype=np.float64) for a in env.agents]).reshape(n, 2),
        agent_pref0=np.array([np.asarray(a.pref_velocity, dtype=np.float64) for a in env.agents]).reshape(n, 2),
        agent_radius=np.array([float(a.radius) for a in env.agents]),
        tracker_radius=np.array([float(env.drone.trackers[i].radius) for i in range(n)]),
        gt_grid=env.map_gt.grid_map.copy(),
        drone0=np.array([float(env.drone.x), float(env.drone.y), float(env.drone.yaw)]),
        agent_vel0=np.array([np.asarray(a.velocity, dtype=np.float64) for a in env.agents]).reshape(n, 2),
        obstacles=np.array([np.asarray(o, dtype=np.float64) for o in env.obstacles]).reshape(len(env.obstacles), 3),
    )


def run_episode(steps, actions=None, policy=None, set_pose=None, stop_on_done=False, record_trackers=True,
                record_oxford=False, record_gt=False, **param_kw):
    """Step the reference and record everything the parity tests compare.

    actions: sequence of floats (cycled) when policy is None; policy: 'Oxford' uses the reference's own
    Oxford class exactly as experiment.py:33-34,69 does (class object used as the instance).
    set_pose: optional (x, y, yaw) written onto env.drone before the first step, the way
    script/difficulty_calculator/glob_survivability_calculator.py:36-37 does.
    """
    ref_utils, ref_env, ref_traj, ref_yaw = import_reference()
    env, params = make_env(**param_kw)
    if set_pose is not None:
        env.drone.x, env.drone.y = set_pose[0], set_pose[1]
        if len(set_pose) > 2:
            env.drone.yaw = set_pose[2]
    world = world_snapshot(env)
    st = np.random.get_state()           # the reference uses the GLOBAL legacy stream (drone_v2.py:80, utils.py:605)
    world["rng_key"] = np.asarray(st[1], dtype=np.uint32).copy()
    world["rng_pos"] = np.array(int(st[2]))
    world["rng_has_gauss"] = np.array(int(st[3]))
    world["rng_gauss"] = np.array(float(st[4]))
    n = len(env.agents)

    pol = None
    if policy is not None:
        pol = getattr(ref_yaw, policy)      # class object used as the instance, as experiment.py:33-34 does
        pol.__init__(pol, params)

    # observe castRays' measurement list without touching the reference source
    seen = {}
    orig_cast = env.drone.raycast.castRays

    def cast_spy(player, gt, agents):
        rays, newly, meas = orig_cast(player, gt, agents)
        seen["hit"] = np.array([0 if m is None else 1 for m in meas], dtype=np.int8)
        seen["newly"] = newly
        return rays, newly, meas

    env.drone.raycast.castRays = cast_spy

    plans = []
    orig_plan = env.planner.plan
    orig_replan = env.planner.replan_check

    def plan_spy(drone, dt):
        before = len(env.planner.trajectory)
        ok = orig_plan(drone, dt)
        seen["plan_ok"] = bool(ok)
        seen["planned"] = (before == 0) and hasattr(env.planner, "u_space")
        if seen["planned"] and ok:
            tr = env.planner.trajectory
            plans.append(dict(step=env.steps,
                              positions=np.array(tr.positions, dtype=np.float64).reshape(-1, 2),
                              velocities=np.array(tr.velocities, dtype=np.float64).reshape(-1, 2)))
        return ok

    def replan_spy(drone):
        r = orig_replan(drone)
        seen["replan"] = bool(r[0])
        return r

    env.planner.plan = plan_spy
    env.planner.replan_check = replan_spy

    rec = {k: [] for k in ("action", "agent_pos", "agent_pref", "agent_vel", "belief", "hit", "newly", "collision", "done",
                           "dead_lock", "freezing", "state_machine", "fail_count", "drone", "drone_vel",
                           "local_map", "yaw_obs", "traj_len", "replan", "plan_ok", "planned", "target",
                           "trk_active", "trk_mu", "trk_sigma", "trk_radius", "trk_ts", "buf_count", "buf_ts",
                           "ox_last", "tracked_agent", "gt_dyn")}
    with _reference_cwd():
        for t in range(steps):
            if pol is not None:
                a = pol.plan(pol, env.info)
                if record_oxford and policy == "Oxford":
                    rec["ox_last"].append(pol.last_time_observed_map.copy())
            else:
                a = actions[t % len(actions)]
            state, reward, done, info = env.step(a)
            assert reward == 0
            rec["action"].append(float(a))
            rec["agent_pos"].append(np.array([np.asarray(ag.position, dtype=np.float64) for ag in env.agents]).reshape(n, 2))
            rec["agent_pref"].append(np.array([np.asarray(ag.pref_velocity, dtype=np.float64) for ag in env.agents]).reshape(n, 2))
            rec["agent_vel"].append(np.array([np.asarray(ag.velocity, dtype=np.float64) for ag in env.agents]).reshape(n, 2))
            rec["belief"].append(env.drone.map.grid_map.copy())
            rec["hit"].append(seen["hit"].copy())
            rec["newly"].append(int(seen["newly"]))
            rec["collision"].append(int(info["collision_flag"]))
            rec["done"].append(bool(done))
            rec["dead_lock"].append(int(info["dead_lock_flag"]))
            rec["freezing"].append(int(info["freezing_flag"]))
            rec["state_machine"].append(int(info["state_machine"]))
            rec["fail_count"].append(int(env.fail_count))
            rec["drone"].append([float(env.drone.x), float(env.drone.y), float(env.drone.yaw)])
            rec["drone_vel"].append(np.asarray(env.drone.velocity, dtype=np.float64).copy())
            lm = state["local_map"]
            assert lm.dtype == np.uint8 and lm.shape == (1, 33, 33), (lm.dtype, lm.shape)
            assert np.array_equal(lm, state["swep_map"])
            rec["local_map"].append(lm[0].copy())
            rec["yaw_obs"].append(np.float32(state["yaw_angle"][0]))
            rec["traj_len"].append(len(env.planner.trajectory))
            rec["replan"].append(bool(seen.get("replan", False)))
            rec["plan_ok"].append(bool(seen.get("plan_ok", True)))
            rec["planned"].append(bool(seen.get("planned", False)))
            rec["target"].append(np.asarray(env.planner.target, dtype=np.float64).copy())
            rec["tracked_agent"].append(int(env.tracked_agent))
            if record_gt:      # env.map_gt.grid_map with the DYNAMIC_OCCUPIED marks of update_dynamic_grid (utils.py:527-540)
                rec["gt_dyn"].append(env.map_gt.grid_map.copy())
            if record_trackers:
                trk = env.drone.trackers[:n]
                rec["trk_active"].append(np.array([bool(k.active) for k in trk]))
                rec["trk_mu"].append(np.array([np.asarray(k.mu_upds[-1], dtype=np.float64).reshape(4) for k in trk]).reshape(n, 4))
                rec["trk_sigma"].append(np.array([np.asarray(k.Sigma_upds[-1], dtype=np.float64).reshape(16) for k in trk]).reshape(n, 16))
                rec["trk_radius"].append(np.array([float(k.radius) for k in trk]))
                rec["trk_ts"].append(np.array([len(k.ts) for k in trk], dtype=np.int64))
                rec["buf_count"].append(len(env.tracker_buffer))
                rec["buf_ts"].append(int(sum(len(k.ts) for k in env.tracker_buffer)))
            if done and stop_on_done:
                break

    out = dict(world)
    for k, v in rec.items():
        if len(v) == 0:
            continue
        out[k] = np.array(v)
    out["plans"] = plans
    out["n_agents"] = n
    out["params"] = {k: v for k, v in vars(params).items()}
    return out
