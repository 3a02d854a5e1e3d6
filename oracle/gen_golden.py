"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (needs /root/reference; build container only).

    python oracle/gen_golden.py            # rewrites every fixture

TEST INFRASTRUCTURE ONLY.  The reference ships no tests / golden vectors, so these fixtures --
outputs of the reference itself on seeded inputs -- are what pins the oracle and, through it, the CUDA path.
Provenance recorded in each file: python / numpy / glibc versions and CPU model of the generating box
(results depend on libm and on OpenBLAS' FMA usage at the ulp level).
"""
import json
import os
import platform
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner as rr  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def provenance():
    cpu = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    cpu = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return dict(python=platform.python_version(), numpy=np.__version__, libc=" ".join(platform.libc_ver()), cpu=cpu,
                reference="smoggy-P/gym-Drone2D-ActivePerception @ /root/reference (unmodified)")


def _jsonable(params):
    out = {}
    for k, v in params.items():
        if isinstance(v, (np.integer,)):
            v = int(v)
        elif isinstance(v, (np.floating,)):
            v = float(v)
        out[k] = v
    return out


def save(name, r, keys, extra=None):
    d = {k: r[k] for k in keys if k in r}
    d["params_json"] = np.array(json.dumps(_jsonable(r["params"])))
    d["provenance_json"] = np.array(json.dumps(provenance()))
    d["n_agents"] = np.array(r["n_agents"])
    if r.get("plans"):
        d["plan_steps"] = np.array([p["step"] for p in r["plans"]])
        for i, p in enumerate(r["plans"]):
            d["plan%d_pos" % i] = p["positions"]
            d["plan%d_vel" % i] = p["velocities"]
    if extra:
        d.update(extra)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-40s %8.1f KB  steps=%d N=%d" % (name, os.path.getsize(path) / 1024, len(r["done"]), r["n_agents"]))


WORLD = ["agent_pos0", "agent_pref0", "agent_radius", "tracker_radius", "gt_grid", "drone0",
         "rng_key", "rng_pos", "rng_has_gauss", "rng_gauss"]
STEP_CORE = ["action", "agent_pos", "agent_pref", "belief", "hit", "newly", "collision", "done", "dead_lock",
             "freezing", "state_machine", "fail_count", "drone", "drone_vel", "local_map", "yaw_obs", "target"]
TRACK = ["trk_active", "trk_mu", "trk_sigma", "trk_radius", "trk_ts", "buf_count", "buf_ts", "tracked_agent"]
PLAN = ["traj_len", "replan", "plan_ok", "planned"]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""      # optional substring filter: regenerate matching fixtures only
    os.makedirs(OUT, exist_ok=True)
    acts = rr.oxford_action_set()
    # --- perception + dynamics (NoMove), the four shipped maps (BASELINE.json configs 2-5 worlds)
    cases = [
        ("nomove_empty_s1", dict(map_id=1), 120, None, True),
        ("nomove_empty_s8_border", dict(map_id=8), 60, None, True),   # seed 8: an agent disc overrides a border cell
        ("nomove_empty_s2_pose", dict(map_id=2), 100, (123.4567, 301.25, 33.3), True),  # runs into a collision
        ("nomove_obstacle_s0", dict(map_id=0, static_map="maps/obstacle_map.npy", agent_radius=10), 80, None, True),
        ("nomove_shaped_s3", dict(map_id=3, static_map="maps/shaped_obstacle_map.npy", agent_number=50,
                                  agent_radius=10, agent_max_speed=40), 40, (250.0, 250.0, 10.0), False),
        ("nomove_random0_s5", dict(map_id=5, static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15,
                                   agent_max_speed=40), 30, (300.5, 200.25, 200.0), False),
        ("nomove_slow_agents_s4", dict(map_id=4, agent_max_speed=4, agent_number=6), 80, (250.0, 250.0, 0.0), True),
        # measurement noise: sigma * np.random.randn(2) per in-view agent continues the seeded legacy stream (utils.py:605)
        ("nomove_noise_s3", dict(map_id=3, agent_number=12, var_cam=1), 120, (250.0, 250.0, 0.0), True),
    ]
    for name, kw, steps, pose, trk in cases:
        if only not in name:
            continue
        r = rr.run_episode(steps, actions=acts, set_pose=pose, planner="NoMove", **kw)
        save(name, r, WORLD + STEP_CORE + (TRACK if trk else []))
    # --- full loop: Primitive planner + Oxford gaze (BASELINE.json config 1 and variants)
    eps = [
        ("episode_cfg1_s1", dict(map_id=1)),
        ("episode_s2", dict(map_id=2)),
        ("episode_speed20_s4", dict(map_id=4, drone_max_speed=20, agent_number=30)),
        ("episode_obstacle_s0", dict(map_id=0, static_map="maps/obstacle_map.npy", agent_radius=10)),
        ("episode_noise_s6", dict(map_id=6, agent_number=12, var_cam=1)),
    ]
    for name, kw in eps:
        if only not in name:
            continue
        r = rr.run_episode(800, policy="Oxford", planner="Primitive", stop_on_done=True, record_oxford=True, **kw)
        save(name, r, WORLD + STEP_CORE + TRACK + PLAN + ["ox_last"])
    # --- env.map_gt.grid_map as callers of the facade see it (value-3 marks; seed 8 has a border cell overridden at init)
    if only in "gtgrid_empty_s8":
        r = rr.run_episode(12, actions=acts, planner="NoMove", record_gt=True, record_trackers=False, map_id=8)
        save("gtgrid_empty_s8", r, WORLD + ["action", "gt_dyn"])
    # --- Owl gaze policy (yaw_planner.py:151-222) with the reference's call pattern (class object as instance, experiment.py:33-34)
    owl = [
        ("owl_s3", dict(map_id=3, agent_number=8)),
        ("owl_crowd_s7", dict(map_id=7, agent_number=20, agent_radius=15, agent_max_speed=20)),
        ("owl_obstacle_s5", dict(map_id=5, agent_number=30, static_map="maps/obstacle_map.npy")),
        ("owl_speed20_s9", dict(map_id=9, agent_number=10, drone_max_speed=20)),
    ]
    for name, kw in owl:
        if only not in name:
            continue
        r = rr.run_episode(800, policy="Owl", planner="Primitive", gaze_method="Owl", stop_on_done=True, **kw)
        save(name, r, WORLD + STEP_CORE + TRACK + PLAN)
    # --- Jerk_Primitive planner (traj_planner.py:403-516).  jerk_s3 is the canonical scenario (start (50, 50), target (50, 460):
    #     goal bearing exactly 90 degrees, every pair of headings ties in the unstable argsort) with a blocked straight line.
    #     The fixture records how THIS machine's numpy ordered the ties (oracle.jerk_tie_orders) next to the episode.
    import oracle as _oracle
    jerk = [
        ("jerk_s3", dict(map_id=3, agent_number=8), "LookAhead"),
        ("jerk_crowd_s7", dict(map_id=7, agent_number=20, agent_radius=15, agent_max_speed=20), "LookAhead"),
        ("jerk_obstacle_s5", dict(map_id=5, agent_number=30, static_map="maps/obstacle_map.npy"), "LookAhead"),
        ("jerk_speed20_s9", dict(map_id=9, agent_number=10, drone_max_speed=20), "LookGoal"),
        ("jerk_offset_s2", dict(map_id=2, agent_number=10, target_list=[[73, 460]]), "Rotating"),
    ]
    for name, kw, pol in jerk:
        if only not in name:
            continue
        r = rr.run_episode(800, policy=pol, planner="Jerk_Primitive", gaze_method=pol, stop_on_done=True, **kw)
        save(name, r, WORLD + STEP_CORE + TRACK + PLAN, extra={"jerk_tie_orders": _oracle.jerk_tie_orders()})
    # --- RVO motion profile (utils.py:299-460): agents avoid each other and the pillars; velocity is its own array
    rvo = [
        ("rvo_nomove_pillars_s3", dict(map_id=3, agent_number=6, pillar_number=2), 60, None, "NoMove"),
        ("rvo_nomove_crowd_s1", dict(map_id=1, agent_number=40, agent_radius=15, agent_max_speed=20), 30, (250.0, 250.0, 90.0), "NoMove"),
        ("rvo_primitive_s7", dict(map_id=7, agent_number=8, pillar_number=3), 120, None, "Primitive"),
    ]
    for name, kw, steps, pose, planner in rvo:
        if only not in name:
            continue
        r = rr.run_episode(steps, actions=acts, set_pose=pose, planner=planner, motion_profile="RVO", stop_on_done=True, **kw)
        save(name, r, WORLD + ["agent_vel0", "obstacles", "agent_vel"] + STEP_CORE + TRACK + (PLAN if planner == "Primitive" else []))


if __name__ == "__main__":
    main()
