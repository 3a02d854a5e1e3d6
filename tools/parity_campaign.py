#!/usr/bin/env python
"""Randomised parity campaign: >= 1e7 env-steps of the CUDA path (default kernels, auto-reset on) held to the oracle,
every env every step, and the residual DISCRETE mismatch count reported instead of assumed zero.

    python tools/parity_campaign.py [--scale 1.0] [--out gpurun_out/r2_parity_campaign.json]

An env whose discrete outputs (belief grid, hit mask, observation, flags, done, integer bookkeeping) differ from the
oracle's is counted once, recorded (leg, env, step, fields) and excluded from the rest of its leg (it has diverged);
continuous state is tracked as the maximum relative error over the envs still in parity.
The oracle is test infrastructure: this tool is a checker, never a product path."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

LEGS = [
    dict(name="cfg2 empty N=10 NoMove (bench default workload)", static_map="maps/empty_map.npy", agent_number=10,
         agent_radius=15, agent_max_speed=20, planner="NoMove", B=8192, steps=400),
    dict(name="cfg2 scattered poses", static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20,
         planner="NoMove", B=8192, steps=300, scatter=True),
    dict(name="cfg2 scattered poses through d2d_rollout (16 steps per launch, compared after every launch)", static_map="maps/empty_map.npy",
         agent_number=10, agent_radius=15, agent_max_speed=20, planner="NoMove", B=4096, steps=640, scatter=True, rollout=16),
    dict(name="obstacle_map N=24 NoMove scattered", static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10,
         agent_max_speed=20, planner="NoMove", B=8192, steps=200, scatter=True),
    dict(name="shaped_obstacle_map N=96 NoMove scattered", static_map="maps/shaped_obstacle_map.npy", agent_number=50,
         agent_radius=10, agent_max_speed=40, planner="NoMove", B=4096, steps=150, scatter=True),
    dict(name="random_map_0 N=142 NoMove scattered", static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15,
         agent_max_speed=40, planner="NoMove", B=4096, steps=100, scatter=True),
    dict(name="obstacle_map Primitive + Oxford (config 4 loop)", static_map="maps/obstacle_map.npy", agent_number=10,
         agent_radius=10, agent_max_speed=20, planner="Primitive", gaze="Oxford", B=8192, steps=300),
    dict(name="empty_map Primitive + Oxford (config 1 loop)", static_map="maps/empty_map.npy", agent_number=10,
         agent_radius=15, agent_max_speed=20, planner="Primitive", gaze="Oxford", B=4096, steps=250),
    dict(name="empty_map Primitive + Owl", static_map="maps/empty_map.npy", agent_number=10, agent_radius=15,
         agent_max_speed=20, planner="Primitive", gaze="Owl", B=4096, steps=250),
    dict(name="empty_map Jerk_Primitive planner, scripted gaze", static_map="maps/empty_map.npy", agent_number=10, agent_radius=15,
         agent_max_speed=20, planner="Jerk_Primitive", B=4096, steps=250),
    dict(name="random_map_0 N=142 Primitive scripted gaze (config 3)", static_map="maps/random_map_0.npy", agent_number=20,
         agent_radius=15, agent_max_speed=40, planner="Primitive", B=4096, steps=150),
    dict(name="empty_map noisy measurements var_cam=0.5", static_map="maps/empty_map.npy", agent_number=10, agent_radius=15,
         agent_max_speed=20, planner="NoMove", var_cam=0.5, B=4096, steps=100, scatter=True),
    dict(name="empty_map RVO motion profile N=10 + 3 pillars", static_map="maps/empty_map.npy", agent_number=10,
         agent_radius=15, agent_max_speed=20, planner="NoMove", motion_profile="RVO", pillar_number=3, B=2048, steps=300,
         scatter=True),
]


def run_leg(leg, scale, seed0):
    import torch
    import util
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B = max(64, int(leg["B"] * min(1.0, scale)))
    steps = max(10, int(leg["steps"] * max(1.0, scale) if scale >= 1 else leg["steps"] * scale))
    gaze = leg.get("gaze")
    use_ox = gaze in ("Oxford", "Owl")                     # a gaze policy on the device picks the actions
    rvo = leg.get("motion_profile") == "RVO"
    p = Params(debug=False, planner=leg["planner"], gaze_method=gaze if use_ox else "NoControl", map_id=seed0,
               static_map=leg["static_map"], agent_number=leg["agent_number"], agent_radius=leg["agent_radius"],
               agent_max_speed=leg["agent_max_speed"], var_cam=leg.get("var_cam", 0),
               motion_profile=leg.get("motion_profile", "CVM"), pillar_number=leg.get("pillar_number", 0))
    t0 = time.time()
    worlds = generate_worlds(p, seed0 + np.arange(B))
    env = Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", auto_reset=True)     # policy state follows p.gaze_method
    n = env.num_agents
    rng = np.random.RandomState(seed0)
    poses = None
    if leg.get("scatter"):
        poses = worlds["drone_pose"].copy()
        poses[:, 0] = rng.uniform(12, 488, B); poses[:, 1] = rng.uniform(12, 488, B); poses[:, 2] = rng.uniform(0, 360, B)
        lat = rng.rand(B) < 0.25                                   # a quarter on the integer lattice, yaw multiples of 45
        poses[lat, :2] = np.round(poses[lat, :2])
        poses[lat, 2] = 45.0 * rng.randint(0, 8, int(lat.sum()))
        env.set_drone_pose(poses)
        env.buffer("drone_pose0").copy_(torch.as_tensor(poses.T.copy(), device="cuda:0"))
    ob = util.oracle_batch(p, worlds, poses)
    n_way = int(env.cfg.n_way) if leg["planner"] in ("Primitive", "Jerk_Primitive") else 0
    fields = util.BATCH_FIELDS + util.TRACKER_FIELDS + (util.PLANNER_FIELDS if n_way else []) + (["agent_vel"] if rvo else [])
    table = util.action_table()
    alive = np.ones(B, dtype=bool)
    compared = 0
    events = []
    by_field = {}
    max_rel = {}
    episodes = 0
    t_setup = time.time() - t0
    t0 = time.time()
    chunk = int(leg.get("rollout", 0))
    for t in range(steps):
        act_bad = np.zeros(B, dtype=bool)
        if chunk:
            # resident multi-step kernel: `chunk` steps per launch; the oracle steps beside it, compared after the launch
            if t % chunk:
                continue
            acts_k = table[rng.randint(0, 6, (chunk, B))]
            for q in range(chunk):
                ob.step(acts_k[q], auto_reset=True)
                if q < chunk - 1:
                    episodes += int(ob.gather(trackers=False)["done"].sum())
            env.rollout(torch.as_tensor(acts_k, device="cuda:0"))
        elif use_ox:
            a_dev = env.plan_gaze(gaze)
            want, _ = ob.step(policy=gaze, auto_reset=True)
            act_bad = a_dev.cpu().numpy() != want
            a_dev = torch.as_tensor(want, device="cuda:0")        # keep diverged envs from cascading through the action
        else:
            acts = table[rng.randint(0, 6, B)]
            a_dev = torch.as_tensor(acts, device="cuda:0")
            ob.step(acts, auto_reset=True)
        if not chunk:
            env.step(a_dev)
        h = util.gpu_fields(env, fields)
        o = ob.gather(trackers=True, rvo=rvo)
        d, r = util.batch_mismatch(h, o, n, trackers=True, planner=n_way)
        if use_ox:
            d["gaze_action"] = act_bad
        if rvo:
            a, b = h["agent_vel"][:, :n].reshape(B, -1), o["agent_vel"][:, :n].reshape(B, -1)
            r["agent_vel"] = (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max(1)
        compared += int(alive.sum()) * (chunk if chunk else 1)
        bad = np.zeros(B, dtype=bool)
        for k, m in d.items():
            mk = m & alive
            if mk.any():
                by_field[k] = by_field.get(k, 0) + int(mk.sum())
                bad |= mk
        for k, v in r.items():
            va = v[alive]
            if va.size:
                max_rel[k] = max(max_rel.get(k, 0.0), float(va.max()))
            over = (v > 1e-9) & alive
            if over.any():
                by_field[k + ">1e-9"] = by_field.get(k + ">1e-9", 0) + int(over.sum())
                bad |= over
        for i in np.nonzero(bad)[0][:50]:
            events.append({"env": int(i), "step": t, "fields": [k for k, m in d.items() if m[i]] +
                           [k for k, v in r.items() if v[i] > 1e-9]})
        alive &= ~bad
        episodes += int(o["done"].sum())
    dt = time.time() - t0
    st = env.stats()
    ob.close()
    env.close()
    return {"leg": leg["name"], "envs": B, "steps": steps, "env_steps_compared": compared, "episodes_finished": episodes,
            "envs_diverged": int((~alive).sum()), "mismatch_env_steps_by_field": by_field, "first_events": events[:20],
            "max_rel_err": max_rel, "agents_per_env": n, "setup_s": round(t_setup, 1), "run_s": round(dt, 1),
            "gpu_env_steps": int(st[0])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="<1: fewer envs and steps per leg; >1: more steps")
    ap.add_argument("--legs", default="", help="comma-separated leg indices (default all)")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_parity_campaign.json"))
    args = ap.parse_args()
    import oracle
    oracle.build()
    sel = [int(x) for x in args.legs.split(",") if x] or list(range(len(LEGS)))
    res = []
    for j in sel:
        r = run_leg(LEGS[j], args.scale, args.seed + 100003 * j)
        print(json.dumps({k: r[k] for k in ("leg", "env_steps_compared", "envs_diverged", "mismatch_env_steps_by_field",
                                             "run_s")}), flush=True)
        res.append(r)
    tot = sum(r["env_steps_compared"] for r in res)
    div = sum(r["envs_diverged"] for r in res)
    summary = {"what": "CUDA path (default kernels, auto-reset on) vs oracle, every env every step; an env is counted once "
                       "when any discrete output differs or continuous state exceeds 1e-9 relative, then excluded",
               "env_steps_compared": tot, "envs_diverged": div,
               "discrete_mismatch_rate_per_env_step": (div / tot) if tot else None, "legs": res}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(summary, f, indent=1)
    print("TOTAL env-steps compared %d, envs diverged %d" % (tot, div))


if __name__ == "__main__":
    main()
