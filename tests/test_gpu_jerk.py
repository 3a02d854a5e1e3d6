"""Jerk_Primitive planner on the device (d2d_step_jerk_warp_kernel) against the reference's own episodes
(tests/golden/jerk_*.npz, incl. the canonical scenario whose goal bearing of exactly 90 degrees makes every pair of headings
tie in numpy's unstable argsort) and against the oracle on seeded batches with auto-reset.  The tie orders are the ones
recorded on the machine that generated the fixtures, handed to both sides."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(p, B, worlds, **kw):
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    return Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", **kw)


def _tie_orders():
    return util.load_golden(util.golden_files("jerk_s3")[0])["jerk_tie_orders"]


@pytest.mark.parametrize("path", util.golden_files("jerk_"), ids=lambda p: p.split("/")[-1][:-4])
def test_jerk_matches_reference_episode(path):
    g = util.load_golden(path)
    p = util.params_from_golden(g)
    n, B = int(g["n_agents"]), 3
    env = _env(p, B, util.world_from_golden(g, B), auto_reset=False, jerk_tie_orders=g["jerk_tie_orders"])
    for t in range(len(g["done"])):
        a = torch.full((B,), float(g["action"][t]), dtype=torch.float64, device="cuda:0")
        _, _, done, _ = env.step(a)
        h = util.gpu_fields(env, ["belief", "drone_x", "drone_y", "drone_yaw", "drone_vx", "drone_vy", "collision_flag", "done",
                                  "state_machine", "fail_count", "plan_ok", "local_map", "agent_pos"])
        for i in (0, B - 1):
            assert (h["drone_x"][i], h["drone_y"][i]) == tuple(g["drone"][t][:2]), ("position", t, h["drone_x"][i], h["drone_y"][i], g["drone"][t])
            assert util.rel_err([h["drone_yaw"][i], h["drone_vx"][i], h["drone_vy"][i]], [g["drone"][t][2], *g["drone_vel"][t]]) <= 1e-9, ("state", t)
            assert np.array_equal(h["belief"][i], g["belief"][t]) and np.array_equal(h["local_map"][i, 0], g["local_map"][t]), ("grids", t)
            assert bool(h["done"][i]) == bool(g["done"][t]) and h["collision_flag"][i] == g["collision"][t], ("done", t)
            assert h["state_machine"][i] == g["state_machine"][t] and h["fail_count"][i] == g["fail_count"][t], ("sm", t)
            assert bool(h["plan_ok"][i]) == bool(g["plan_ok"][t]), ("plan_ok", t)
            assert util.rel_err(h["agent_pos"][i, :n], g["agent_pos"][t]) <= 1e-9
    env.close()


@pytest.mark.parametrize("cfg", [
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, drone_max_speed=40, B=128, steps=260, policy="LookAhead"),
    dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20, drone_max_speed=40, B=64, steps=220, policy="scripted"),
    dict(static_map="maps/empty_map.npy", agent_number=40, agent_radius=10, agent_max_speed=40, drone_max_speed=20, B=48, steps=200, policy="Oxford"),
], ids=["empty_lookahead", "obstacle_scripted", "crowded_speed20_oxford"])
def test_jerk_batch_matches_oracle_with_auto_reset(cfg):
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps, pol = cfg["B"], cfg["steps"], cfg["policy"]
    p = Params(debug=False, planner="Jerk_Primitive", gaze_method=pol if pol != "scripted" else "NoControl", map_id=300,
               static_map=cfg["static_map"], agent_number=cfg["agent_number"], agent_radius=cfg["agent_radius"],
               agent_max_speed=cfg["agent_max_speed"], drone_max_speed=cfg["drone_max_speed"])
    ties = _tie_orders()
    worlds = generate_worlds(p, 300 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True, jerk_tie_orders=ties)
    n = env.num_agents
    ob = util.oracle_batch(p, worlds, jerk_tie_orders=ties)
    fields = util.BATCH_FIELDS + util.TRACKER_FIELDS + ["plan_ok", "replan", "traj_nseg", "traj_cursor", "drone_acc"]
    table, rng = util.action_table(), np.random.RandomState(3)
    episodes = fails = 0
    for t in range(steps):
        if pol == "scripted":
            acts = table[rng.randint(0, 6, B)]
            a = torch.as_tensor(acts, device="cuda:0")
            ob.step(acts, auto_reset=True)
        else:
            a = env.plan_gaze(pol)
            want, _ = ob.step(policy=pol, auto_reset=True)
            got = a.cpu().numpy()
            assert util.rel_err(got, want) <= 1e-12, ("gaze action", t, np.nonzero(got != want)[0][:8])
            a = torch.as_tensor(want, device="cuda:0")
        env.step(a)
        h = util.gpu_fields(env, fields)
        o = ob.gather(trackers=True)
        d, r = util.batch_mismatch(h, o, n, trackers=True, planner=int(env.cfg.n_way))
        for k, m in d.items():
            assert not m.any(), (k, t, np.nonzero(m)[0][:8])
        for k, v in r.items():
            assert float(v.max()) <= 1e-9, (k, t, float(v.max()))
        oacc = np.array([[e.c.ax, e.c.ay] for e in ob.envs])
        assert util.rel_err(h["drone_acc"], oacc) <= 1e-9, ("acceleration", t)
        episodes += int(o["done"].sum())
        fails += int((o["plan_ok"] == 0).sum())
    assert episodes > B // 2
    ob.close()
    env.close()
