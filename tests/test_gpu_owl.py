"""Owl gaze policy on the device (d2d_owl_kernel behind d2d_plan_gaze) against the reference's own episodes
(tests/golden/owl_*.npz, generated with the reference's real call pattern: the class object is the instance,
experiment.py:33-34) and against the oracle on seeded batches with auto-reset."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(p, B, worlds, **kw):
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    return Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", **kw)


@pytest.mark.parametrize("path", util.golden_files("owl_"), ids=lambda p: p.split("/")[-1][:-4])
def test_owl_actions_match_reference_episode(path):
    """Every action of whole Primitive-planner episodes equals the reference's, and so does everything downstream of it."""
    g = util.load_golden(path)
    p = util.params_from_golden(g)
    n, B = int(g["n_agents"]), 3
    env = _env(p, B, util.world_from_golden(g, B), auto_reset=False, owl=True)
    for t in range(len(g["done"])):
        a = env.plan_gaze("Owl")
        torch.cuda.synchronize()
        got = a.cpu().numpy()
        assert np.all(got == g["action"][t]), ("owl action", t, got, g["action"][t])
        _, _, done, _ = env.step(a)
        bel = env.buffer("belief").cpu().numpy()
        for i in (0, B - 1):
            assert np.array_equal(bel[i], g["belief"][t]), ("belief", t)
            assert bool(done[i]) == bool(g["done"][t]) and int(env.buffer("collision_flag")[i]) == g["collision"][t]
        assert util.rel_err(env.buffer("drone_yaw").cpu().numpy()[0], g["drone"][t][2]) <= 1e-9
    env.close()


@pytest.mark.parametrize("cfg", [
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, drone_max_speed=40, B=96, steps=320),
    dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20, drone_max_speed=40, B=64, steps=260),
    dict(static_map="maps/empty_map.npy", agent_number=40, agent_radius=10, agent_max_speed=40, drone_max_speed=20, B=48, steps=200),
], ids=["empty", "obstacle", "crowded_speed20"])
def test_owl_batch_matches_oracle_with_auto_reset(cfg):
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = cfg["B"], cfg["steps"]
    p = Params(debug=False, planner="Primitive", gaze_method="Owl", map_id=900, static_map=cfg["static_map"],
               agent_number=cfg["agent_number"], agent_radius=cfg["agent_radius"], agent_max_speed=cfg["agent_max_speed"],
               drone_max_speed=cfg["drone_max_speed"])
    worlds = generate_worlds(p, 900 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True)                       # owl state follows params.gaze_method
    n = env.num_agents
    ob = util.oracle_batch(p, worlds)
    fields = util.BATCH_FIELDS + util.PLANNER_FIELDS + util.TRACKER_FIELDS
    episodes = 0
    for t in range(steps):
        a = env.plan_gaze("Owl")
        want, _ = ob.step(policy="Owl", auto_reset=True)
        got = a.cpu().numpy()
        assert np.array_equal(got, want), ("owl action", t, np.nonzero(got != want)[0][:8], got[got != want][:4], want[got != want][:4])
        env.step(a)
        h = util.gpu_fields(env, fields + ["owl_U", "owl_queue_len"])
        o = ob.gather(trackers=True)
        d, r = util.batch_mismatch(h, o, n, trackers=True, planner=int(env.cfg.n_way))
        for k, m in d.items():
            assert not m.any(), (k, t, np.nonzero(m)[0][:8])
        for k, v in r.items():
            assert float(v.max()) <= 1e-9, (k, t, float(v.max()))
        oU = np.array([np.array(e.c.owl_U[:]) for e in ob.envs])
        oq = np.array([e.c.owl_q for e in ob.envs])
        assert np.array_equal(h["owl_U"], oU) and np.array_equal(h["owl_queue_len"], oq), ("owl state", t)
        episodes += int(o["done"].sum())
    assert episodes > B // 2, "the case must exercise the policy reset at episode boundaries"
    ob.close()
    env.close()


def test_owl_needs_its_state():
    from gym_drone2d_activeperception_b200 import _native
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    p = Params(debug=False, planner="Primitive", gaze_method="LookAhead", map_id=3, agent_number=4)
    env = _env(p, 4, generate_worlds(p, 3 + np.arange(4)))
    with pytest.raises(_native.Drone2DNativeError):
        env.plan_gaze("Owl")
    env.close()
