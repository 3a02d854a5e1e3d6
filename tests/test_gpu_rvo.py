"""RVO motion profile on the device (d2d_rvo_kernel + Agent.step with a velocity of its own; utils.py:299-460, 472-493,
drone_v2.py:169-175).  The reference / oracle compute atan2, asin, sin, cos with glibc, the kernel with CUDA's double
precision functions (<= 2 ulp), so continuous agent state is held to the north star's 1e-9 relative tolerance; the discrete
outputs (belief, hit lists, flags, done, observation) must still agree exactly on these seeds."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _env(p, B, worlds, **kw):
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    return Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", **kw)


@pytest.mark.parametrize("path", util.golden_files("rvo_"), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_rvo_matches_reference_golden(path):
    g = util.load_golden(path)
    p = util.params_from_golden(g)
    n = int(g["n_agents"])
    B = 3
    env = _env(p, B, util.world_from_golden(g, B), auto_reset=False, oxford=False)
    init = np.array([float(p.init_position[0]), float(p.init_position[1]), 270.0])
    if not np.array_equal(g["drone0"], init):
        env.set_drone_pose(np.stack([g["drone0"]] * B))
    T = len(g["done"])
    for t in range(T):
        a = torch.full((B,), float(g["action"][t]), dtype=torch.float64, device="cuda:0")
        env.step(a)
        torch.cuda.synchronize()
        pos, vel, pref = (env.buffer(k).cpu().numpy() for k in ("agent_pos", "agent_vel", "agent_pref"))
        bel, hit, lm = (env.buffer(k).cpu().numpy() for k in ("belief", "hit", "local_map"))
        dn, col = env.buffer("done").cpu().numpy(), env.buffer("collision_flag").cpu().numpy()
        for i in (0, B - 1):
            assert util.rel_err(pos[i, :n], g["agent_pos"][t]) <= RTOL, ("agent_pos", t)
            assert util.rel_err(vel[i, :n], g["agent_vel"][t]) <= RTOL, ("agent_vel", t)
            assert util.rel_err(pref[i, :n], g["agent_pref"][t]) <= RTOL, ("agent_pref", t)
            assert np.array_equal(bel[i], g["belief"][t]) and np.array_equal(hit[i, :n], g["hit"][t]), ("belief/hit", t)
            assert np.array_equal(lm[i, 0], g["local_map"][t]), ("local_map", t)
            assert bool(dn[i]) == bool(g["done"][t]) and int(col[i]) == int(g["collision"][t]), ("done", t)
            assert util.rel_err([float(env.buffer("drone_x")[i]), float(env.buffer("drone_y")[i])], g["drone"][t][:2]) <= RTOL
    env.close()


@pytest.mark.parametrize("cfg", [
    dict(planner="NoMove", kw=dict(agent_number=10, agent_radius=15, agent_max_speed=20), B=9, steps=80, epb=0),
    dict(planner="NoMove", kw=dict(agent_number=40, agent_radius=15, agent_max_speed=20), B=4, steps=30, epb=0),
    dict(planner="NoMove", kw=dict(agent_number=12, pillar_number=5, agent_max_speed=40), B=6, steps=60, epb=8),
    dict(planner="Primitive", kw=dict(agent_number=8, pillar_number=3, agent_radius=12), B=7, steps=150, epb=0),
    dict(planner="NoMove", kw=dict(agent_number=6, static_map="maps/obstacle_map.npy"), B=3, steps=30, epb=0),
], ids=["cfg2_like", "crowd_fallback_branch", "pillars_block_kernel", "primitive_auto_reset", "map_agents"])
def test_cuda_rvo_matches_oracle_batch(cfg):
    """Seeded batch under RVO vs the oracle, with auto-reset (the RVO kernel must plan from the snapshot of an env that the
    same step re-initialises) and both kernel families."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = cfg["B"], cfg["steps"]
    p = Params(debug=False, planner=cfg["planner"], motion_profile="RVO", map_id=40, **cfg["kw"])
    worlds = generate_worlds(p, 40 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True, oxford=False, envs_per_block=cfg["epb"])
    n = env.num_agents
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    table = util.action_table()
    rng = np.random.RandomState(3)
    resets = 0
    for t in range(steps):
        acts = table[rng.randint(0, 6, B)]
        for i in range(B):      # auto-reset twin: a done oracle env is rebuilt from its world before its next step
            if oracles[i].c.done:
                oracles[i].close()
                oracles[i] = util.oracle_env_from_world(p, worlds, i)
                resets += 1
        env.step(torch.as_tensor(acts, device="cuda:0"))
        for i, e in enumerate(oracles):
            e.step(float(acts[i]))
        torch.cuda.synchronize()
        pos, vel, pref = (env.buffer(k).cpu().numpy() for k in ("agent_pos", "agent_vel", "agent_pref"))
        bel, hit, lm = (env.buffer(k).cpu().numpy() for k in ("belief", "hit", "local_map"))
        dn = env.buffer("done").cpu().numpy()
        dx, dy = env.buffer("drone_x").cpu().numpy(), env.buffer("drone_y").cpu().numpy()
        for i, e in enumerate(oracles):
            assert util.rel_err(pos[i, :n], e.apos[:n]) <= RTOL, ("agent_pos", t, i)
            assert util.rel_err(vel[i, :n], e.avel[:n]) <= RTOL, ("agent_vel", t, i)
            assert util.rel_err(pref[i, :n], e.apref[:n]) <= RTOL, ("agent_pref", t, i)
            assert np.array_equal(bel[i], e.belief) and np.array_equal(hit[i, :n], e.hit[:n]), ("belief/hit", t, i)
            assert np.array_equal(lm[i, 0], e.local_map) and int(dn[i]) == e.c.done, ("obs/done", t, i)
            assert util.rel_err([dx[i], dy[i]], [e.c.x, e.c.y]) <= RTOL, ("drone", t, i)
    if cfg["planner"] == "Primitive":
        assert resets > 0
    if "crowd" in str(cfg["kw"].get("agent_number")) or cfg["kw"]["agent_number"] == 40:
        assert sum(int(e.c.rvo_fallbacks) for e in oracles) > 0
    for e in oracles:
        e.close()
    env.close()


def test_rvo_needs_its_world_arrays():
    from gym_drone2d_activeperception_b200 import _native
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    p = Params(debug=False, planner="NoMove", motion_profile="RVO", agent_number=4)
    w = {k: v for k, v in generate_worlds(p, [1, 2]).items() if k not in ("agent_vel", "obstacles")}
    with pytest.raises(ValueError):
        _env(p, 2, w)
    with pytest.raises(ValueError):      # more pillars than the device table holds
        _env(Params(debug=False, planner="NoMove", motion_profile="RVO", agent_number=4, pillar_number=17), 1, None)
