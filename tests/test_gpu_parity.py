"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the reference's golden outputs.
Bit-exact for grids / masks / flags / integer bookkeeping; continuous state within 1e-9 relative (north_star)."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _env(p, B, worlds, **kw):
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    return Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", **kw)


def _cmp_env_to_oracle(env, i, e, n, t, tag, trackers=True):
    """env: Drone2DVecEnv (host copies passed in `env` dict), e: OracleEnv"""
    h = env
    assert np.array_equal(h["belief"][i], e.belief), (tag, "belief", t, i)
    assert np.array_equal(h["hit"][i, :n], e.hit[:n]), (tag, "hit", t, i)
    assert int(h["collision_flag"][i]) == e.c.collision and int(h["done"][i]) == e.c.done, (tag, "collision/done", t, i)
    assert int(h["dead_lock_flag"][i]) == e.c.dead_lock and int(h["freezing_flag"][i]) == e.c.freezing, (tag, "flags", t, i)
    assert int(h["state_machine"][i]) == e.c.state_machine and int(h["fail_count"][i]) == e.c.fail_count, (tag, "sm", t, i)
    assert int(h["steps"][i]) == e.c.steps
    assert np.array_equal(h["local_map"][i, 0], e.local_map), (tag, "local_map", t, i)
    assert h["yaw_angle"][i, 0] == np.float32(e.c.yaw_obs), (tag, "yaw obs", t, i)
    for name, ref in (("drone_x", e.c.x), ("drone_y", e.c.y), ("drone_yaw", e.c.yaw), ("drone_vx", e.c.vx), ("drone_vy", e.c.vy)):
        assert util.rel_err(h[name][i], ref) <= RTOL, (tag, name, t, i, h[name][i], ref)
    assert util.rel_err(h["agent_pos"][i, :n], e.apos[:n]) <= RTOL, (tag, "agent_pos", t, i)
    assert util.rel_err(h["agent_pref"][i, :n], e.apref[:n]) <= RTOL, (tag, "agent_pref", t, i)
    if trackers:
        act = e.trk_active[:n].astype(bool)
        assert np.array_equal(h["tracker_active"][i, :n].astype(bool), act), (tag, "trk_active", t, i)
        assert np.array_equal(h["tracker_ts"][i, :n][act], e.trk_ts[:n][act]), (tag, "trk_ts", t, i)
        assert np.array_equal(h["tracker_radius"][i, :n], e.trk_radius[:n]), (tag, "trk_radius", t, i)
        if act.any():
            assert util.rel_err(h["tracker_mu"][i, :n][act], e.trk_mu[:n][act]) <= RTOL, (tag, "trk_mu", t, i)
            assert util.rel_err(h["tracker_sigma"][i, :n][act].reshape(-1, 16), e.trk_sigma[:n][act]) <= RTOL, (tag, "trk_sigma", t, i)
        assert int(h["tracker_buffer_count"][i]) == e.c.buf_count and int(h["tracker_buffer_ts"][i]) == e.c.buf_ts, (tag, "buffer", t, i)


FIELDS = ["belief", "hit", "collision_flag", "done", "dead_lock_flag", "freezing_flag", "state_machine", "fail_count", "steps",
          "local_map", "yaw_angle", "drone_x", "drone_y", "drone_yaw", "drone_vx", "drone_vy", "agent_pos", "agent_pref",
          "tracker_active", "tracker_ts", "tracker_radius", "tracker_mu", "tracker_sigma", "tracker_buffer_count",
          "tracker_buffer_ts", "traj_nseg", "traj_cursor", "plan_ok", "replan"]


def _host(env):
    torch.cuda.synchronize()
    return {k: env.buffer(k).cpu().numpy() for k in FIELDS}


@pytest.mark.parametrize("path", util.golden_files("nomove_"), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_reference_golden_nomove(path):
    """CUDA step vs the reference's recorded outputs (perception + dynamics + trackers, NoMove planner)."""
    g = util.load_golden(path)
    p = util.params_from_golden(g)
    n = int(g["n_agents"])
    B = 5   # identical copies; B not a multiple of envs_per_block exercises the ragged last block
    env = _env(p, B, util.world_from_golden(g, B), auto_reset=False)
    init = np.array([float(p.init_position[0]), float(p.init_position[1]), 270.0])
    if not np.array_equal(g["drone0"], init):
        env.set_drone_pose(np.stack([g["drone0"]] * B))
    T = len(g["done"])
    first_done = T
    for t in range(T):
        a = torch.full((B,), float(g["action"][t]), dtype=torch.float64, device="cuda:0")
        obs, rew, done, info = env.step(a)
        h = _host(env)
        for i in (0, B - 1):
            assert np.array_equal(h["belief"][i], g["belief"][t]), ("belief", t)
            assert np.array_equal(h["hit"][i, :n], g["hit"][t]), ("hit", t)
            assert h["collision_flag"][i] == g["collision"][t] and bool(h["done"][i]) == bool(g["done"][t]), ("done", t)
            assert np.array_equal(h["local_map"][i, 0], g["local_map"][t]), ("local_map", t)
            assert h["yaw_angle"][i, 0] == g["yaw_obs"][t]
            assert util.rel_err([h["drone_x"][i], h["drone_y"][i], h["drone_yaw"][i]], g["drone"][t]) <= RTOL
            assert util.rel_err(h["agent_pos"][i, :n], g["agent_pos"][t]) <= RTOL, ("agent_pos", t)
            assert util.rel_err(h["agent_pref"][i, :n], g["agent_pref"][t]) <= RTOL
            assert h["state_machine"][i] == g["state_machine"][t]
            if "trk_active" in g:
                act = g["trk_active"][t]
                assert np.array_equal(h["tracker_active"][i, :n].astype(bool), act), ("trk_active", t)
                if act.any():
                    assert util.rel_err(h["tracker_mu"][i, :n][act], g["trk_mu"][t][act]) <= RTOL
                    assert util.rel_err(h["tracker_sigma"][i, :n][act].reshape(-1, 16), g["trk_sigma"][t][act]) <= RTOL
                    assert np.array_equal(h["tracker_ts"][i, :n][act], g["trk_ts"][t][act])
                if t <= first_done:
                    assert (h["tracker_buffer_count"][i], h["tracker_buffer_ts"][i]) == (g["buf_count"][t], g["buf_ts"][t])
        assert float(rew.abs().sum()) == 0.0
        if g["done"][t] and first_done == T:
            first_done = t
    env.close()


@pytest.mark.parametrize("cfg", [
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, B=96, steps=160, epb=8),
    dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20, B=40, steps=100, epb=4),
    dict(static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15, agent_max_speed=40, B=24, steps=40, epb=16),
    dict(static_map="maps/shaped_obstacle_map.npy", agent_number=50, agent_radius=10, agent_max_speed=40, B=20, steps=40, epb=8),
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, B=7, steps=60, epb=8),
    dict(static_map="maps/empty_map.npy", agent_number=6, agent_radius=10, agent_max_speed=40, B=11, steps=60, epb=4),
    dict(static_map="maps/empty_map.npy", agent_number=6, agent_radius=10, agent_max_speed=40, B=18, steps=40, epb=16),
    dict(static_map="maps/empty_map.npy", agent_number=3, agent_radius=10, agent_max_speed=40, B=1, steps=40, epb=8),
    # the same seeded batches on the DEFAULT kernel (envs_per_block = 0: d2d_step_fused_warp_kernel, the one bench.py times)
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, B=96, steps=160, epb=0),
    dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20, B=40, steps=100, epb=0),
    dict(static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15, agent_max_speed=40, B=24, steps=40, epb=0),
    dict(static_map="maps/shaped_obstacle_map.npy", agent_number=50, agent_radius=10, agent_max_speed=40, B=20, steps=40, epb=0),
    dict(static_map="maps/empty_map.npy", agent_number=6, agent_radius=10, agent_max_speed=40, B=11, steps=60, epb=0),
    dict(static_map="maps/empty_map.npy", agent_number=3, agent_radius=10, agent_max_speed=40, B=1, steps=40, epb=0),
], ids=["cfg2_empty", "cfg4_obstacle", "cfg3_random0", "cfg5_shaped", "ragged7", "ragged11", "ragged18", "single",
        "cfg2_empty_warp", "cfg4_obstacle_warp", "cfg3_random0_warp", "cfg5_shaped_warp", "ragged11_warp", "single_warp"])
def test_cuda_matches_oracle_batch_nomove(cfg):
    """Seeded batch (world generation on the host, random actions from the Oxford action set) stepped by the CUDA
    path and by the oracle; every field compared every step; envs keep stepping after done (auto_reset off)."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = cfg["B"], cfg["steps"]
    p = Params(debug=False, planner="NoMove", map_id=100, static_map=cfg["static_map"], agent_number=cfg["agent_number"],
               agent_radius=cfg["agent_radius"], agent_max_speed=cfg["agent_max_speed"])
    seeds = 100 + np.arange(B)
    worlds = generate_worlds(p, seeds)
    env = _env(p, B, worlds, auto_reset=False, envs_per_block=cfg["epb"])
    n = env.num_agents
    # scatter drones over the map (non-integer poses, some next to walls) like the metric scripts do
    rng = np.random.RandomState(7)
    poses = worlds["drone_pose"].copy()
    poses[B // 4:, 0] = rng.uniform(15, 485, B - B // 4)
    poses[B // 4:, 1] = rng.uniform(15, 485, B - B // 4)
    poses[B // 4:, 2] = rng.uniform(0, 360, B - B // 4)
    env.set_drone_pose(poses)
    oracles = [util.oracle_env_from_world(p, worlds, i, drone=poses[i]) for i in range(B)]
    table = util.action_table()
    for t in range(steps):
        acts = table[rng.randint(0, 6, B)]
        env.step(torch.as_tensor(acts, device="cuda:0"))
        for i, e in enumerate(oracles):
            e.step(float(acts[i]))
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, cfg["static_map"])
    st = env.stats()
    assert st[0] == B * steps
    for e in oracles:
        e.close()
    env.close()


def test_auto_reset_restores_initial_world():
    """done -> next step starts from the initial snapshot (reference reset() re-runs __init__ with the same seed)."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = 32, 200
    p = Params(debug=False, planner="NoMove", map_id=300, agent_number=14, agent_radius=15, agent_max_speed=40,
               init_pos=[250, 250])
    worlds = generate_worlds(p, 300 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True)
    n = env.num_agents
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    rng = np.random.RandomState(3)
    table = util.action_table()
    resets = 0
    for t in range(steps):
        acts = table[rng.randint(0, 6, B)]
        for i in range(B):
            if oracles[i].c.done:
                oracles[i].close()
                oracles[i] = util.oracle_env_from_world(p, worlds, i)
                resets += 1
            oracles[i].step(float(acts[i]))
        env.step(torch.as_tensor(acts, device="cuda:0"))
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, "auto_reset")
    assert resets > 0
    st = env.stats()
    assert st[1] >= resets and st[1] == st[3] + st[4] + st[5] + st[6] + st[2]
    env.close()


def test_step_host_buffers_and_explicit_reset():
    from gym_drone2d_activeperception_b200.params import Params
    B = 16
    p = Params(debug=False, planner="NoMove", map_id=5, agent_number=10, agent_radius=15, agent_max_speed=20)
    env = _env(p, B, None, auto_reset=False)
    acts = torch.full((B,), 1.0, dtype=torch.float64).pin_memory()
    lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory()
    yaw = torch.empty((B,), dtype=torch.float32).pin_memory()
    dn = torch.empty((B,), dtype=torch.uint8).pin_memory()
    for _ in range(5):
        env.step_host(acts, lm, yaw, dn)
    assert torch.equal(lm, env.buffer("local_map").cpu()) and torch.equal(yaw, env.buffer("yaw_angle").cpu()[:, 0])
    assert float(yaw[0]) == np.float32(310.0) and int(lm.sum()) > 0
    obs = env.reset()
    torch.cuda.synchronize()
    assert int(obs["local_map"].sum()) == 0 and float(obs["yaw_angle"][0, 0]) == 270.0
    assert int(env.buffer("steps").sum()) == 0 and int(env.buffer("belief").sum()) == 0
    env.close()


@pytest.mark.parametrize("case", [
    dict(name="no_agents", kw=dict(agent_number=0), B=6, steps=40),
    dict(name="pillars", kw=dict(agent_number=6, pillar_number=4, agent_radius=10, agent_max_speed=40), B=10, steps=60),
    dict(name="wide_fov_dense_rays", kw=dict(agent_number=8, drone_view_range=180), B=6, steps=40, strip_width=5),
    dict(name="full_circle_250_rays", kw=dict(agent_number=8, drone_view_range=360), B=5, steps=30, strip_width=2),
    dict(name="random_radius", kw=dict(agent_number=12, agent_radius=-1, agent_max_speed=60), B=9, steps=60),
    dict(name="many_agents_radius5", kw=dict(agent_number=64, agent_radius=5, agent_max_speed=30), B=5, steps=40),
    dict(name="culled_list_over_32", kw=dict(agent_number=400, agent_radius=5, agent_max_speed=30), B=6, steps=30),
], ids=lambda c: c["name"])
def test_cuda_matches_oracle_edge_configs(case):
    """Edge configurations of the reference's Params: empty agent list, static pillars (per-env ground truth),
    wider FOV and denser ray fans (BASELINE config 5 sweeps strip_width 10/5/2 and view_range 90/180/360),
    `agent_radius == -1` (uniform 5..15 radii, drone_v2.py:31), crowded worlds."""
    import oracle
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = case["B"], case["steps"]
    sw = case.get("strip_width", 10)
    p = Params(debug=False, planner="NoMove", map_id=900, **case["kw"])
    worlds = generate_worlds(p, 900 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=False, strip_width=sw)
    n = env.num_agents
    rng = np.random.RandomState(11)
    poses = worlds["drone_pose"].copy()
    poses[1:, 0] = rng.uniform(15, 485, B - 1)
    poses[1:, 1] = rng.uniform(15, 485, B - 1)
    poses[1:, 2] = rng.uniform(0, 360, B - 1)
    env.set_drone_pose(poses)
    op = util.oracle_params(p)
    op.n_rays = int(np.ceil(p.map_size[0] / sw))
    oracles = [oracle.OracleEnv(op, worlds["agent_pos"][i], worlds["agent_pref"][i], worlds["agent_radius"][i],
                                worlds["gt_grid"][i], worlds["tracker_radius"][i], drone=poses[i], targets=p.target_list)
               for i in range(B)]
    table = util.action_table()
    for t in range(steps):
        acts = table[rng.randint(0, 6, B)]
        env.step(torch.as_tensor(acts, device="cuda:0"))
        for i, e in enumerate(oracles):
            e.step(float(acts[i]))
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, case["name"])
    for e in oracles:
        e.close()
    env.close()


def test_error_paths_are_loud():
    """Unsupported configurations and misuse raise instead of silently falling back."""
    from gym_drone2d_activeperception_b200 import _native
    from gym_drone2d_activeperception_b200.params import Params
    with pytest.raises(_native.Drone2DNativeError):      # noisy measurements exist only on the default warp-per-env kernels
        _env(Params(debug=False, planner="NoMove", var_cam=1), 4, None, envs_per_block=8)
    noisy = _env(Params(debug=False, planner="NoMove", var_cam=1, agent_number=3), 4, None)
    w = {k: v for k, v in __import__("gym_drone2d_activeperception_b200").generate_worlds(noisy.params, [0, 1, 2, 3]).items()
         if not k.startswith("rng_")}
    noisy2 = _env(Params(debug=False, planner="NoMove", var_cam=1, agent_number=3), 4, w)      # worlds without the RNG state
    with pytest.raises(_native.Drone2DNativeError):
        noisy2.step(torch.zeros(4, dtype=torch.float64, device="cuda:0"))
    noisy.close()
    noisy2.close()
    with pytest.raises(ValueError):
        _env(Params(debug=False, planner="MPC"), 4, None)
    with pytest.raises(ValueError):
        _env(Params(debug=False, planner="NoMove", motion_profile="ORCA"), 4, None)
    env = _env(Params(debug=False, planner="NoMove", agent_number=3), 4, None, oxford=False)
    with pytest.raises(_native.Drone2DNativeError):
        env.plan_oxford()
    with pytest.raises(ValueError):
        env.step(torch.zeros(3, dtype=torch.float64, device="cuda:0"))
    env.close()


def test_trackers_disabled_and_primitive_step_host():
    """cfg.trackers = 0 skips the Kalman filters only (perception, flags and observation unchanged); the host-buffer
    entry point drives the Primitive path as well."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B = 10
    p = Params(debug=False, planner="NoMove", map_id=60, agent_number=10, agent_radius=15, agent_max_speed=20)
    worlds = generate_worlds(p, 60 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=False, trackers=False)
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    for t in range(40):
        env.step(torch.full((B,), -1.0 / 3, dtype=torch.float64, device="cuda:0"))
        for e in oracles:
            e.step(-1.0 / 3)
        torch.cuda.synchronize()
        bel, hit = env.buffer("belief").cpu().numpy(), env.buffer("hit").cpu().numpy()
        lm, dn = env.buffer("local_map").cpu().numpy(), env.buffer("done").cpu().numpy()
        for i, e in enumerate(oracles):
            assert np.array_equal(bel[i], e.belief) and np.array_equal(hit[i], e.hit) and int(dn[i]) == e.c.done, (t, i)
            assert np.array_equal(lm[i, 0], e.local_map)
    assert int(env.buffer("tracker_active").sum()) == 0
    env.close()

    pp = Params(debug=False, planner="Primitive", map_id=61, agent_number=8, agent_radius=15, agent_max_speed=20)
    w2 = generate_worlds(pp, 61 + np.arange(B))
    env = _env(pp, B, w2, auto_reset=False, oxford=False)
    oracles = [util.oracle_env_from_world(pp, w2, i) for i in range(B)]
    acts = torch.full((B,), 2.0 / 3, dtype=torch.float64).pin_memory()
    lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory()
    yaw = torch.empty((B,), dtype=torch.float32).pin_memory()
    dn = torch.empty((B,), dtype=torch.uint8).pin_memory()
    for t in range(60):
        env.step_host(acts, lm, yaw, dn)
        for i, e in enumerate(oracles):
            if e.c.done:
                continue
            e.step(2.0 / 3)
            assert np.array_equal(lm[i, 0].numpy(), e.local_map) and int(dn[i]) == e.c.done, (t, i)
            assert float(yaw[i]) == float(np.float32(e.c.yaw_obs))
    env.close()


def test_cuda_matches_oracle_awkward_poses():
    """Drone poses the belief-first / general ray marches must treat exactly like the reference loop `while 0 < x < W and
    0 < y < H` (utils.py:654): inside border wall cells, exactly on x == 0 / y == 0, exactly on cell boundaries and on
    the far map edge, and outside the map; yaws on the axes and diagonals (rays running along cell boundaries)."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    xy = [(5.0, 5.0), (0.0, 250.0), (250.0, 0.0), (0.0, 0.0), (10.0, 10.0), (20.0, 490.0), (495.0, 495.0), (499.999, 250.0),
          (500.0, 250.0), (-5.0, 100.0), (100.0, 505.0), (250.0, 250.0), (30.0, 30.0), (470.0, 20.0), (9.999999999, 40.0),
          (490.0, 250.0), (40.0, 0.0), (250.0, 489.99999)]
    yaws = [0.0, 90.0, 180.0, 270.0, 45.0, 135.0, 225.0, 315.0]
    poses = np.array([(x, y, yw) for (x, y) in xy for yw in yaws], dtype=np.float64)
    B, steps = len(poses), 24
    p = Params(debug=False, planner="NoMove", map_id=300, agent_number=12, agent_radius=15, agent_max_speed=40)
    worlds = generate_worlds(p, 300 + np.arange(B) % 7)
    env = _env(p, B, worlds, auto_reset=False)
    n = env.num_agents
    env.set_drone_pose(poses)
    oracles = [util.oracle_env_from_world(p, worlds, i, drone=poses[i]) for i in range(B)]
    table = util.action_table()
    rng = np.random.RandomState(5)
    for t in range(steps):
        acts = table[rng.randint(0, 6, B)] if t % 3 else np.zeros(B)      # zero action: the yaw stays on the axis / diagonal
        env.step(torch.as_tensor(acts, device="cuda:0"))
        for i, e in enumerate(oracles):
            e.step(float(acts[i]))
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, "pose %s" % (poses[i],))
    for e in oracles:
        e.close()
    env.close()
