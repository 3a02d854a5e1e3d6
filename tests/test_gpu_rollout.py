"""d2d_rollout (K steps of every env in ONE launch, env state resident on chip) against K d2d_step launches on a twin env:
every state / observation buffer and the statistics must be BIT-identical after each chunk, for chunk lengths 1, 7 and 33,
with auto-reset inside a chunk (scattered poses), more agents than lanes (N = 96) and noisy measurements (per-env RNG
stream in HBM).  A second test holds the rollout to the oracle directly (oracle.OracleBatch stepped beside it)."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ALL_FIELDS = util.BATCH_FIELDS + util.TRACKER_FIELDS + ["env_records"]     # env_records: every persistent scalar of an env

CASES = {
    "cfg2_scattered": dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, B=1024,
                           scatter=True),
    "cfg5_n96": dict(static_map="maps/shaped_obstacle_map.npy", agent_number=50, agent_radius=10, agent_max_speed=40, B=512,
                     scatter=True),
    "cfg2_noise": dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, B=256,
                       var_cam=0.5, scatter=True),
    # above two waves of warps d2d_rollout steps with K per-step launches: same contract
    "large_batch_dispatch": dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, B=8400,
                                 scatter=True, worlds=256),
    "obstacle_start_pose": dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20, B=512),
}


def _make(cfg, n_env=None):
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B = cfg["B"]
    p = Params(debug=False, planner="NoMove", gaze_method="NoControl", map_id=1, static_map=cfg["static_map"],
               agent_number=cfg["agent_number"], agent_radius=cfg["agent_radius"], agent_max_speed=cfg["agent_max_speed"],
               var_cam=cfg.get("var_cam", 0.0))
    nw = min(B, cfg.get("worlds", B))
    worlds = generate_worlds(p, 1 + np.arange(nw))
    if nw < B:
        worlds = {k: np.concatenate([v] * (-(-B // nw)))[:B] for k, v in worlds.items()}
    poses = None
    if cfg.get("scatter"):
        rng = np.random.RandomState(5)
        poses = worlds["drone_pose"].copy()
        poses[:, 0] = rng.uniform(12, 488, B); poses[:, 1] = rng.uniform(12, 488, B); poses[:, 2] = rng.uniform(0, 360, B)
        poses[::5, :2] = np.round(poses[::5, :2])

    def mk():
        env = Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", auto_reset=True)
        if poses is not None:
            env.set_drone_pose(poses)
            env.buffer("drone_pose0").copy_(torch.as_tensor(poses.T.copy(), device="cuda:0"))
        return env
    return p, worlds, poses, mk


def _fields(env, noisy):
    names = ALL_FIELDS + (["rng_key", "rng_pos", "rng_gauss"] if noisy else [])
    return {k: env.buffer(k).cpu().numpy() for k in names}


@pytest.mark.parametrize("name", list(CASES))
def test_rollout_equals_single_steps(name):
    cfg = CASES[name]
    B = cfg["B"]
    _, _, _, mk = _make(cfg)
    a, b = mk(), mk()
    table = torch.as_tensor(util.action_table(), device="cuda:0")
    g = torch.Generator(device="cuda:0"); g.manual_seed(3)
    episodes_seen = 0
    for chunk in (1, 7, 33, 2, 60):
        acts = table[torch.randint(0, 6, (chunk, B), device="cuda:0", generator=g)].contiguous()
        for t in range(chunk):
            a.step(acts[t])
        b.rollout(acts)
        torch.cuda.synchronize()
        fa, fb = _fields(a, "var_cam" in cfg), _fields(b, "var_cam" in cfg)
        for k in fa:
            same = (fa[k] == fb[k]) | ((fa[k] != fa[k]) & (fb[k] != fb[k])) if fa[k].dtype.kind == "f" else (fa[k] == fb[k])
            assert same.all(), (name, k, "chunk", chunk, "first mismatching env", np.argwhere(~same)[:3].tolist())
        sa, sb = np.asarray(a.stats()), np.asarray(b.stats())
        assert np.array_equal(sa, sb), (name, chunk, sa, sb)
        episodes_seen = int(sa[1])
    if cfg.get("scatter"):
        assert episodes_seen > B // 8, "auto-reset inside a rollout chunk must be exercised"
    assert b.launch_count() < a.launch_count() or B > 8288
    a.close(); b.close()


def test_rollout_matches_oracle():
    cfg = CASES["cfg2_scattered"]
    B = cfg["B"]
    p, worlds, poses, mk = _make(cfg)
    env = mk()
    ob = util.oracle_batch(p, worlds, poses)
    n = env.num_agents
    rng = np.random.RandomState(17)
    table = util.action_table()
    fields = util.BATCH_FIELDS + util.TRACKER_FIELDS
    episodes = 0
    for chunk in (16, 1, 40, 25):
        acts = table[rng.randint(0, 6, (chunk, B))]
        for t in range(chunk):
            ob.step(acts[t], auto_reset=True)
            episodes += int(ob.gather(trackers=False)["done"].sum())
        env.rollout(torch.as_tensor(acts, device="cuda:0"))
        h = util.gpu_fields(env, fields)
        o = ob.gather(trackers=True)
        d, r = util.batch_mismatch(h, o, n, trackers=True, planner=0)
        for k, m in d.items():
            assert not m.any(), (k, "chunk", chunk, "envs", np.nonzero(m)[0][:8].tolist(), "of", int(m.sum()))
        for k, v in r.items():
            assert float(v.max()) <= 1e-9, (k, "chunk", chunk, "env", int(v.argmax()), float(v.max()))
    st = env.stats()
    assert st[0] == B * 82 and st[1] == episodes and episodes > B // 8
    ob.close(); env.close()


def test_rollout_rejects_other_planners():
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    from gym_drone2d_activeperception_b200 import _native
    p = Params(debug=False, planner="Primitive", gaze_method="NoControl", map_id=1, agent_number=10)
    env = Drone2DVecEnv(p, 8, seeds=1 + np.arange(8), device="cuda:0", auto_reset=True)
    with pytest.raises(_native.Drone2DNativeError):
        env.rollout(torch.zeros((3, 8), dtype=torch.float64, device="cuda:0"))
    env.close()


@pytest.mark.parametrize("planner,gaze", [("NoMove", None), ("Primitive", "Oxford")])
def test_lazy_reset_equals_eager_reset(planner, gaze):
    """d2d_request_reset (re-initialisation inside the next step, the auto-reset path) against d2d_reset + the same step."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    B = 96
    p = Params(debug=False, planner=planner, gaze_method=gaze or "NoControl", map_id=1, agent_number=10, agent_radius=15,
               agent_max_speed=20)
    mk = lambda: Drone2DVecEnv(p, B, seeds=1 + np.arange(B), device="cuda:0", auto_reset=False, oxford=gaze == "Oxford")
    a, b = mk(), mk()
    table = torch.as_tensor(util.action_table(), device="cuda:0")
    g = torch.Generator(device="cuda:0"); g.manual_seed(9)
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda:0"); mask[::3] = 1
    names = util.BATCH_FIELDS + util.TRACKER_FIELDS + ["env_records"] + (["oxford_calls"] if gaze else [])

    def step(env, acts):
        env.step(env.plan_oxford() if gaze else acts)
    for t in range(70):
        acts = table[torch.randint(0, 6, (B,), device="cuda:0", generator=g)]
        if t in (25, 26, 50):
            a.reset(mask)
            b.reset(mask, lazy=True)
        step(a, acts); step(b, acts)
        if t >= 25:
            for k in names:
                x, y = a.buffer(k).cpu().numpy(), b.buffer(k).cpu().numpy()
                same = (x == y) | ((x != x) & (y != y)) if x.dtype.kind == "f" else (x == y)
                assert same.all(), (planner, k, "step", t, np.argwhere(~same)[:3].tolist())
    a.close(); b.close()
