"""Parity of the Primitive planner path (replan check, A*, brake / step_pos) and the Oxford gaze kernel."""
import numpy as np
import pytest

import util
from test_gpu_parity import _env, _host, _cmp_env_to_oracle, RTOL

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _traj(env, h, i):
    from gym_drone2d_activeperception_b200.vec_env import trajectory_waypoints
    coeff = env.buffer("traj_coeff")[i].cpu().numpy()
    return trajectory_waypoints(env.cfg, coeff, h["traj_nseg"][i], h["traj_cursor"][i])


@pytest.mark.parametrize("path", util.golden_files("episode_"), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_reference_golden_episode(path):
    """Whole episodes of the reference (Primitive + Kalman trackers + Oxford): actions chosen on the device must equal
    the reference's, and every recorded field must match step by step."""
    g = util.load_golden(path)
    p = util.params_from_golden(g)
    n = int(g["n_agents"])
    B = 3
    env = _env(p, B, util.world_from_golden(g, B), auto_reset=False, oxford=True)
    T = len(g["done"])
    plan_i = 0
    for t in range(T):
        a = env.plan_oxford()
        torch.cuda.synchronize()
        ah = a.cpu().numpy()
        assert ah[0] == g["action"][t] and ah[B - 1] == g["action"][t], ("oxford action", t, ah, g["action"][t])
        ox = env.buffer("oxford_last_time_observed").cpu().numpy()
        assert np.array_equal(ox[0], g["ox_last"][t]), ("oxford last_time_observed_map", t)
        env.step(a)
        h = _host(env)
        for i in (0, B - 1):
            assert np.array_equal(h["belief"][i], g["belief"][t]), ("belief", t)
            assert np.array_equal(h["hit"][i, :n], g["hit"][t]), ("hit", t)
            assert h["collision_flag"][i] == g["collision"][t] and bool(h["done"][i]) == bool(g["done"][t]), ("done", t)
            assert h["dead_lock_flag"][i] == g["dead_lock"][t] and h["freezing_flag"][i] == g["freezing"][t]
            assert h["state_machine"][i] == g["state_machine"][t] and h["fail_count"][i] == g["fail_count"][t], ("sm", t)
            assert np.array_equal(h["local_map"][i, 0], g["local_map"][t]), ("local_map", t)
            assert h["yaw_angle"][i, 0] == g["yaw_obs"][t]
            assert util.rel_err([h["drone_x"][i], h["drone_y"][i], h["drone_yaw"][i]], g["drone"][t]) <= RTOL, ("drone", t)
            assert util.rel_err([h["drone_vx"][i], h["drone_vy"][i]], g["drone_vel"][t]) <= RTOL, ("vel", t)
            assert h["traj_nseg"][i] * 20 - h["traj_cursor"][i] == g["traj_len"][t], ("traj_len", t)
            assert bool(h["replan"][i]) == bool(g["replan"][t]) and bool(h["plan_ok"][i]) == bool(g["plan_ok"][t]), ("plan", t)
            act = g["trk_active"][t]
            assert np.array_equal(h["tracker_active"][i, :n].astype(bool), act), ("trk_active", t)
            if act.any():
                assert util.rel_err(h["tracker_mu"][i, :n][act], g["trk_mu"][t][act]) <= RTOL
            assert (h["tracker_buffer_count"][i], h["tracker_buffer_ts"][i]) == (g["buf_count"][t], g["buf_ts"][t])
        if g["planned"][t] and g["plan_ok"][t]:
            pos, vel = _traj(env, h, 0)
            assert np.array_equal(pos, g["plan%d_pos" % plan_i][1:]), ("plan positions", t)
            assert util.rel_err(vel, g["plan%d_vel" % plan_i][1:]) <= RTOL, ("plan velocities", t)
            plan_i += 1
    assert plan_i == len(g["plan_steps"])
    st = env.stats()
    assert st[1] == B                                    # every copy finished exactly one episode
    env.close()


@pytest.mark.parametrize("cfg", [
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, drone_max_speed=40, B=23, steps=260),
    dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20, drone_max_speed=40, B=14, steps=200),
    dict(static_map="maps/empty_map.npy", agent_number=30, agent_radius=10, agent_max_speed=40, drone_max_speed=20, B=9, steps=200),
], ids=["cfg1_like", "cfg4_obstacle", "speed20_crowded"])
@pytest.mark.parametrize("epb", [4, 0], ids=["block4", "warp"])
def test_cuda_matches_oracle_primitive_oxford_with_auto_reset(cfg, epb):
    """Seeded batch, Primitive planner + Oxford policy on the device vs the oracle, with auto-reset: when an episode
    ends the oracle env and its policy state are rebuilt from the same world (what the reference's reset() does)."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = cfg["B"], cfg["steps"]
    p = Params(debug=False, planner="Primitive", gaze_method="Oxford", map_id=500, static_map=cfg["static_map"],
               agent_number=cfg["agent_number"], agent_radius=cfg["agent_radius"], agent_max_speed=cfg["agent_max_speed"],
               drone_max_speed=cfg["drone_max_speed"])
    worlds = generate_worlds(p, 500 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True, oxford=True, envs_per_block=epb)   # 0: d2d_step_prim_warp_kernel (default)
    n = env.num_agents
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    episodes = 0
    for t in range(steps):
        a = env.plan_oxford()
        acts = np.zeros(B)
        for i in range(B):
            if oracles[i].c.done:
                oracles[i].close()
                oracles[i] = util.oracle_env_from_world(p, worlds, i)
                episodes += 1
            acts[i] = oracles[i].oxford_plan()
        torch.cuda.synchronize()
        ah = a.cpu().numpy()
        assert np.array_equal(ah, acts), ("oxford actions", t, np.nonzero(ah != acts))
        env.step(a)
        for i in range(B):
            oracles[i].step(acts[i])
        h = _host(env)
        ox = env.buffer("oxford_last_time_observed").cpu().numpy()
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, "primitive")
            assert h["traj_nseg"][i] * 20 - h["traj_cursor"][i] == e.c.traj_len, ("traj_len", t, i)
            assert bool(h["replan"][i]) == bool(e.c.replan) and bool(h["plan_ok"][i]) == bool(e.c.plan_ok), ("plan", t, i)
            assert np.array_equal(ox[i], e.ox_last), ("ox_last", t, i)
            if e.c.planned and e.c.plan_ok:
                pos, vel = _traj(env, h, i)
                opos, ovel = e.trajectory()
                assert np.array_equal(pos, opos) and util.rel_err(vel, ovel) <= RTOL, ("trajectory", t, i)
    assert episodes > 0
    st = env.stats()
    assert st[11] > 0 and st[1] >= episodes
    from gym_drone2d_activeperception_b200 import _native
    if "psmall40" in _native.LIB_PATH and epb == 0 and cfg["drone_max_speed"] == 40:     # 8 x 8 primitives: the small kernel runs
        # forced-overflow A/B build (build.build_variant, tools/ab_plan.sh): the searches above 40 nodes were abandoned by
        # d2d_plan_small_kernel and redone by d2d_plan_kernel -- the comparisons above held on that path
        assert st[15] > 0
    for e in oracles:
        e.close()
    env.close()


def test_primitive_multiple_targets_and_no_agents():
    """target_list with two goals (state machine GOAL_REACHED -> WAIT_FOR_GOAL -> PLANNING, drone_v2.py:156-163) and an
    empty agent list; scripted gaze; oracle comparison until the episode ends."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B = 5
    p = Params(debug=False, planner="Primitive", map_id=70, agent_number=0, target_list=[[200, 120], [60, 300]])
    worlds = generate_worlds(p, 70 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=False, oxford=False)
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    table = util.action_table()
    rng = np.random.RandomState(5)
    done_at = None
    for t in range(400):
        acts = table[rng.randint(0, 6, B)]
        env.step(torch.as_tensor(acts, device="cuda:0"))
        for i, e in enumerate(oracles):
            e.step(float(acts[i]))
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, 0, t, "targets", trackers=False)
            assert h["traj_nseg"][i] * 20 - h["traj_cursor"][i] == e.c.traj_len
        if oracles[0].c.done:
            done_at = t
            break
    assert done_at is not None and oracles[0].c.state_machine == 1 and oracles[0].c.target_cursor == 2
    env.close()


def test_primitive_crowded_random_map():
    """Primitive + Oxford on random_map_0 (142 agents per env: 20 discs + 122 moving map cells), vs the oracle."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = 6, 80
    p = Params(debug=False, planner="Primitive", gaze_method="Oxford", map_id=20, static_map="maps/random_map_0.npy",
               agent_number=20, agent_radius=15, agent_max_speed=40)
    worlds = generate_worlds(p, 20 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True, oxford=True)
    n = env.num_agents
    assert n == 142
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    for t in range(steps):
        a = env.plan_oxford()
        acts = np.zeros(B)
        for i in range(B):
            if oracles[i].c.done:
                oracles[i].close()
                oracles[i] = util.oracle_env_from_world(p, worlds, i)
            acts[i] = oracles[i].oxford_plan()
        torch.cuda.synchronize()
        assert np.array_equal(a.cpu().numpy(), acts), ("oxford actions", t)
        env.step(a)
        for i in range(B):
            oracles[i].step(acts[i])
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, "crowded")
            assert h["traj_nseg"][i] * 20 - h["traj_cursor"][i] == e.c.traj_len
    env.close()


@pytest.mark.parametrize("planner", ["NoMove", "Primitive"])
def test_cuda_graph_capture_replays_steps(planner):
    """d2d_step / d2d_plan_oxford only enqueue kernels on the caller's stream, so a step loop can be captured in a
    CUDA graph (after one eager warm-up step) and must reproduce the eager results."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B = 64
    p = Params(debug=False, planner=planner, gaze_method="Oxford", map_id=11, agent_number=10, agent_radius=15,
               agent_max_speed=20)
    worlds = generate_worlds(p, 11 + np.arange(B))
    ox = planner == "Primitive"
    eager = _env(p, B, worlds, auto_reset=True, oxford=ox)
    graphed = _env(p, B, worlds, auto_reset=True, oxford=ox)
    acts = torch.full((B,), 1.0 / 3, dtype=torch.float64, device="cuda:0")
    static_a = torch.empty(B, dtype=torch.float64, device="cuda:0")

    def one(env):
        if ox:
            env.plan_oxford(static_a)
            env.step(static_a)
        else:
            env.step(acts)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        one(eager)
        one(graphed)                       # warm-up (sets kernel attributes) outside the capture
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5):
            one(graphed)
    for _ in range(4):
        g.replay()
    for _ in range(20):
        one(eager)
    torch.cuda.synchronize()
    for name in ("belief", "local_map", "drone_x", "drone_yaw", "agent_pos", "done", "steps", "tracker_mu"):
        assert torch.equal(eager.buffer(name), graphed.buffer(name)), name
    eager.close()
    graphed.close()


@pytest.mark.parametrize("policy,kind", [("LookAhead", 2), ("LookGoal", 3), ("Rotating", 1), ("NoControl", 0)])
def test_scalar_gaze_policies_on_device(policy, kind):
    """d2d_plan_gaze vs the oracle's restatement of yaw_planner.py.  atan2 on the device is CUDA's (not glibc's), so
    unsaturated actions are held to 1e-12 and the oracle is driven with the DEVICE's actions (state compared as usual)."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = 12, 150
    p = Params(debug=False, planner="Primitive", gaze_method=policy, map_id=800, agent_number=8, agent_radius=15,
               agent_max_speed=20)
    worlds = generate_worlds(p, 800 + np.arange(B))
    env = _env(p, B, worlds, auto_reset=True, oxford=False)
    n = env.num_agents
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    unsat = 0
    for t in range(steps):
        a = env.plan_gaze(policy)
        torch.cuda.synchronize()
        ah = a.cpu().numpy()
        for i in range(B):
            if oracles[i].c.done:
                oracles[i].close()
                oracles[i] = util.oracle_env_from_world(p, worlds, i)
            ref = oracles[i].policy_plan(kind)
            assert abs(ah[i] - ref) <= 1e-12 * max(1.0, abs(ref)), (policy, t, i, ah[i], ref)
            unsat += int(abs(ref) not in (0.0, 1.0))
        env.step(a)
        for i in range(B):
            oracles[i].step(float(ah[i]))
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, policy)
    if policy in ("LookAhead", "LookGoal"):
        assert unsat > 0
    env.close()


def test_measurement_noise_with_auto_reset():
    """var_cam = 1: noisy measurements drawn from the per-env legacy np.random stream on the device (restored at reset)
    vs the oracle; Primitive planner + Oxford, auto-reset."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, steps = 16, 220
    p = Params(debug=False, planner="Primitive", gaze_method="Oxford", map_id=40, agent_number=14, agent_radius=15,
               agent_max_speed=20, var_cam=1)
    worlds = generate_worlds(p, 40 + np.arange(B))
    assert "rng_key" in worlds and worlds["rng_key"].shape == (B, 624)
    env = _env(p, B, worlds, auto_reset=True, oxford=True)
    n = env.num_agents
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    resets = 0
    for t in range(steps):
        a = env.plan_oxford()
        acts = np.zeros(B)
        for i in range(B):
            if oracles[i].c.done:
                oracles[i].close()
                oracles[i] = util.oracle_env_from_world(p, worlds, i)
                resets += 1
            acts[i] = oracles[i].oxford_plan()
        torch.cuda.synchronize()
        assert np.array_equal(a.cpu().numpy(), acts), ("oxford actions", t)
        env.step(a)
        for i in range(B):
            oracles[i].step(acts[i])
        h = _host(env)
        for i, e in enumerate(oracles):
            _cmp_env_to_oracle(h, i, e, n, t, "noise")
    assert resets > 0 and int(env.buffer("tracker_active").sum()) >= 0
    env.close()


@pytest.mark.parametrize("static_map,B,speed,planner", [("maps/obstacle_map.npy", 700, 40, "Primitive"), ("maps/empty_map.npy", 97, 40, "Primitive"),
                                                        ("maps/obstacle_map.npy", 150, 20, "Primitive"),      # 27 x 27 primitives: large A* kernel
                                                        ("maps/obstacle_map.npy", 64, 40, "NoMove")],         # nothing to overlap: the two calls
                         ids=["obstacle", "empty", "speed20_large_kernel", "nomove"])
def test_step_plan_oxford_equals_step_then_plan(static_map, B, speed, planner):
    """d2d_step_plan_oxford (A* searches on the side stream beside the Oxford scoring of the envs that did not plan) against
    the two calls it stands for, on twin envs with auto-reset: actions, every exposed field, trajectory coefficients, the
    Oxford state and the statistics stay identical step by step -- also in place (next actions written over the inputs)."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    from test_gpu_parity import FIELDS
    p = Params(debug=False, planner=planner, gaze_method="Oxford", map_id=900, static_map=static_map,
               agent_number=10, agent_radius=10, agent_max_speed=20, drone_max_speed=speed)
    worlds = generate_worlds(p, 900 + np.arange(B))
    ref = _env(p, B, worlds, auto_reset=True, oxford=True)
    fused = _env(p, B, worlds, auto_reset=True, oxford=True)
    a_f = fused.plan_oxford()
    extra = ("traj_coeff", "oxford_last_time_observed", "need_plan")
    planned = 0
    for t in range(150):
        a_r = ref.plan_oxford()
        assert torch.equal(a_r, a_f), ("actions", t)
        ref.step(a_r)
        a_f = fused.step_plan_oxford(a_f, out=a_f)          # in place
        torch.cuda.synchronize()
        for k in tuple(FIELDS) + (extra[:1] if planner == "Primitive" else ()) + extra[2:]:     # NoMove: no trajectory buffer to compare
            assert torch.equal(ref.buffer(k), fused.buffer(k)), (k, t)
        planned += int((ref.buffer("need_plan") == 1).sum())
    # the policy state after the fused call is one plan() ahead of the reference env: compare after the reference catches up
    a_r = ref.plan_oxford()
    assert torch.equal(a_r, a_f)
    assert torch.equal(ref.buffer(extra[1]), fused.buffer(extra[1]))
    sr, sf = ref.stats(), fused.stats()
    assert np.array_equal(sr[:14], sf[:14])
    assert planner == "NoMove" or (sr[1] > 0 and planned > 0)        # episodes ended and searches ran on the overlapped path
    ref.close()
    fused.close()
