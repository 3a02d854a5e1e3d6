"""Live cross-check of the oracle against the UNMODIFIED reference (only where /root/reference is mounted)."""
import numpy as np
import pytest

import util

ref_runner = pytest.importorskip("ref_runner")
pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_runner.reference_available(), reason="reference checkout not present")]


def _compare(steps, policy=None, set_pose=None, **kw):
    from gym_drone2d_activeperception_b200.params import Params
    acts = util.action_table().tolist()
    r = ref_runner.run_episode(steps, actions=acts, policy=policy, set_pose=set_pose, stop_on_done=policy is not None,
                               record_oxford=policy == "Oxford", **kw)
    P = r["params"]
    p = Params(debug=False, **{k: P[k] for k in util.PARAM_KEYS if k in P}, init_pos=P["init_position"],
               target_list=P["target_list"])
    world = dict(agent_pos=r["agent_pos0"], agent_pref=r["agent_pref0"], agent_radius=r["agent_radius"],
                 tracker_radius=r["tracker_radius"], gt_grid=r["gt_grid"], drone_pose=r["drone0"],
                 agent_vel=r["agent_vel0"], obstacles=r["obstacles"])
    e = util.oracle_env_from_world(p, world)
    n = r["n_agents"]
    for t in range(len(r["done"])):
        if policy == "Oxford":
            assert e.oxford_plan() == r["action"][t]
            assert np.array_equal(e.ox_last, r["ox_last"][t])
        d = e.step(float(r["action"][t]))
        assert np.array_equal(e.belief, r["belief"][t]) and np.array_equal(e.hit[:n], r["hit"][t]), t
        assert (e.c.collision, d) == (r["collision"][t], bool(r["done"][t])), t
        assert (e.c.x, e.c.y, e.c.yaw, e.c.vx, e.c.vy) == (*r["drone"][t], *r["drone_vel"][t]), t
        assert np.array_equal(e.apos[:n], r["agent_pos"][t]), t
        if p.motion_profile == "RVO":
            assert np.array_equal(e.avel[:n], r["agent_vel"][t]) and np.array_equal(e.apref[:n], r["agent_pref"][t]), t
        assert np.array_equal(e.local_map, r["local_map"][t]), t
        assert e.c.traj_len == r["traj_len"][t] and e.c.state_machine == r["state_machine"][t], t
        assert np.array_equal(e.trk_active[:n].astype(bool), r["trk_active"][t]), t
    e.close()


@pytest.mark.parametrize("seed", [3, 11])
@pytest.mark.parametrize("smap", ["maps/empty_map.npy", "maps/obstacle_map.npy"])
def test_nomove_live(seed, smap):
    _compare(40, planner="NoMove", map_id=seed, static_map=smap, set_pose=(200.5 + seed, 260.25, 77.0))


@pytest.mark.parametrize("kw", [dict(map_id=13, agent_number=7, pillar_number=2), dict(map_id=2, agent_number=40, agent_radius=15, agent_max_speed=20)],
                         ids=["pillars", "crowd"])
def test_rvo_live(kw):
    """RVO motion profile (utils.py:299-460): velocities, preferred velocities and positions bit for bit."""
    _compare(25, planner="NoMove", motion_profile="RVO", **kw)


def test_full_episode_live():
    _compare(400, policy="Oxford", planner="Primitive", map_id=6, agent_number=12)


@pytest.mark.parametrize("kw", [
    dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20),
    dict(static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15, agent_max_speed=40),
    dict(static_map="maps/shaped_obstacle_map.npy", agent_number=50, agent_radius=10, agent_max_speed=40),
    dict(static_map="maps/empty_map.npy", agent_number=12, agent_radius=-1, agent_max_speed=30, pillar_number=3),
], ids=["empty", "obstacle", "random0", "shaped50", "pillars_randradius"])
def test_world_generation_live(kw):
    """Host world generator (product code) vs the reference's __init__ on many seeds."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_world
    for seed in range(200, 212):
        env, params = ref_runner.make_env(planner="NoMove", map_id=seed, **kw)
        ws = ref_runner.world_snapshot(env)
        w = generate_world(Params(debug=False, planner="NoMove", map_id=seed, **kw), seed)
        assert np.array_equal(w["agent_pos"], ws["agent_pos0"]) and np.array_equal(w["agent_pref"], ws["agent_pref0"])
        assert np.array_equal(w["agent_radius"], ws["agent_radius"]) and np.array_equal(w["tracker_radius"], ws["tracker_radius"])
        assert np.array_equal(w["gt_grid"] == 1, ws["gt_grid"] == 1), seed
        # inputs of the RVO motion profile: Agent.velocity after __init__ and the pillars (drone_v2.py:14-27, 19, 62)
        assert np.array_equal(w["agent_vel"], ws["agent_vel0"]) and np.array_equal(w["obstacles"], ws["obstacles"]), seed


@pytest.mark.parametrize("policy,kind", [("LookAhead", 2), ("LookGoal", 3), ("Rotating", 1), ("NoControl", 0)])
def test_scalar_gaze_policies_live(policy, kind):
    """Oracle restatement of the scalar gaze policies (yaw_planner.py:10-39, 136-142, 225-255) vs the reference."""
    from gym_drone2d_activeperception_b200.params import Params
    kw = dict(planner="Primitive", map_id=7, agent_number=8, agent_radius=15, agent_max_speed=20, gaze_method=policy)
    r = ref_runner.run_episode(150, policy=policy, stop_on_done=True, **kw)
    P = r["params"]
    p = Params(debug=False, **{k: P[k] for k in util.PARAM_KEYS if k in P}, init_pos=P["init_position"],
               target_list=P["target_list"])
    world = dict(agent_pos=r["agent_pos0"], agent_pref=r["agent_pref0"], agent_radius=r["agent_radius"],
                 tracker_radius=r["tracker_radius"], gt_grid=r["gt_grid"], drone_pose=r["drone0"],
                 agent_vel=r["agent_vel0"], obstacles=r["obstacles"])
    e = util.oracle_env_from_world(p, world)
    for t in range(len(r["done"])):
        a = e.policy_plan(kind)
        assert a == r["action"][t], (policy, t, a, r["action"][t])
        e.step(a)
        assert np.array_equal(e.belief, r["belief"][t]) and (e.c.x, e.c.y, e.c.yaw) == tuple(r["drone"][t]), t
    e.close()


def test_owl_policy_live():
    """Owl (yaw_planner.py:151-222) with the reference's real call pattern (experiment.py:33-34: the class object is the
    instance, so the @classmethod helpers resolve): every action of a whole episode, bit for bit."""
    from gym_drone2d_activeperception_b200.params import Params
    kw = dict(planner="Primitive", map_id=21, agent_number=14, agent_radius=12, agent_max_speed=30, gaze_method="Owl")
    r = ref_runner.run_episode(400, policy="Owl", stop_on_done=True, **kw)
    P = r["params"]
    p = Params(debug=False, **{k: P[k] for k in util.PARAM_KEYS if k in P}, init_pos=P["init_position"],
               target_list=P["target_list"])
    world = dict(agent_pos=r["agent_pos0"], agent_pref=r["agent_pref0"], agent_radius=r["agent_radius"],
                 tracker_radius=r["tracker_radius"], gt_grid=r["gt_grid"], drone_pose=r["drone0"],
                 agent_vel=r["agent_vel0"], obstacles=r["obstacles"])
    e = util.oracle_env_from_world(p, world)
    for t in range(len(r["done"])):
        a = e.owl_plan()
        assert a == r["action"][t], (t, a, r["action"][t])
        e.step(a)
        assert np.array_equal(e.belief, r["belief"][t]) and (e.c.x, e.c.y, e.c.yaw) == tuple(r["drone"][t]), t
    e.close()


@pytest.mark.parametrize("kw,policy", [(dict(map_id=13, agent_number=16, agent_radius=12, agent_max_speed=30), "LookAhead"),
                                       (dict(map_id=3, agent_number=8), "LookAhead"),
                                       (dict(map_id=4, agent_number=25, agent_radius=15, agent_max_speed=20), "Oxford")],
                         ids=["generic", "canonical_tie", "oxford"])
def test_jerk_primitive_live(kw, policy):
    """Jerk_Primitive (traj_planner.py:403-516) restated in the oracle vs the live reference; the tie orders of the unstable
    argsort are taken from this machine's numpy, which is also the one the reference runs on here."""
    from gym_drone2d_activeperception_b200.params import Params
    r = ref_runner.run_episode(400, policy=policy, stop_on_done=True, planner="Jerk_Primitive", gaze_method=policy, **kw)
    P = r["params"]
    p = Params(debug=False, **{k: P[k] for k in util.PARAM_KEYS if k in P}, init_pos=P["init_position"],
               target_list=P["target_list"])
    world = dict(agent_pos=r["agent_pos0"], agent_pref=r["agent_pref0"], agent_radius=r["agent_radius"],
                 tracker_radius=r["tracker_radius"], gt_grid=r["gt_grid"], drone_pose=r["drone0"])
    e = util.oracle_env_from_world(p, world)
    for t in range(len(r["done"])):
        if policy == "Oxford":
            assert e.oxford_plan() == r["action"][t]
        e.step(float(r["action"][t]))
        assert (e.c.x, e.c.y, e.c.yaw, e.c.vx, e.c.vy) == (*r["drone"][t], *r["drone_vel"][t]), t
        assert np.array_equal(e.belief, r["belief"][t]) and e.c.done == int(r["done"][t]), t
        assert bool(e.c.plan_ok) == bool(r["plan_ok"][t]) and e.c.traj_len == r["traj_len"][t], t
    e.close()


@pytest.mark.parametrize("blocker", ["belief_cell", "probe_cell", "tracker", "none"])
def test_reference_search_from_blocked_start_expands_once(blocker):
    """The rule d2d_step_prim_warp_kernel uses to decide a search without running it, on the reference's own Primitive.plan
    (traj_planner.py:125-218): when is_free(start, 0) is False every primitive breaks at its first sample (the start position
    at global time 0), nothing is added to the open set and plan() returns False after ONE expansion -- is_free is called once
    per speed-feasible primitive and never again.  With a free start the same call explores further."""
    ref_utils, ref_env, ref_traj, ref_yaw = ref_runner.import_reference()
    env, params = ref_runner.make_env(planner="Primitive", map_id=7, agent_number=4)
    drone, planner = env.drone, env.planner
    planner.set_target(np.array([50, 460]))
    x, y = drone.x, drone.y
    assert np.linalg.norm(np.array([x, y]) - planner.target[:2]) > planner.search_threshold
    scale = params.map_scale
    if blocker == "belief_cell":
        drone.map.grid_map[int(x // scale), int(y // scale)] = 1
    elif blocker == "probe_cell":                         # one of the four probes at drone_radius + 10 (traj_planner.py:35-47)
        drone.map.grid_map[int((x + params.drone_radius + 10) // scale), int(y // scale)] = 1
    elif blocker == "tracker":
        trk = drone.trackers[0]
        trk.active = True
        trk.mu_upds.append(np.array([[x + 3.0], [y - 2.0], [1.0], [1.0]]))       # estimate_pos reads mu_upds[-1] (utils.py:220-223)
    calls = []
    orig = planner.is_free

    def spy(position, t, occupancy_map, trackers):
        r = orig(position, t, occupancy_map, trackers)
        calls.append((tuple(np.asarray(position, dtype=float)), float(t), bool(r)))
        return r
    planner.is_free = spy
    planner.trajectory.clear()
    ok = planner.plan(drone, params.dt)
    n_prim = len(planner.u_space) ** 2
    if blocker == "none":
        assert len(calls) > n_prim                        # a free start: later samples and later expansions are tested
        return
    assert ok is False and len(planner.trajectory) == 0
    assert 0 < len(calls) <= n_prim                       # one expansion, one (failed) sample per speed-feasible primitive
    assert all(c == ((float(np.around(x)), float(np.around(y))), 0.0, False) for c in calls)
