"""The batched oracle harness (oracle.OracleBatch: snapshot / restore / step_batch / gather) against the plain
one-env-at-a-time oracle it wraps: a restored env must be indistinguishable from a freshly created one, and the packed
arrays must equal the per-env views.  CPU only."""
import numpy as np

import util


def _worlds(planner, static_map, agent_number, B, gaze="NoControl", **kw):
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    p = Params(debug=False, planner=planner, gaze_method=gaze, map_id=3, static_map=static_map, agent_number=agent_number, **kw)
    return p, generate_worlds(p, 3 + np.arange(B))


def _same(ob, i, e, n):
    g = ob.gather()
    assert np.array_equal(g["belief"][i], e.belief) and np.array_equal(g["hit"][i, :n], e.hit[:n])
    assert np.array_equal(g["local_map"][i], e.local_map)
    assert (g["collision_flag"][i], g["done"][i], g["steps"][i], g["state_machine"][i]) == \
           (e.c.collision, e.c.done, e.c.steps, e.c.state_machine)
    assert (g["drone_x"][i], g["drone_y"][i], g["drone_yaw"][i]) == (e.c.x, e.c.y, e.c.yaw)
    assert np.array_equal(g["agent_pos"][i, :n], e.apos[:n]) and np.array_equal(g["agent_pref"][i, :n], e.apref[:n])
    assert np.array_equal(g["tracker_active"][i, :n], e.trk_active[:n])
    assert np.array_equal(g["tracker_mu"][i, :n], e.trk_mu[:n]) and np.array_equal(g["tracker_sigma"][i, :n], e.trk_sigma[:n])
    assert (g["tracker_buffer_count"][i], g["tracker_buffer_ts"][i]) == (e.c.buf_count, e.c.buf_ts)


def test_batch_auto_reset_equals_fresh_env_nomove():
    B, steps = 6, 260
    p, worlds = _worlds("NoMove", "maps/empty_map.npy", 14, B, agent_radius=18, agent_max_speed=40)
    rng = np.random.RandomState(1)
    poses = worlds["drone_pose"].copy()
    poses[:, 0] = rng.uniform(60, 440, B); poses[:, 1] = rng.uniform(60, 440, B)     # mid-map: collisions end episodes
    ob = util.oracle_batch(p, worlds, poses, threads=3)
    singles = [util.oracle_env_from_world(p, worlds, i, drone=poses[i]) for i in range(B)]
    table = util.action_table()
    resets = 0
    for t in range(steps):
        acts = table[rng.randint(0, 6, B)]
        for i in range(B):
            if singles[i].c.done:               # what auto-reset means: a brand-new env from the same world
                singles[i].close()
                singles[i] = util.oracle_env_from_world(p, worlds, i, drone=poses[i])
                resets += 1
        applied, _ = ob.step(acts, auto_reset=True)
        assert np.array_equal(applied, acts)
        for i in range(B):
            singles[i].step(float(acts[i]))
        if t % 13 == 0 or t == steps - 1:
            for i in range(B):
                _same(ob, i, singles[i], ob.n)
    assert resets >= 3, "the scenario must exercise the restore path"
    ob.close()


def test_batch_auto_reset_equals_fresh_env_primitive_oxford():
    B, steps = 3, 420
    p, worlds = _worlds("Primitive", "maps/empty_map.npy", 10, B, gaze="Oxford", agent_radius=15, agent_max_speed=20)
    ob = util.oracle_batch(p, worlds, threads=2)
    singles = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    resets = 0
    for t in range(steps):
        want = np.zeros(B)
        for i in range(B):
            if singles[i].c.done:
                singles[i].close()
                singles[i] = util.oracle_env_from_world(p, worlds, i)
                resets += 1
            want[i] = singles[i].oxford_plan()
            singles[i].step(want[i])
        applied, _ = ob.step(policy="Oxford", auto_reset=True)
        assert np.array_equal(applied, want), t
        if t % 29 == 0 or t == steps - 1:
            for i in range(B):
                _same(ob, i, singles[i], ob.n)
                assert np.array_equal(ob.envs[i].ox_last, singles[i].ox_last)
                assert ob.envs[i].c.traj_len == singles[i].c.traj_len
    assert resets >= 2
    ob.close()


def test_search_from_a_start_that_is_not_free_fails():
    """The Primitive step kernel decides a search without running it when the start position itself is not free (sample 0 of
    every primitive is the start at global time 0, traj_planner.py:179-183 -> no successor -> open set empty -> plan() False,
    :149-153) unless the start lies within the goal threshold (:160).  Held here against the oracle's full A*: Planner.is_free
    (traj_planner.py:28-59) restated in Python on the state the search saw -- the drone's pre-step position (step_pos comes
    after plan), the belief grid and the trackers as the step left them."""
    import math
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B = 24
    p = Params(debug=False, planner="Primitive", map_id=40, static_map="maps/obstacle_map.npy", agent_number=16,
               agent_radius=15, agent_max_speed=20, drone_max_speed=40)
    worlds = generate_worlds(p, 40 + np.arange(B))
    envs = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    table = np.arange(-80, 80, 80 / 3) / 80
    rng = np.random.RandomState(3)

    def grid(e, x, y):                                   # OccupancyGridMap.get_grid utils.py:545-548
        if x >= p.map_size[0] or x < 0 or y >= p.map_size[1] or y < 0:
            return 1
        return int(e.belief[int(x // p.map_scale), int(y // p.map_scale)])

    def start_is_free(e, x, y):
        sd = p.drone_radius + 10
        for px, py in ((x - sd, y), (x, y), (x + sd, y), (x, y - sd), (x, y + sd)):
            if grid(e, px, py) == 1:
                return False
        for k in range(e.n):
            if e.trk_active[k]:
                ex, ey = e.trk_mu[k, 0] + 0.0 * e.trk_mu[k, 2], e.trk_mu[k, 1] + 0.0 * e.trk_mu[k, 3]
                if math.sqrt((x - ex) ** 2 + (y - ey) ** 2) <= p.drone_radius + e.trk_radius[k] + 5 + p.var_cam:
                    return False
        return True

    searches = blocked = blocked_ok = free_failed = 0
    for t in range(400):
        for i, e in enumerate(envs):
            if e.c.done:
                e.close()
                envs[i] = e = util.oracle_env_from_world(p, worlds, i)
            x0, y0 = e.c.x, e.c.y
            e.step(float(table[rng.randint(0, 6)]))
            if not e.c.planned:
                continue
            searches += 1
            at_goal = math.hypot(x0 - e.c.target[0], y0 - e.c.target[1]) <= 10
            if not at_goal and not start_is_free(e, float(np.around(x0)), float(np.around(y0))):
                blocked += 1
                blocked_ok += int(e.c.plan_ok)
            elif not e.c.plan_ok:
                free_failed += 1
    for e in envs:
        e.close()
    assert searches > 200 and blocked > 20 and free_failed > 0      # all three kinds of search occurred
    assert blocked_ok == 0                                           # ... and none from a blocked start succeeded
