"""The B=1 facade + Experiment runner reproduce BASELINE config 1 (`main.py --gaze_method Oxford --planner Primitive
--agent_number 10 --agent_max_speed 20 --agent_radius 15 --drone_max_speed 40 --map_id 1`) end to end on the device."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_experiment_config1_matches_reference_episode():
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.experiment import Experiment
    g = util.load_golden(util.golden_files("episode_cfg1")[0])
    p = Params.from_parser(["--debug", "--gaze_method", "Oxford", "--planner", "Primitive", "--agent_number", "10",
                            "--agent_max_speed", "20", "--agent_radius", "15", "--drone_max_speed", "40", "--map_id", "1"])
    ex = Experiment(p, None)
    row = ex.run()
    T = len(g["done"])
    assert abs(row["Flight time"] - T * 0.1) < 1e-9 and row["Success"] == 1            # 210 steps, goal reached
    assert row["Static Collision"] == 0 and row["Dynamic Collision"] == 0 and row["state machine"] == 1
    assert row["Grid discovered"] == int((g["belief"][-1] != 0).sum())
    assert row["Agent tracked"] == int(g["buf_count"][-1])
    assert abs(row["Agent tracked time"] - g["buf_ts"][-1] * 0.1 / g["buf_count"][-1]) < 1e-9
    env = ex.env
    assert (env.drone.x, env.drone.y) == (42.0, 455.0) and abs(env.drone.yaw - g["drone"][-1][2]) < 1e-9
    assert np.array_equal(env.drone.map.grid_map, g["belief"][-1])
    assert np.array_equal(env.map_gt.grid_map == 1, g["gt_grid"] == 1)
    assert np.allclose(env.agents[0].position, g["agent_pos"][-1][0], rtol=1e-12)
    env.close()


def test_facade_info_and_pose_write():
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.env import make
    from gym_drone2d_activeperception_b200.policies import LookAhead, LookGoal, NoControl, Rotating
    p = Params(debug=False, planner="Primitive", gaze_method="LookAhead", map_id=3, agent_number=6)
    env = make("gym-2d-perception-v2", params=p)
    assert env.reset() == {}
    info = env.info
    assert len(info["trajectory"]) == 0 and info["state_machine"] == 0 and info["drone"].yaw == 270.0
    pol = LookAhead(p)
    for _ in range(5):
        a = pol.plan(env.info)
        state, rew, done, info = env.step(a)
    assert rew == 0 and state["local_map"].shape == (1, 33, 33) and state["local_map"].dtype == np.uint8
    assert state["yaw_angle"].dtype == np.float32 and len(info["trajectory"].positions) == len(info["trajectory"])
    assert info["trajectory"].positions[0].shape == (2,) and -1 <= LookGoal(p).plan(info) <= 1
    assert NoControl(p).plan(info) == 0 and Rotating(p).plan(info) == 1
    env.drone.x = 123.5                    # scripts write the pose directly (glob_survivability_calculator.py:36-37)
    assert env.drone.x == 123.5
    env.close()


def test_global_survivability_matches_oracle():
    """metrics.global_survivability vs the oracle run the way glob_survivability_calculator.py:30-41 runs the reference."""
    import oracle
    from gym_drone2d_activeperception_b200 import Params, generate_worlds
    from gym_drone2d_activeperception_b200.metrics import global_survivability, survivability_positions, mean_survival_time
    p = Params(debug=False, planner="NoMove", gaze_method="NoControl", agent_number=20, agent_radius=15, agent_max_speed=40)
    seeds = [4, 9]
    T = 6
    got = global_survivability(p, seeds, T=T)
    xs, ys = survivability_positions(p)
    assert got.shape == (2, 8, 8, 60) and xs[0] == 20 and xs[-1] == 440
    worlds = generate_worlds(p, seeds)
    op = util.oracle_params(p)
    for wi in range(2):
        for a, x in enumerate(xs[::3]):
            for b, y in enumerate(ys[::3]):
                e = oracle.OracleEnv(op, worlds["agent_pos"][wi], worlds["agent_pref"][wi], worlds["agent_radius"][wi],
                                     worlds["gt_grid"][wi], worlds["tracker_radius"][wi], drone=(x, y, 270.0),
                                     targets=p.target_list)
                ref = []
                for t in range(60):
                    e.c.x, e.c.y = float(x), float(y)
                    e.step(0.0)
                    ref.append(1 if e.c.collision == 2 else 0)
                assert np.array_equal(got[wi, a * 3, b * 3], np.array(ref, dtype=np.uint8)), (wi, x, y)
                e.close()
    assert got.sum() > 0 and mean_survival_time(got).shape == (2, 8, 8)
