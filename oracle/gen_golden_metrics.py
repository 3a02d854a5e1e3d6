"""Golden values of the reference's difficulty metrics (needs /root/reference; build container only).

    python oracle/gen_golden_metrics.py

TEST INFRASTRUCTURE ONLY.  script/difficulty_calculator/survivability_calculator.py and density_calculator.py run a
parameter sweep at import time, so their `env_metrics` bodies are restated here (loop for loop) around the UNMODIFIED
reference env; the outputs pin `metrics.survivability` / `metrics.obstacle_density`."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner as rr  # noqa: E402
from gen_golden import provenance, OUT  # noqa: E402
import json  # noqa: E402


def survive_times(env, params, T, position_step=60):
    # survivability_calculator.py:24-47
    from numpy.linalg import norm
    x_range = range(params.map_scale + params.drone_radius, params.map_size[0] - params.map_scale - params.drone_radius, position_step)
    y_range = range(params.map_scale + params.drone_radius, params.map_size[1] - params.map_scale - params.drone_radius, position_step)
    st = np.ones((len(x_range), len(y_range))) * T
    env.step(0)
    for t in np.arange(0, T, 0.1):
        for x in x_range:
            for y in y_range:
                p = np.array([x, y])
                for agent in env.agents:
                    if norm(agent.position - p) < agent.radius + env.drone.radius:
                        st[x_range.index(x), y_range.index(y)] = min(t, st[x_range.index(x), y_range.index(y)])
        env.step(0)
    st = st - 0.1
    st[st < 0] = 0
    return st


def global_collision_state(env, params, T, position_step):
    # glob_survivability_calculator.py:24-41, loop for loop (the script itself runs a sweep at import time)
    x_range = range(params.map_scale + params.drone_radius, params.map_size[0] - params.map_scale - params.drone_radius, position_step)
    y_range = range(params.map_scale + params.drone_radius, params.map_size[1] - params.map_scale - params.drone_radius, position_step)
    collision_state = np.zeros((len(x_range), len(y_range), int(T / 0.1)))
    for x in x_range:
        for y in y_range:
            env.reset()
            for t in np.arange(0, T, 0.1):
                env.drone.x = x
                env.drone.y = y
                _, _, done, info = env.step(0)
                if info['collision_flag'] == 2:
                    collision_state[int((x - params.map_scale - params.drone_radius) / position_step),
                                    int((y - params.map_scale - params.drone_radius) / position_step), int(t / 0.1)] = 1
    return collision_state


def main_global():
    """T = 24 as in the script: int(t / 0.1) maps 10 of the 240 steps to the previous slot (k = 43, 81, 86, ...)."""
    T, step, seed = 24, 120, 5
    kw = dict(agent_number=30, agent_radius=20, agent_max_speed=40)
    env, params = rr.make_env(planner="NoMove", gaze_method="NoControl", map_id=seed, debug=False, **kw)
    with rr._reference_cwd():
        cs = global_collision_state(env, params, T, step)
    print("global survivability golden: collisions", int(cs.sum()), "shape", cs.shape)
    np.savez_compressed(os.path.join(OUT, "metrics_global_survivability.npz"), collision_state=cs.astype(np.uint8), T=T,
                        position_step=step, seed=seed, params=json.dumps(kw), provenance=json.dumps(provenance()))


def main():
    cases = [dict(agent_number=10, agent_radius=15, agent_max_speed=20), dict(agent_number=20, agent_radius=8, agent_max_speed=45)]
    seeds = [0, 3, 7]
    T = 6
    out = {"T": T, "seeds": np.array(seeds), "provenance": json.dumps(provenance()), "cases": json.dumps(cases)}
    for ci, c in enumerate(cases):
        sts, dens = [], []
        for s in seeds:
            env, params = rr.make_env(planner="NoMove", gaze_method="NoControl", map_id=s, debug=False, **c)   # debug only switches rendering on (utils.py); the script turns it off again
            obs_area = 0                                       # density_calculator.py:27-30 (a plain += loop, not sum())
            for agent in env.agents:
                obs_area += 3.14 * agent.radius ** 2
            dens.append(obs_area / (params.map_size[0] * params.map_size[1]))
            sts.append(survive_times(env, params, T))
        out["survive_%d" % ci] = np.array(sts)
        out["density_%d" % ci] = np.array(dens)
        print(c, [float(x.mean()) for x in sts], dens)
    np.savez_compressed(os.path.join(OUT, "metrics_survivability.npz"), **out)


if __name__ == "__main__":
    if "global" in sys.argv[1:]:
        main_global()
    else:
        main()
