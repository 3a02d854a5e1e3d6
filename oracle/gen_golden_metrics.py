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
    main()
