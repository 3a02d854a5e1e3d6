"""`Experiment`: the reference's episode driver (experiment.py:26-106) on the device-backed facade, plus a batched
runner that plays many seeded episodes at once and reports the same per-episode statistics as totals."""
import os

import numpy as np

from . import _native
from .env import make
from .params import state_machine
from .policies import policy_list

CSV_COLUMNS = ["Method", "Planner", "Motion Profile", "Map ID", "Agent size", "Number of agents", "Number of pillars",
               "Agent speed", "Drone speed", "Depth variance", "Initial position", "Target position", "Flight time",
               "Grid discovered", "Agent tracked", "Agent tracked time", "Success", "Static Collision",
               "Dynamic Collision", "Freezing", "Dead Lock", "state machine"]


class Experiment(object):
    def __init__(self, params, dir=None):
        if params.gaze_method == "NoControl":
            params.drone_view_range = 360                        # experiment.py:28-29
        self.params = params
        self.env = make(params.env, params=params)
        self.dt = params.dt
        self.policy = policy_list[params.gaze_method](params)
        self.result_dir = dir
        self.rows = []

    def run(self, max_steps=None):
        """One episode (experiment.py:65-106); returns the CSV row as a dict (and appends it to `dir` if recording)."""
        self.env.reset()
        done, n = False, 0
        info = self.env.info
        while not done:
            a = self.policy.plan(info)
            _, _, done, info = self.env.step(a)
            n += 1
            if max_steps is not None and n >= max_steps:
                break
        p = self.params
        buf = info["tracker_buffer"]
        tracking_time = float(np.array([len(t.ts) * 0.1 for t in buf]).sum())
        grid = info["drone"].map.grid_map
        row = dict(zip(CSV_COLUMNS, [
            p.gaze_method, p.planner, p.motion_profile, p.map_id, p.agent_radius, p.agent_number, p.pillar_number,
            p.agent_max_speed, p.drone_max_speed, p.var_cam, p.init_position, p.target_list[0], info["flight_time"],
            int(grid.shape[0] * grid.shape[1] - np.sum(grid == 0)), len(buf),
            tracking_time / len(buf) if len(buf) else float("nan"),
            1 if info["state_machine"] == state_machine["GOAL_REACHED"] else 0, 1 if info["collision_flag"] == 1 else 0,
            1 if info["collision_flag"] == 2 else 0, info["freezing_flag"], info["dead_lock_flag"], info["state_machine"]]))
        self.rows.append(row)
        if self.result_dir and getattr(p, "record", False):
            import pandas as pd
            df = pd.DataFrame([row])
            df.to_csv(self.result_dir, mode="a", index=False, header=not os.path.isfile(self.result_dir))
        return row


def run_batched(params, num_envs, steps, device="cuda:0", seeds=None):
    """Plays `steps` env-steps of `num_envs` seeded envs (auto-reset on) with the params' gaze method (Oxford, LookAhead,
    LookGoal, Rotating, NoControl) evaluated on the device.  Returns the statistics dict."""
    import torch
    from .vec_env import Drone2DVecEnv
    if params.gaze_method == "NoControl":
        params.drone_view_range = 360                            # experiment.py:28-29
    env = Drone2DVecEnv(params, num_envs, seeds=seeds, device=device, auto_reset=True,
                        oxford=params.gaze_method == "Oxford")
    out = torch.empty(num_envs, dtype=torch.float64, device=env.device)
    for _ in range(steps):
        env.step(env.plan_gaze(params.gaze_method, out))
    st = env.stats()
    env.close()
    return dict(zip(_native.STAT_NAMES, [int(v) for v in st]))
