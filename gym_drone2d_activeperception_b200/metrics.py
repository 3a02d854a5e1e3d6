"""Difficulty metrics of the reference on the batched env.

`global_survivability` is script/difficulty_calculator/glob_survivability_calculator.py:13-44: for every world and every
start position of an 8x8 grid (x, y in range(map_scale + drone_radius, size - map_scale - drone_radius, 60)) the drone is
pinned at the position (`env.drone.x = x` before every step), the env is stepped with the NoMove planner and action 0 for
T / dt steps, and `collision_state[ix, iy, t] = 1` whenever `info['collision_flag'] == 2`.  The reference runs the
64 x 240 steps of one world sequentially on one env object; here all (world, position) pairs are one batch."""
import numpy as np
import torch

from .params import Params
from .vec_env import Drone2DVecEnv
from .world import generate_worlds


def survivability_positions(params, position_step=60):
    lo = params.map_scale + params.drone_radius
    xs = list(range(lo, params.map_size[0] - params.map_scale - params.drone_radius, position_step))
    ys = list(range(lo, params.map_size[1] - params.map_scale - params.drone_radius, position_step))
    return xs, ys


def global_survivability(params, seeds, T=24, position_step=60, device="cuda:0"):
    """Returns collision_state uint8 [len(seeds), len(xs), len(ys), int(T / dt)] (numpy)."""
    if params.planner != "NoMove":
        raise ValueError("the metric is defined with planner='NoMove' (glob_survivability_calculator.py:20)")
    xs, ys = survivability_positions(params, position_step)
    seeds = np.asarray(seeds, dtype=np.int64)
    nw, npos = len(seeds), len(xs) * len(ys)
    steps = int(T / params.dt)
    w = generate_worlds(params, seeds)
    worlds = {k: np.repeat(v, npos, axis=0) for k, v in w.items()}          # env = world * npos + position
    env = Drone2DVecEnv(params, nw * npos, seeds=np.repeat(seeds, npos), worlds=worlds, device=device, auto_reset=False,
                        trackers=True, oxford=False)
    pose = worlds["drone_pose"].copy()
    grid = np.array([(x, y) for x in xs for y in ys], dtype=np.float64)
    pose[:, 0] = np.tile(grid[:, 0], nw)
    pose[:, 1] = np.tile(grid[:, 1], nw)
    env.set_drone_pose(pose)            # NoMove never moves the drone, so pinning once == pinning before every step
    zero = torch.zeros(nw * npos, dtype=torch.float64, device=env.device)
    out = torch.empty((steps, nw * npos), dtype=torch.uint8, device=env.device)
    col = env.buffer("collision_flag")
    for t in range(steps):
        env.step(zero)
        out[t] = (col == 2)
    res = out.cpu().numpy().T.reshape(nw, len(xs), len(ys), steps)
    env.close()
    return res


def mean_survival_time(collision_state, dt=0.1):
    """Time of the first dynamic collision per (world, position), T if none -- the quantity survivability_calculator.py
    averages (utils of the paper's 'survivability' metric)."""
    steps = collision_state.shape[-1]
    hit = collision_state.astype(bool)
    first = np.where(hit.any(-1), hit.argmax(-1), steps)
    return first * dt
