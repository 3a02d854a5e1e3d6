"""Difficulty metrics of the reference on the batched env.

`global_survivability` is script/difficulty_calculator/glob_survivability_calculator.py:13-44: for every world and every
start position of an 8x8 grid (x, y in range(map_scale + drone_radius, size - map_scale - drone_radius, 60)) the drone is
pinned at the position (`env.drone.x = x` before every step), the env is stepped with the NoMove planner and action 0 for
T / dt steps, and `collision_state[ix, iy, int(t / 0.1)] = 1` whenever `info['collision_flag'] == 2`.  The reference runs the
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
    out = torch.zeros((steps, nw * npos), dtype=torch.uint8, device=env.device)
    col = env.buffer("collision_flag")
    # the reference walks `for t in np.arange(0, T, 0.1)` and writes slot int(t / 0.1) (glob_survivability_calculator.py:34-39):
    # t is k * 0.1 rounded, so for some k the slot is k - 1 (T = 24: k = 43, 81, 86, ...) -- that slot then collects two
    # steps and slot k stays 0.  Reproduced, not corrected.
    slots = [int(t / 0.1) for t in np.arange(0, T, 0.1)][:steps]
    for k in range(len(slots)):
        env.step(zero)
        out[slots[k]] |= (col == 2).to(torch.uint8)
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


def survivability(params, seeds, T=12, position_step=60, device="cuda:0"):
    """script/difficulty_calculator/survivability_calculator.py:13-48 for every seed at once.  The reference builds one
    NoMove env per world, steps it once, then for t in np.arange(0, T, 0.1): marks every position of the 8x8 grid that lies
    inside an agent disc grown by the drone radius (`norm(agent.position - p) < agent.radius + drone.radius`) with
    min(t, ...), and steps again (it keeps stepping after `done`; nothing is reset).  Returns (survive_times
    float64 [len(seeds), len(xs), len(ys)], mean per seed) with the script's post-processing (-0.1, clamped at 0).
    One env per world here; the 64 positions are evaluated from the agent positions on the device."""
    if params.planner != "NoMove":
        raise ValueError("the metric is defined with planner='NoMove' (survivability_calculator.py:20)")
    xs, ys = survivability_positions(params, position_step)
    seeds = np.asarray(seeds, dtype=np.int64)
    nw = len(seeds)
    env = Drone2DVecEnv(params, nw, seeds=seeds, device=device, auto_reset=False, trackers=True, oxford=False)
    dev = env.device
    grid = torch.tensor([(x, y) for x in xs for y in ys], dtype=torch.float64, device=dev)        # [P, 2]
    zero = torch.zeros(nw, dtype=torch.float64, device=dev)
    pos, rad = env.buffer("agent_pos"), env.buffer("agent_radius")
    ts = np.arange(0, T, 0.1)
    first = torch.full((nw, grid.shape[0]), len(ts), dtype=torch.int64, device=dev)               # index of the first hit
    env.step(zero)
    for i in range(len(ts)):
        dx = pos[:, :, None, 0] - grid[None, None, :, 0]
        dy = pos[:, :, None, 1] - grid[None, None, :, 1]
        hit = (torch.sqrt(dx * dx + dy * dy) < (rad + float(params.drone_radius))[:, :, None]).any(1)   # [nw, P]
        first = torch.where(hit & (first == len(ts)), torch.full_like(first, i), first)
        env.step(zero)
    first = first.cpu().numpy()
    env.close()
    tv = np.concatenate([ts, [float(T)]])               # never hit: stays at T (np.ones(...) * T)
    st = tv[first].reshape(nw, len(xs), len(ys)) - 0.1
    st[st < 0] = 0
    return st, st.reshape(nw, -1).mean(1)


def obstacle_density(params, seeds):
    """script/difficulty_calculator/density_calculator.py:13-30: sum of 3.14 * r^2 over the agents / map area (host only)."""
    w = generate_worlds(params, np.asarray(seeds, dtype=np.int64))
    out = []
    for radii in w["agent_radius"]:
        area = 0
        for r in radii.tolist():        # the script's own accumulation order and scalar `**` (libm pow)
            area += 3.14 * r ** 2
        out.append(area / (params.map_size[0] * params.map_size[1]))
    return np.array(out)
