"""Host-side world generation: what `Drone2DEnv2.__init__` produces from a seed (reference envs/drone_v2.py:69-117,
init_obstacles_random_size :12-66, OccupancyGridMap.init_obstacles utils.py:508-525).

`map_id` is only an RNG seed in the reference (drone_v2.py:79-80): the global `random` and `np.random` streams
are both seeded with it, consumed in a fixed order (pillars -> agents -> 100 group headings).  Here each env gets
private `random.Random(seed)` / `np.random.RandomState(seed)` instances, which produce the identical streams.
"""
import os
import random

import numpy as np
from numpy.linalg import norm

_MAP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps")


def load_static_map(static_map):
    """Resolves the reference's 'maps/<name>.npy' paths against the package's own copy of the four shipped maps
    (data files), or any readable path."""
    if isinstance(static_map, np.ndarray):
        return static_map
    if os.path.isfile(static_map):
        return np.load(static_map)
    cand = os.path.join(_MAP_DIR, os.path.basename(static_map))
    if os.path.isfile(cand):
        return np.load(cand)
    raise FileNotFoundError(static_map)


def count_agents(params, static_map=None):
    m = load_static_map(params.static_map if static_map is None else static_map)
    return int(params.agent_number) + int(np.count_nonzero(m))


def generate_world(params, seed, static_map=None):
    """Returns dict(agent_pos [N,2], agent_pref [N,2], agent_radius [N], tracker_radius [N], gt_grid [gw,gh] u8
    (1 = OCCUPIED, 2 otherwise), drone_pose [3], obstacles [P,3])."""
    rnd = random.Random(seed)
    nrs = np.random.RandomState(seed)
    W, H = params.map_size
    scale = params.map_scale
    dx, dy = params.init_position
    drone_xy = np.array([dx, dy])

    # pillars (drone_v2.py:14-26)
    obstacles = []
    while len(obstacles) < params.pillar_number:
        obs = np.array([rnd.randint(50, W - 50), rnd.randint(50, H - 50), rnd.randint(15, 20)])
        free = True
        for target in params.target_list:
            if norm(np.asarray(target) - obs[:-1]) <= params.drone_radius + 20 + obs[-1]:
                free = False
                break
        if norm(drone_xy - obs[:-1]) <= params.drone_radius + 70:
            free = False
        if free:
            obstacles.append(obs)

    # random disc agents (drone_v2.py:28-47)
    pos, pref, rad, trk_rad = [], [], [], []
    n_rand = params.agent_number
    while len(pos) < n_rand:
        p = np.array((rnd.uniform(20, W - 20), rnd.uniform(20, H - 20)))
        r = rnd.uniform(5, 15) if params.agent_radius == -1 else rnd.uniform(params.agent_radius - 2,
                                                                                  params.agent_radius + 2)
        k = len(pos)
        pv = -params.agent_max_speed * np.array([np.cos(2 * np.pi * k / n_rand), np.sin(2 * np.pi * k / n_rand)])
        free = True
        for q, rq in zip(pos, rad):
            if norm(q - p) <= rq + r:
                free = False
        for ob in obstacles:
            if norm(np.array([ob[0], ob[1]]) - p) <= ob[2] + r + 10:
                free = False
        if norm(p - drone_xy) <= params.drone_radius + 70:
            free = False
        if free:
            pos.append(p)
            pref.append(pv)
            rad.append(r)
            trk_rad.append(r)           # drone_v2.py:46

    # every non-zero cell of the "static" map is a moving radius-5 agent (drone_v2.py:49-66)
    smap = load_static_map(params.static_map if static_map is None else static_map)
    vel = params.agent_max_speed
    direction = nrs.rand(100) * 2 * np.pi          # same stream as 100 sequential np.random.rand() calls
    vels = np.stack([vel * np.cos(direction), vel * np.sin(direction)], axis=1)
    xs, ys = np.nonzero(smap)                      # row-major order == the reference's nested x, y loops
    for x, y in zip(xs.tolist(), ys.tolist()):
        pos.append(np.array([5 + x * 10, 5 + y * 10], dtype=np.float64))
        pref.append(vels[int(smap[x][y])].copy())
        rad.append(5.0)
        trk_rad.append(float(params.agent_radius))  # KalmanFilter.__init__ default, utils.py:184

    n = len(pos)
    agent_pos = np.array(pos, dtype=np.float64).reshape(n, 2)
    agent_pref = np.array(pref, dtype=np.float64).reshape(n, 2)
    agent_radius = np.array(rad, dtype=np.float64).reshape(n)
    tracker_radius = np.array(trk_rad, dtype=np.float64).reshape(n)

    # ground truth grid (utils.py:495-525).  Only cells == 1 matter downstream (utils.py:666,770).  NB: a cell
    # whose CENTRE lies inside an agent disc is overwritten with DYNAMIC_OCCUPIED even if it was a border / pillar
    # cell, and update_dynamic_grid (utils.py:527-530) then turns it into UNOCCUPIED for good -- so such cells are
    # not occupancy from the first perception pass on.
    gw, gh = W // scale, H // scale
    gt = np.full((gw, gh), 2, dtype=np.uint8)
    gt[0, :] = 1
    gt[-1, :] = 1
    gt[:, 0] = 1
    gt[:, -1] = 1
    for ob in obstacles:
        for i in range(max(0, int((ob[0] - ob[2]) // scale) - 1), min(gw, int((ob[0] + ob[2]) // scale) + 2)):
            for j in range(max(0, int((ob[1] - ob[2]) // scale) - 1), min(gh, int((ob[1] + ob[2]) // scale) + 2)):
                if norm(np.array([scale * (i + 0.5), scale * (j + 0.5)]) - np.array([ob[0], ob[1]])) <= ob[2]:
                    gt[i, j] = 1
    occ_i, occ_j = np.nonzero(gt == 1)
    if n:
        for i, j in zip(occ_i.tolist(), occ_j.tolist()):
            cx, cy = scale * (i + 0.5), scale * (j + 0.5)
            # cheap reject, then the reference's scalar expression (`**2` is libm pow there, utils.py:523)
            near = np.nonzero((np.abs(agent_pos[:, 0] - cx) <= agent_radius + 1) &
                              (np.abs(agent_pos[:, 1] - cy) <= agent_radius + 1))[0]
            for k in near.tolist():
                ax, ay, r = float(agent_pos[k, 0]), float(agent_pos[k, 1]), float(agent_radius[k])
                if (cx - ax) ** 2 + (cy - ay) ** 2 <= r ** 2:
                    gt[i, j] = 2
                    break
    yaw0 = -90 % 360                                # Drone2D(init_yaw=-90): yaw = init_yaw % 360 (utils.py:718)
    return dict(agent_pos=agent_pos, agent_pref=agent_pref, agent_radius=agent_radius, tracker_radius=tracker_radius,
                gt_grid=gt, drone_pose=np.array([float(dx), float(dy), float(yaw0)]),
                obstacles=np.array(obstacles, dtype=np.float64).reshape(len(obstacles), 3))


def generate_worlds(params, seeds, static_map=None):
    """Stacks generate_world over seeds -> arrays with a leading env axis."""
    smap = load_static_map(params.static_map if static_map is None else static_map)
    ws = [generate_world(params, int(s), smap) for s in seeds]
    keys = ("agent_pos", "agent_pref", "agent_radius", "tracker_radius", "gt_grid", "drone_pose")
    return {k: np.ascontiguousarray(np.stack([w[k] for w in ws])) for k in keys}
