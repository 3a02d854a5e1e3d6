"""Gaze policies with the reference's interface (yaw_planner.py): `policy.plan(observation) -> a in [-1, 1]` where
`observation` is `env.info`.  `Oxford` runs on the device (d2d_plan_oxford); the scalar policies are restated on the
host.  Like the reference's `experiment.py:33-34`, the classes also work when the CLASS object is used as the instance
(`policy = Oxford; policy.__init__(policy, params); policy.plan(policy, env.info)`)."""
import math

import numpy as np


class NoControl(object):                     # yaw_planner.py:10-16
    def __init__(self, params):
        self.params = params

    def plan(self, state):
        return 0


class Rotating(object):                      # yaw_planner.py:136-142
    def __init__(self, params):
        self.params = params

    def plan(self, observation):
        return 1


def _clamped_turn(target_yaw, yaw, dt, max_yaw_speed):
    # yaw_planner.py:35-39 / 250-255
    if abs(target_yaw - yaw) < 180:
        yaw_vel = max(min((target_yaw - yaw) / dt, max_yaw_speed), -max_yaw_speed)
    else:
        yaw_vel = -max(min((target_yaw - yaw) / dt, max_yaw_speed), -max_yaw_speed)
    return yaw_vel / max_yaw_speed


class LookAhead(object):                     # yaw_planner.py:18-39
    def __init__(self, params):
        self.dt = params.dt
        self.params = params

    def plan(self, state):
        v = state["drone"].velocity
        if v[1] == 0 and v[0] == 0:
            return 0
        target_yaw = math.degrees(math.atan2(-v[1], v[0])) % 360
        return _clamped_turn(target_yaw, state["drone"].yaw, self.dt, self.params.drone_max_yaw_speed)


class LookGoal(object):                      # yaw_planner.py:225-255
    def __init__(self, params):
        self.params = params

    def plan(self, observation):
        trajectory, drone = observation["trajectory"], observation["drone"]
        if len(trajectory) == 0:
            return 0
        positions = trajectory.positions
        x_look, y_look = positions[-1][0], positions[-1][1]
        grid = drone.map.grid_map                 # one device read; then OccupancyGridMap.get_grid (utils.py:545-548)
        w, h = self.params.map_size
        scale = self.params.map_scale
        for position in positions:
            x, y = position[0], position[1]
            val = 1 if (x >= w or x < 0 or y >= h or y < 0) else grid[int(x // scale), int(y // scale)]
            if val == 0:
                x_look, y_look = x, y
                break
        target_yaw = math.degrees(math.atan2(-(y_look - drone.y), x_look - drone.x)) % 360
        return _clamped_turn(target_yaw, drone.yaw, self.params.dt, self.params.drone_max_yaw_speed)


class Oxford(object):                        # yaw_planner.py:41-127, executed by d2d_oxford_kernel
    def __init__(self, params):
        self.params = params
        self.v_yaw_space = np.arange(-params.drone_max_yaw_speed, params.drone_max_yaw_speed,
                                     params.drone_max_yaw_speed / 3)

    def plan(self, observation):
        env = observation["drone"]._env      # the facade that produced this info dict
        return float(env._vec.plan_oxford()[0].item())

    @property
    def last_time_observed_map(self):
        raise AttributeError("policy state lives on the device: env._vec.buffer('oxford_last_time_observed')")


class Owl(object):                           # yaw_planner.py:151-222, executed by d2d_owl_kernel
    """State (U_list, the repeated-action queue) lives on the device and is re-initialised with the env."""

    def __init__(self, params):
        self.params = params
        self.dt = 0.8
        self.u_space = np.arange(-params.drone_max_yaw_speed, params.drone_max_yaw_speed, params.drone_max_yaw_speed / 10)

    def plan(self, observation):
        env = observation["drone"]._env
        return float(env._vec.plan_gaze("Owl")[0].item())

    @property
    def U_list(self):
        raise AttributeError("policy state lives on the device: env._vec.buffer('owl_U')")


policy_list = {"LookAhead": LookAhead, "NoControl": NoControl, "Oxford": Oxford, "Rotating": Rotating,
               "LookGoal": LookGoal, "Owl": Owl}
