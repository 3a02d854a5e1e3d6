"""`Params`: the configuration object of the reference (utils.py:65-171), mirrored field for field so that code
written against the reference (`Params(...)`, `Params.from_parser()`, attribute mutation by scripts) keeps working."""
import argparse


class Params(object):
    def __init__(self, env='gym-2d-perception-v2', debug=True, record_img=False, trained_policy=False,
                 policy_dir='./trained_policy/lookahead.zip', dt=0.1, map_scale=10, map_size=[500, 500],
                 agent_radius=10, drone_max_acceleration=40, drone_radius=10, drone_max_yaw_speed=80,
                 drone_view_depth=80, drone_view_range=90, img_dir='./', max_flight_time=80, gaze_method='LookAhead',
                 planner='Primitive', var_cam=0, drone_max_speed=40, motion_profile='CVM', pillar_number=0,
                 agent_number=10, agent_max_speed=40, map_id=0, init_pos=[50, 50], target_list=[[50, 460]],
                 static_map='maps/empty_map.npy'):
        self.env = env
        # utils.py:75-80: debug=True -> render on / record off
        self.render = bool(debug)
        self.record = not bool(debug)
        self.record_img = record_img
        self.trained_policy = trained_policy
        self.policy_dir = policy_dir
        self.dt = dt
        self.map_scale = map_scale
        self.map_size = list(map_size)
        self.agent_radius = agent_radius
        self.drone_max_acceleration = drone_max_acceleration
        self.drone_radius = drone_radius
        self.drone_max_yaw_speed = drone_max_yaw_speed
        self.drone_view_depth = drone_view_depth
        self.drone_view_range = drone_view_range
        self.img_dir = img_dir
        self.max_flight_time = max_flight_time
        self.gaze_method = gaze_method
        self.planner = planner
        self.var_cam = var_cam
        self.drone_max_speed = drone_max_speed
        self.motion_profile = motion_profile
        self.pillar_number = pillar_number
        self.agent_number = agent_number
        self.agent_max_speed = agent_max_speed
        self.map_id = map_id
        self.init_position = list(init_pos)
        self.target_list = [list(t) for t in target_list]
        self.static_map = static_map

    @classmethod
    def from_parser(cls, argv=None):
        """Same flags as the reference (utils.py:108-171); note `--debug` is store_false there (passing it turns
        rendering OFF and recording ON), reproduced as is."""
        ap = argparse.ArgumentParser(description='Initialize Params class with command-line arguments')
        ap.add_argument('--env', default='gym-2d-perception-v2')
        ap.add_argument('--debug', action='store_false')
        ap.add_argument('--record_img', action='store_true')
        ap.add_argument('--trained_policy', action='store_true')
        ap.add_argument('--policy_dir', default='./trained_policy/lookahead.zip')
        ap.add_argument('--dt', type=float, default=0.1)
        ap.add_argument('--map_scale', type=int, default=10)
        ap.add_argument('--map_size', nargs=2, type=int, default=[500, 500])
        ap.add_argument('--agent_radius', type=int, default=10)
        ap.add_argument('--drone_max_acceleration', type=int, default=40)
        ap.add_argument('--drone_radius', type=int, default=10)
        ap.add_argument('--drone_max_yaw_speed', type=int, default=80)
        ap.add_argument('--drone_view_depth', type=int, default=80)
        ap.add_argument('--drone_view_range', type=int, default=90)
        ap.add_argument('--img_dir', default='./')
        ap.add_argument('--max_flight_time', type=int, default=80)
        ap.add_argument('--gaze_method', default='LookAhead')
        ap.add_argument('--planner', default='Primitive')
        ap.add_argument('--var_cam', type=int, default=0)
        ap.add_argument('--drone_max_speed', type=int, default=40)
        ap.add_argument('--motion_profile', default='CVM')
        ap.add_argument('--pillar_number', type=int, default=0)
        ap.add_argument('--agent_number', type=int, default=10)
        ap.add_argument('--agent_max_speed', type=int, default=40)
        ap.add_argument('--map_id', type=int, default=0)
        ap.add_argument('--init_pos', nargs=2, type=int, default=[50, 50])
        ap.add_argument('--target_list', nargs='+', type=int, default=[[50, 460]])
        ap.add_argument('--static_map', default='maps/empty_map.npy')
        a = ap.parse_args(argv)
        tl = a.target_list
        if tl and not isinstance(tl[0], (list, tuple)):
            tl = [tl[i:i + 2] for i in range(0, len(tl), 2)]
        return cls(env=a.env, debug=a.debug, record_img=a.record_img, trained_policy=a.trained_policy,
                   policy_dir=a.policy_dir, dt=a.dt, map_scale=a.map_scale, map_size=a.map_size,
                   agent_radius=a.agent_radius, drone_max_acceleration=a.drone_max_acceleration,
                   drone_radius=a.drone_radius, drone_max_yaw_speed=a.drone_max_yaw_speed,
                   drone_view_depth=a.drone_view_depth, drone_view_range=a.drone_view_range, img_dir=a.img_dir,
                   max_flight_time=a.max_flight_time, gaze_method=a.gaze_method, planner=a.planner, var_cam=a.var_cam,
                   drone_max_speed=a.drone_max_speed, motion_profile=a.motion_profile, pillar_number=a.pillar_number,
                   agent_number=a.agent_number, agent_max_speed=a.agent_max_speed, map_id=a.map_id,
                   init_pos=a.init_pos, target_list=tl, static_map=a.static_map)


# constants of the reference (utils.py:11-29)
grid_type = {'DYNAMIC_OCCUPIED': 3, 'OCCUPIED': 1, 'UNOCCUPIED': 2, 'UNEXPLORED': 0}
state_machine = {'WAIT_FOR_GOAL': 0, 'GOAL_REACHED': 1, 'PLANNING': 2, 'EXECUTING': 3}
