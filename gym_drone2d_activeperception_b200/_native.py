"""ctypes binding of libdrone2d.so (include/drone2d.h).  There is NO fallback: if the CUDA library is missing or
a call fails, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("D2D_LIB") or os.path.join(HERE, "libdrone2d.so")     # D2D_LIB: an A/B build of the same ABI (tools/)

MAX_TARGETS, MAX_U, MAX_SAMP, MAX_WAY, MAX_YAW = 8, 64, 32, 64, 16
NUM_STATS = 16
GAZE = {"NoControl": 0, "Rotating": 1, "LookAhead": 2, "LookGoal": 3, "Owl": 4}
POLICY_OXFORD, POLICY_OWL, MAX_OWL_U = 1, 2, 32
BELIEF_STRIDE = 2560
STAT_NAMES = ["env_steps", "episodes", "success", "static_collision", "dynamic_collision", "freezing", "dead_lock",
              "flight_steps", "grid_discovered", "agents_tracked", "tracked_steps", "plans", "plan_failures", "replans",
              "mirror_bytes", "plan_overflows"]


class D2DConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("num_envs", C.c_int32), ("num_agents", C.c_int32),
        ("planner", C.c_int32), ("trackers", C.c_int32), ("auto_reset", C.c_int32), ("oxford", C.c_int32),
        ("envs_per_block", C.c_int32), ("n_rays", C.c_int32), ("n_targets", C.c_int32),
        ("n_u", C.c_int32), ("n_samp", C.c_int32), ("n_way", C.c_int32), ("n_yaw", C.c_int32), ("motion_profile", C.c_int32),
        ("dt", C.c_double), ("map_scale", C.c_double), ("map_w", C.c_double), ("map_h", C.c_double),
        ("agent_radius", C.c_double),
        ("drone_max_acceleration", C.c_double), ("drone_radius", C.c_double), ("drone_max_yaw_speed", C.c_double),
        ("drone_view_depth", C.c_double), ("drone_view_range", C.c_double), ("max_flight_time", C.c_double),
        ("var_cam", C.c_double), ("drone_max_speed", C.c_double), ("ox_cos_thresh", C.c_double),
        ("targets", (C.c_double * 2) * MAX_TARGETS),
        ("u_space", C.c_double * MAX_U),
        ("t_samp", C.c_double * MAX_SAMP), ("t_samp2", C.c_double * MAX_SAMP),
        ("t_way", C.c_double * MAX_WAY), ("t_way2", C.c_double * MAX_WAY), ("t_way_x2", C.c_double * MAX_WAY),
        ("v_yaw_space", C.c_double * MAX_YAW),
        ("n_owl_u", C.c_int32), ("owl_repeat", C.c_int32), ("owl_u_space", C.c_double * MAX_OWL_U),
    ]


JERK_H, JERK_MAXT = 72, 40


class D2DJerkTables(C.Structure):
    _fields_ = [
        ("dx", C.c_double * JERK_H), ("dy", C.c_double * JERK_H),
        ("T", C.c_double * JERK_H), ("Tp", (C.c_double * 4) * JERK_H),
        ("times", C.c_int32 * JERK_H),
        ("tt", (C.c_double * JERK_MAXT) * JERK_H), ("ttp", ((C.c_double * 4) * JERK_MAXT) * JERK_H),
        ("tie_order", (C.c_uint8 * JERK_H) * 144),
    ]


class D2DBufferInfo(C.Structure):
    _fields_ = [("dev_ptr", C.c_void_p), ("nbytes", C.c_int64), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4), ("strides", C.c_int64 * 4)]


EXPORTS = ["d2d_version", "d2d_last_error", "d2d_create", "d2d_destroy", "d2d_set_world", "d2d_set_rng", "d2d_set_rvo", "d2d_set_jerk_tables", "d2d_reset", "d2d_request_reset", "d2d_step", "d2d_rollout",
           "d2d_step_host", "d2d_bind_host_mirror", "d2d_bind_host_io", "d2d_step_bound", "d2d_step_pipelined", "d2d_plan_oxford", "d2d_step_plan_oxford", "d2d_step_bound_plan_oxford", "d2d_plan_gaze", "d2d_set_drone_pose", "d2d_get_buffer", "d2d_stats",
           "d2d_launch_count"]

_lib = None


class Drone2DNativeError(RuntimeError):
    pass


def load():
    """Loads libdrone2d.so; raises if it has not been built (python -m gym_drone2d_activeperception_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise Drone2DNativeError(
            "libdrone2d.so not found at %s -- build it with `python -m gym_drone2d_activeperception_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.d2d_version.restype = C.c_int
    L.d2d_last_error.restype = C.c_char_p
    L.d2d_last_error.argtypes = [vp]
    L.d2d_create.argtypes = [C.POINTER(D2DConfig), C.POINTER(vp)]
    L.d2d_destroy.argtypes = [vp]
    L.d2d_set_world.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.d2d_set_rng.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.d2d_set_rvo.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32]
    L.d2d_set_jerk_tables.argtypes = [vp, C.POINTER(D2DJerkTables)]
    L.d2d_reset.argtypes = [vp, vp, vp]
    L.d2d_request_reset.argtypes = [vp, vp, vp]
    L.d2d_step.argtypes = [vp, vp, vp]
    L.d2d_rollout.argtypes = [vp, vp, C.c_int32, C.c_int64, vp]
    L.d2d_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.d2d_bind_host_mirror.argtypes = [vp, vp, vp, vp]
    L.d2d_bind_host_io.argtypes = [vp, vp, vp, vp, vp, vp]
    L.d2d_step_bound.argtypes = [vp]
    L.d2d_step_pipelined.argtypes = [vp, C.c_int32]
    L.d2d_plan_oxford.argtypes = [vp, vp, vp]
    L.d2d_step_plan_oxford.argtypes = [vp, vp, vp, vp]
    L.d2d_step_bound_plan_oxford.argtypes = [vp]
    L.d2d_plan_gaze.argtypes = [vp, C.c_int32, vp, vp]
    L.d2d_set_drone_pose.argtypes = [vp, vp, vp]
    L.d2d_get_buffer.argtypes = [vp, C.c_char_p, C.POINTER(D2DBufferInfo)]
    L.d2d_stats.argtypes = [vp, vp, C.c_int32, vp]
    L.d2d_launch_count.argtypes = [vp]
    L.d2d_launch_count.restype = C.c_int64
    for name in EXPORTS:
        if name not in ("d2d_version", "d2d_last_error", "d2d_launch_count"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(handle, rc, what):
    if rc != 0:
        msg = load().d2d_last_error(handle)
        raise Drone2DNativeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
