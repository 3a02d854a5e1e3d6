"""Where the end-to-end step time goes at BASELINE config 2 (GPU box only): python tools/e2e_breakdown.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    cfg = bench.CONFIGS[2]
    B, pk = cfg["envs"], cfg["params"]
    worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B))
    env = Drone2DVecEnv(Params(debug=False, **pk), B, worlds=worlds, device="cuda:0", auto_reset=True, trackers=True)
    K = 300
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80)
    a_host = table[torch.randint(0, 6, (K, B))].contiguous().pin_memory()
    a_dev = a_host.cuda()
    lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory()
    yaw = torch.empty((B,), dtype=torch.float32).pin_memory()
    done = torch.empty((B,), dtype=torch.uint8).pin_memory()
    env.bind_host_mirror(lm, yaw, done)
    for t in range(10):
        env.step_host(a_host[t], lm, yaw, done)

    def timeit(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(K):
            fn(t)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print("%-64s %7.2f us/step  %7.1f M env-steps/s" % (name, dt * 1e6, B / dt / 1e6))

    timeit("env.step_host (python wrapper, mirror bound)", lambda t: env.step_host(a_host[t], lm, yaw, done))
    a_bound = torch.zeros(B, dtype=torch.float64).pin_memory()
    env.bind_host_io(a_bound, lm, yaw, done)
    env.step_bound()

    def bound_step(t):
        a_bound.copy_(a_host[t])
        env.step_bound()
    timeit("env.step_bound (bound buffers, stream sync) + action memcpy", bound_step)
    timeit("env.step_bound only (same actions every step)", lambda t: env.step_bound())
    import ctypes
    dst, nb = a_bound.data_ptr(), B * 8
    src = [a_host[t].data_ptr() for t in range(K)]

    def pipe_step(t):
        ctypes.memmove(dst, src[t], nb)
        env.step_pipelined(prelaunch_next=t != K - 1)
    timeit("env.step_pipelined (next step pre-launched, gated on its actions) + action memcpy", pipe_step)

    def pipe_step_slow_policy(t):
        ctypes.memmove(dst, src[t], nb)
        env.step_pipelined(prelaunch_next=t != K - 1)
        t_end = time.perf_counter() + 20e-6           # a host policy that needs 20 us per step
        while time.perf_counter() < t_end:
            pass
    timeit("env.step_pipelined + a 20 us host policy per step", pipe_step_slow_policy)

    def bound_step_slow_policy(t):
        ctypes.memmove(dst, src[t], nb)
        env.step_bound()
        t_end = time.perf_counter() + 20e-6
        while time.perf_counter() < t_end:
            pass
    timeit("env.step_bound + a 20 us host policy per step", bound_step_slow_policy)
    env.bind_host_io(None, None, None, None)
    env.bind_host_mirror(lm, yaw, done)
    env.step_host(a_host[0], lm, yaw, done)
    L, h = env._lib, env._h
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ap = [C.c_void_p(a_host[t].data_ptr()) for t in range(K)]
    lp, yp, dp = C.c_void_p(lm.data_ptr()), C.c_void_p(yaw.data_ptr()), C.c_void_p(done.data_ptr())
    timeit("d2d_step_host via ctypes, pointers precomputed", lambda t: L.d2d_step_host(h, ap[t], lp, yp, dp, st))
    adp = [C.c_void_p(a_dev[t].data_ptr()) for t in range(K)]
    timeit("d2d_step (device actions) + torch.cuda.synchronize per step",
           lambda t: (L.d2d_step(h, adp[t], st), torch.cuda.synchronize()))
    timeit("d2d_step (device actions), no per-step sync", lambda t: L.d2d_step(h, adp[t], st))
    timeit("torch.cuda.synchronize only", lambda t: torch.cuda.synchronize())
    timeit("ctypes call d2d_launch_count only", lambda t: L.d2d_launch_count(h))
    env.bind_host_mirror(None, None, None)
    timeit("d2d_step (device actions), no per-step sync, mirror UNBOUND", lambda t: L.d2d_step(h, adp[t], st))
    timeit("d2d_step + sync per step, mirror UNBOUND", lambda t: (L.d2d_step(h, adp[t], st), torch.cuda.synchronize()))
    env.close()


if __name__ == "__main__":
    main()
