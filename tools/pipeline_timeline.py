"""Timeline of d2d_step_pipelined (GPU box only: python tools/pipeline_timeline.py) with the D2D_WARP_PROF build: per-warp %globaltimer stamps of the LAST kernel of a
pipelined run (it was pre-launched by the previous call; the host then sleeps before the call that opens its gate)."""
import os, sys, time, ctypes, subprocess
sys.path.insert(0, os.getcwd())
import numpy as np
from gym_drone2d_activeperception_b200 import build as b, _native
lib = os.path.join(os.path.dirname(b.LIB), "libdrone2d_prof.so")
subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + ["-DD2D_WARP_PROF", "-o", lib] +
               [os.path.join(b.CSRC, s) for s in b.SOURCES], check=True)
_native.LIB_PATH = lib
import torch, bench
from gym_drone2d_activeperception_b200 import Params
from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
cfg = bench.CONFIGS[2]; B, pk = cfg["envs"], cfg["params"]
worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B))
env = Drone2DVecEnv(Params(debug=False, **pk), B, worlds=worlds, device="cuda:0", auto_reset=True)
lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory(); yaw = torch.empty((B,), dtype=torch.float32).pin_memory(); done = torch.empty((B,), dtype=torch.uint8).pin_memory()
a_bound = torch.zeros(B, dtype=torch.float64).pin_memory()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    env.bind_host_io(a_bound, lm, yaw, done)
    for t in range(5): env.step_bound()
    for delay in (0.0, 200e-6, 1000e-6):
        for t in range(50):
            env.step_pipelined(prelaunch_next=True)
        te = time.perf_counter() + delay
        while time.perf_counter() < te: pass
        t0 = time.perf_counter()
        env.step_pipelined(prelaunch_next=False)
        call_us = (time.perf_counter() - t0) * 1e6
        torch.cuda.synchronize()
        p = env.buffer("warp_prof").cpu().numpy().astype(np.int64)
        k0 = p[:, 0].min()
        print("host delay %6.0f us | last call %7.1f us | kernel: first start 0, last start %.1f, median pre-gate stamp %.1f, "
              "median post-gate stamp %.1f, last end %.1f us (relative to first warp start); median gate wait %.1f us, max %.1f"
              % (delay * 1e6, call_us, (p[:, 0].max() - k0) / 1e3, np.median(p[:, 8] - k0) / 1e3, np.median(p[:, 9] - k0) / 1e3,
                 (p[:, 3].max() - k0) / 1e3, np.median(p[:, 9] - p[:, 8]) / 1e3, (p[:, 9] - p[:, 8]).max() / 1e3), flush=True)
    env.bind_host_io(None, None, None, None)
env.close()
