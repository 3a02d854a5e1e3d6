"""Timeline of the resident gated kernel behind d2d_step_pipelined (GPU box only: python tools/resident_timeline.py), with the
D2D_WARP_PROF build: %globaltimer stamps of every warp for the LAST step of a run of pipelined steps, plus the host-side cost
of the calls.  Answers: how long is a warp's action-independent part, how long does it wait at the gate for the host, how
long from the gate to the host seeing the step complete."""
import os, sys, time, ctypes, subprocess
sys.path.insert(0, os.getcwd())
import numpy as np
from gym_drone2d_activeperception_b200 import build as b, _native
lib = os.path.join(os.path.dirname(b.LIB), "libdrone2d_prof.so")
subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + ["-DD2D_WARP_PROF=2" if "--counters" in sys.argv else "-DD2D_WARP_PROF", "-o", lib] +
               [os.path.join(b.CSRC, s) for s in b.SOURCES], check=True)
_native.LIB_PATH = lib
import torch, bench
from gym_drone2d_activeperception_b200 import Params
from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
cfg = bench.CONFIGS[2]; B, pk = cfg["envs"], cfg["params"]
worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B))
env = Drone2DVecEnv(Params(debug=False, **pk), B, worlds=worlds, device="cuda:0", auto_reset=True)
table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80)
acts = table[torch.randint(0, 6, (64, B))].contiguous().pin_memory()
lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory(); yaw = torch.empty((B,), dtype=torch.float32).pin_memory(); done = torch.empty((B,), dtype=torch.uint8).pin_memory()
a_bound = torch.zeros(B, dtype=torch.float64).pin_memory()
srcs = [acts[t].data_ptr() for t in range(64)]
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    env.rollout(acts[:64].cuda()); env.rollout(acts[:64].cuda())
    for r in range(6):
        env.rollout(acts.cuda())                       # towards the stationary episode mix
    env.bind_host_io(a_bound, None if os.environ.get("NO_LM") else lm, yaw, done)
    for t in range(5): env.step_bound()
    for n in (200, 200, 200):
        env.buffer("warp_prof")[:, 5:8] = 0
        t0 = time.perf_counter()
        for t in range(n):
            ctypes.memmove(a_bound.data_ptr(), srcs[t % 64], B * 8)
            env.step_pipelined(prelaunch_next=(t != n - 1))
        dt = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        p = env.buffer("warp_prof").cpu().numpy().astype(np.int64)
        k0 = p[:, 0].min()
        pre = (p[:, 8] - p[:, 0]) / 1e3; wait = (p[:, 9] - p[:, 8]) / 1e3; post = (p[:, 3] - p[:, 9]) / 1e3
        rays = (p[:, 2] - p[:, 1]) / 1e3
        pub = (p[:, 3] - p[:, 4]) / 1e3; g0 = p[:, 9].min()
        print("   gate pass spread: med %.1f p90 %.1f max %.1f us after the first | publication (post-gate work done -> end): med %.1f max %.1f | post-gate work med %.1f max %.1f"
              % (np.median(p[:, 9] - g0) / 1e3, np.percentile(p[:, 9] - g0, 90) / 1e3, (p[:, 9].max() - g0) / 1e3, np.median(pub), pub.max(),
                 np.median(p[:, 4] - p[:, 9]) / 1e3, (p[:, 4] - p[:, 9]).max() / 1e3))
        steps_now = env.buffer("steps").cpu().numpy(); hits = (env.buffer("hit").cpu().numpy() > 0).sum(1)
        act = env.buffer("tracker_active").cpu().numpy().sum(1)
        order = np.argsort(-pre)[:12]
        print("   slowest action-independent parts: " + "; ".join(
            "env %d pre %.1f (head %.1f rays %.1f tail %.1f) steps=%d hits=%d trk=%d" % (
                i, pre[i], (p[i, 1] - p[i, 0]) / 1e3, rays[i], (p[i, 8] - p[i, 2]) / 1e3, steps_now[i], hits[i], act[i]) for i in order))
        # counters accumulate over the n steps of the run: per-step means per env against the LAST step's ray time is only a
        # hint; the clean signal is the split by "env was re-initialised in the last step" (steps == 1)
        fresh = steps_now == 1
        print("   ray phase of the last step: fresh envs (%d) med %.1f us | others med %.1f p99 %.1f max %.1f | marks per env-step (mean over run) %.2f, samples per ray %.2f, longest ray %d samples"
              % (int(fresh.sum()), np.median(rays[fresh]) if fresh.any() else 0.0, np.median(rays[~fresh]), np.percentile(rays[~fresh], 99), rays[~fresh].max(),
                 p[:, 5].mean() / n, p[:, 6].mean() / n / 50.0, int(p[:, 7].max())))
        slow_old = np.argsort(-np.where(fresh, 0, rays))[:6]
        print("   slowest non-fresh ray phases: " + "; ".join("env %d rays %.1f marks/step %.2f steps=%d ix,iy=%s hits=%d" % (
            i, rays[i], p[i, 5] / n, steps_now[i], (int(env.buffer("drone_x")[i].item() // 10), int(env.buffer("drone_y")[i].item() // 10)), hits[i]) for i in slow_old))
        print("%d steps: %.1f us per call | last step, per warp (us): start spread %.1f | pre-gate med %.1f p90 %.1f max %.1f (rays med %.1f) | "
              "gate wait med %.1f min %.1f max %.1f | gate -> end med %.1f max %.1f | first start -> last gate arrival %.1f, -> first gate open %.1f, -> last end %.1f"
              % (n, dt, (p[:, 0].max() - k0) / 1e3, np.median(pre), np.percentile(pre, 90), pre.max(), np.median(rays),
                 np.median(wait), wait.min(), wait.max(), np.median(post), post.max(),
                 (p[:, 8].max() - k0) / 1e3, (p[:, 9].min() - k0) / 1e3, (p[:, 3].max() - k0) / 1e3), flush=True)
    env.bind_host_io(None, None, None, None)
env.close()
