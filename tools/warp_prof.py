"""Per-env (= per-warp) timeline of the fused NoMove step kernel: builds libdrone2d_prof.so with -DD2D_WARP_PROF (four
%globaltimer stamps per env: kernel entry, before the ray phase, after it, kernel exit) and prints where a step's
wall time goes -- launch ramp, per-warp duration spread, stragglers.  GPU box only:  python tools/warp_prof.py [--config 2]"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--burn-in", type=int, default=1000, help="untimed steps before profiling (stationary episode mix)")
    args = ap.parse_args()
    from gym_drone2d_activeperception_b200 import build as b, _native
    lib = os.path.join(os.path.dirname(b.LIB), "libdrone2d_prof.so")
    if not os.path.isfile(lib) or "--rebuild" in sys.argv:
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + ["-DD2D_WARP_PROF", "-o", lib] + \
              [os.path.join(b.CSRC, s) for s in b.SOURCES]
        subprocess.run(cmd, check=True)
    _native.LIB_PATH = lib
    import torch
    import bench
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    cfg = bench.CONFIGS[args.config]
    B = args.envs or cfg["envs"]
    pk = cfg["params"]
    worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B), unique=min(B, 16384))
    env = Drone2DVecEnv(Params(debug=False, **pk), B, worlds=worlds, device="cuda:0", auto_reset=True, trackers=True)
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80, device="cuda:0")
    g = torch.Generator(device="cuda:0")
    g.manual_seed(1)
    acts = table[torch.randint(0, 6, (args.steps, B), device="cuda:0", generator=g)].contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    for t in range(args.burn_in):
        env.step(acts[t % args.steps])
    prof = env.buffer("warp_prof")
    rows = []
    fine_rows = []
    by_nc = {}
    for t in range(args.steps):
        flush.fill_(t & 255)
        done_before = env.buffer("done").clone()
        env.step(acts[t])
        torch.cuda.synchronize()
        p = prof.cpu().numpy().astype(np.int64)
        t0 = p[:, 0].min()
        start, pre, rays, end = p[:, 0] - t0, p[:, 1] - p[:, 0], p[:, 2] - p[:, 1], p[:, 3] - p[:, 2]
        dur = p[:, 3] - p[:, 0]
        total = p[:, 3].max() - t0
        rs = done_before.cpu().numpy() != 0
        if t == args.steps - 1:      # one step's raw per-env data for offline analysis
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", "warp_prof_step_cfg%d.npz" % args.config), prof=p,
                                yaw=env.buffer("drone_yaw").cpu().numpy(), x=env.buffer("drone_x").cpu().numpy(),
                                y=env.buffer("drone_y").cpu().numpy(), reset=rs, steps=env.buffer("steps").cpu().numpy(),
                                hit=env.buffer("hit").cpu().numpy(), belief_cells=(env.buffer("belief") != 0).sum((1, 2)).cpu().numpy(),
                                tracker_active=env.buffer("tracker_active").cpu().numpy(), act=acts[t].cpu().numpy())
        if t >= 10:
            seq = [0, 4, 5, 6, 1, 2, 7, 8, 9, 3]        # stamp slots in program order
            fine_rows.append([np.median(p[:, seq[k + 1]] - p[:, seq[k]]) for k in range(len(seq) - 1)])
        if t >= 10:
            # culled-list length per env (agents within depth + 9*sqrt(2) + r of the drone), from the post-step state
            ap = env.buffer("agent_pos").cpu().numpy().reshape(B, -1, 2)
            ar = env.buffer("agent_radius").cpu().numpy().reshape(B, -1)
            dx = ap[:, :, 0] - env.buffer("drone_x").cpu().numpy()[:, None]
            dy = ap[:, :, 1] - env.buffer("drone_y").cpu().numpy()[:, None]
            nc = (np.hypot(dx, dy) <= ar + 80 + 9 * np.sqrt(2.0)).sum(1)
            for v in np.unique(nc):
                by_nc.setdefault(int(v), []).append((float(rays[nc == v].mean()), int((nc == v).sum())))
            rows.append((total, np.percentile(start, 50), start.max(), np.percentile(dur, 50), np.percentile(dur, 90),
                         dur.max(), np.median(pre), np.median(rays), np.median(end), int(rs.sum()),
                         dur[rs].mean() if rs.any() else 0.0, int(np.argmax(p[:, 3])), bool(rs[np.argmax(p[:, 3])]),
                         pre.max(), rays.max(), end.max()))
    r = np.array(rows, dtype=np.float64)
    names = ["kernel_ns(first entry -> last exit)", "start_p50", "start_max", "dur_p50", "dur_p90", "dur_max", "pre_p50",
             "rays_p50", "tail_p50", "n_reset_envs", "dur_mean_of_reset_envs", "last_env", "last_env_was_reset",
             "pre_max", "rays_max", "tail_max"]
    fine = np.array(fine_rows)
    labels = ["entry -> scalars in shared (loads, bulk issue, hit-mask clear)", "agents phase (Agent.step, culling)",
              "leader_begin + patch decision", "wait for the bulk copies", "ray phase", "trackers", "static probes",
              "leader finish / flags / scalar stores", "observation rewrite + done stats"]
    print("median per-warp phase durations, ns (mean over steps):")
    for lab, v in zip(labels, fine.mean(0)):
        print("  %-62s %8.1f" % (lab, v))
    for i, n in enumerate(names):
        print("%-38s mean %10.1f   min %10.1f   max %10.1f" % (n, r[:, i].mean(), r[:, i].min(), r[:, i].max()))
    print("ray-phase ns by culled-list length (mean over steps; envs per step):")
    for v in sorted(by_nc):
        a = np.array(by_nc[v])
        print("  ncull %3d : rays %9.1f ns   envs/step %8.1f" % (v, a[:, 0].mean(), a[:, 1].mean()))
    env.close()


if __name__ == "__main__":
    main()
