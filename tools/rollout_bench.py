#!/usr/bin/env python
"""Times d2d_rollout (K steps per launch, env state resident on chip) against K single-step launches on one GPU, bench.py's
way: burn-in to the stationary episode mix, R replicas of the batch visited round robin (cold L2), timed region >= 100 ms,
median over repeats.  GPU box only.

    python tools/rollout_bench.py [--config 2] [--chunks 1,4,16,64] [--envs B] [--burn-in 1000]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--chunks", default="1,4,16,64")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--burn-in", type=int, default=1000)
    ap.add_argument("--region-ms", type=float, default=100.0)
    ap.add_argument("--planner", default="NoMove")
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    import bench
    cfg = bench.make_cfg(args.config, planner=args.planner, envs=args.envs)
    B, pk = cfg["envs"], cfg["params"]
    worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B), unique=min(B, 16384))
    from gym_drone2d_activeperception_b200 import _native
    if args.lib:
        _native.LIB_PATH = args.lib
    import torch
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    dev = torch.device("cuda:0")
    p = Params(debug=False, **pk)
    chunks = [int(x) for x in args.chunks.split(",")]
    KA = 256
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    actions = table[torch.randint(0, 6, (KA, B), device=dev, generator=gen)].contiguous()
    mk = lambda: Drone2DVecEnv(p, B, worlds=worlds, device=dev, auto_reset=True)
    env = mk()
    N = env.num_agents
    touched = B * (bench.D2D_STATE_BYTES + 56 * N)
    R = max(1, min(64, -(-(256 << 20) // touched)))
    envs = [env] + [mk() for _ in range(R - 1)]
    for e in envs:                                   # burn-in with the rollout kernel itself (bit-identical to single steps)
        left = args.burn_in
        while left > 0:
            k = min(left, KA)
            e.rollout(actions[:k])
            left -= k
    torch.cuda.synchronize()

    def timed(fn, steps_per_call):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        M = int(min(4000, max(7, -(-args.region_ms // max(1e-3, e0.elapsed_time(e1))))))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(M + 1)]
        evs[0].record()
        for m in range(M):
            fn()
            evs[m + 1].record()
        torch.cuda.synchronize()
        ms = np.array([evs[m].elapsed_time(evs[m + 1]) for m in range(M)]) / steps_per_call
        return float(np.median(ms)), float(ms.min()), float(ms.max()), M

    # (a) bench.py's device-resident method: K single-step launches captured as one graph, replicas round robin
    K = 40
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for t in range(K):
                envs[t % R].step(actions[t])
    med, lo, hi, M = timed(graph.replay, K)
    print(json.dumps({"mode": "single-step launches (graph of %d)" % K, "config": args.config, "envs": B, "agents": N, "replicas": R,
                      "us_per_step_median": 1e3 * med, "us_min": 1e3 * lo, "us_max": 1e3 * hi,
                      "env_steps_per_s_M": B / med / 1e3}), flush=True)
    del graph
    # (b) rollout: one launch = `chunk` steps of one replica; replicas visited round robin
    for chunk in chunks:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for r in range(R):
                    envs[r].rollout(actions[:chunk])
        med, lo, hi, M = timed(graph.replay, chunk * R)
        print(json.dumps({"mode": "rollout", "steps_per_launch": chunk, "config": args.config, "envs": B, "agents": N, "replicas": R,
                          "us_per_step_median": 1e3 * med, "us_min": 1e3 * lo, "us_max": 1e3 * hi,
                          "env_steps_per_s_M": B / med / 1e3, "repeats": M}), flush=True)
        del graph
    for e in envs:
        e.close()


if __name__ == "__main__":
    main()
