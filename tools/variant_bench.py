#!/usr/bin/env python
"""Times kernel variants of the NoMove step against each other on one GPU (bench.py's method: R replicas round robin for a
cold L2, burn-in to the stationary episode mix, K-step CUDA graph replayed for >= 100 ms, median replay) and checks that
every variant leaves bit-identical state behind.  GPU box only.

    python tools/variant_bench.py [--config 2] [--variants 2,5,6,7,9] [--envs B] [--burn-in 1000]

variants = d2d_config.envs_per_block codes: 0 default, 2 warp kernel (round-1 default), 1 / 3 its ILP2 forms,
5 / 6 / 7 / 9 pooled kernel with 4 / 7 / 14 / 28 envs per block."""
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
    ap.add_argument("--variants", default="2,5,6,7,9")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--burn-in", type=int, default=1000)
    ap.add_argument("--region-ms", type=float, default=150.0)
    ap.add_argument("--planner", default=None)
    ap.add_argument("--lib", default=None, help="alternative libdrone2d build to load")
    ap.add_argument("--trackers", type=int, default=1, help="0: Kalman trackers off (what-if: cost of the tracker phase)")
    ap.add_argument("--agent-number", type=int, default=-1, help="override the config's agent_number (what-if: 0 = no agents)")
    ap.add_argument("--no-auto-reset", action="store_true", help="what-if: no env is ever re-initialised")
    args = ap.parse_args()
    import bench
    cfg = bench.make_cfg(args.config, planner=args.planner, envs=args.envs)
    B, pk = cfg["envs"], cfg["params"]
    if args.agent_number >= 0:
        pk = dict(pk, agent_number=args.agent_number)
    worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B), unique=min(B, 16384))
    from gym_drone2d_activeperception_b200 import _native
    if args.lib:
        _native.LIB_PATH = args.lib
    import torch
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    dev = torch.device("cuda:0")
    p = Params(debug=False, **pk)
    K = args.steps
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    actions = table[torch.randint(0, 6, (K + 8, B), device=dev, generator=gen)].contiguous()
    ref_state = None
    out = []
    for v in [int(x) for x in args.variants.split(",")]:
        mk = lambda: Drone2DVecEnv(p, B, worlds=worlds, device=dev, auto_reset=not args.no_auto_reset, envs_per_block=v,
                                   trackers=bool(args.trackers))
        env = mk()
        N = env.num_agents
        touched = B * (bench.D2D_STATE_BYTES + 56 * N)
        R = max(1, min(64, -(-(256 << 20) // touched)))
        envs = [env] + [mk() for _ in range(R - 1)]
        for t in range(args.burn_in):
            for e in envs:
                e.step(actions[t % (K + 8)])
        torch.cuda.synchronize()
        state = {k: env.buffer(k).clone() for k in ("belief", "local_map", "agent_pos", "drone_yaw", "done", "steps", "hit",
                                                    "tracker_mu", "tracker_sigma", "tracker_active", "tracker_buffer_ts")}
        same = None
        if ref_state is None:
            ref_state = state
        else:
            same = all(torch.equal(state[k], ref_state[k]) for k in state)
        side = torch.cuda.Stream(device=dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for t in range(K):
                    envs[t % R].step(actions[t])
        torch.cuda.synchronize()
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        M = int(min(4000, max(7, -(-args.region_ms // max(1e-3, e0.elapsed_time(e1))))))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(M + 1)]
        evs[0].record()
        for m in range(M):
            graph.replay()
            evs[m + 1].record()
        torch.cuda.synchronize()
        ms = np.array([evs[m].elapsed_time(evs[m + 1]) for m in range(M)]) / K
        r = {"variant": v, "config": args.config, "envs": B, "agents": N, "replicas": R, "us_per_step_median": 1e3 * float(np.median(ms)),
             "us_min": 1e3 * float(ms.min()), "us_max": 1e3 * float(ms.max()), "env_steps_per_s_M": B / float(np.median(ms)) / 1e3,
             "state_equal_to_first_variant": same}
        print(json.dumps(r), flush=True)
        out.append(r)
        del graph
        for e in envs:
            e.close()
    return out


if __name__ == "__main__":
    main()
