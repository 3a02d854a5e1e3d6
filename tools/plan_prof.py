"""Per-search timeline of d2d_plan_small_kernel: builds libdrone2d_planprof.so with -DD2D_PLAN_PROF (%globaltimer at the start
and the end of every A* search, expansions, nodes, outcome, SM) and prints what sets the length of the kernel -- how long
an expansion takes, how long the longest searches run, how busy the search slots are.  GPU box only:
    python tools/plan_prof.py --config 4 [--steps 8]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--burn-in", type=int, default=200)
    args = ap.parse_args()
    from gym_drone2d_activeperception_b200 import build as b, _native
    lib = os.path.join(os.path.dirname(b.LIB), "libdrone2d_planprof.so")
    if not os.path.isfile(lib) or "--rebuild" in sys.argv:
        b.build_variant("planprof", ["-DD2D_PLAN_PROF"])
    _native.LIB_PATH = lib
    import torch
    import bench
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    cfg = bench.make_cfg(args.config)
    B = args.envs or cfg["envs"]
    pk = cfg["params"]
    ox = cfg["gaze"] == "Oxford"
    worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B), unique=min(B, 8192))
    env = Drone2DVecEnv(Params(debug=False, **pk), B, worlds=worlds, device="cuda:0", auto_reset=True, trackers=True, oxford=ox)
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80, device="cuda:0")
    g = torch.Generator(device="cuda:0")
    g.manual_seed(1)

    def act():
        return env.plan_oxford() if ox else table[torch.randint(0, 6, (B,), device="cuda:0", generator=g)].contiguous()
    for _ in range(args.burn_in):
        env.step(act())
    prof = env.buffer("warp_prof")
    print("# %s, %d envs; per step: searches, kernel span = last end - first start" % (cfg["name"], B))
    per_exp, long_d, spans = [], [], []
    for t in range(args.steps):
        a = act()
        prof.zero_()
        env.step(a)
        torch.cuda.synchronize()
        p = prof.cpu().numpy().astype(np.int64)
        r = p[p[:, 1] > 0]
        t0, t1, nodes, outcome, itr = r[:, 0], r[:, 1], r[:, 2], r[:, 3], r[:, 8]
        dur = (t1 - t0) * 1e-3
        span = (t1.max() - t0.min()) * 1e-3
        spans.append(span)
        full = itr >= 100
        order = np.argsort(t1)
        last = order[-5:]
        print("step %d: %5d searches (%4d full-length, %4d ok), span %6.1f us, start of last search at %6.1f us, "
              "median search %5.1f us, full-length searches %5.1f / %5.1f / %5.1f us (min / med / max), max nodes %d"
              % (t, len(r), full.sum(), (outcome == 1).sum(), span, (t0.max() - t0.min()) * 1e-3, np.median(dur),
                 dur[full].min() if full.any() else 0, np.median(dur[full]) if full.any() else 0, dur[full].max() if full.any() else 0,
                 nodes.max()))
        print("        the 5 searches that end last: " + ", ".join(
            "[start %.0f us, %.0f us long, %d expansions, %d nodes, %d trackers]" % ((t0[i] - t0.min()) * 1e-3, dur[i], itr[i] - 1, nodes[i], r[i, 7])
            for i in last))
        m = itr >= 12
        per_exp.append(dur[m] / (itr[m] - 1))
        long_d.append(dur[full])
        sm_busy = {}
        for s_, d_ in zip(r[:, 4], dur):
            sm_busy[s_] = sm_busy.get(s_, 0.0) + d_
        v = np.array(list(sm_busy.values()))
        print("        search-time per SM (sum of its searches' durations / span): min %.2f  median %.2f  max %.2f slots busy"
              % (v.min() / span, np.median(v) / span, v.max() / span))
    pe = np.concatenate(per_exp)
    print("time per expansion over searches of >= 11 expansions: p10 %.2f  median %.2f  p90 %.2f us" %
          (np.percentile(pe, 10), np.median(pe), np.percentile(pe, 90)))
    print("kernel span: median %.1f us over %d steps" % (np.median(spans), len(spans)))
    env.close()


if __name__ == "__main__":
    main()
