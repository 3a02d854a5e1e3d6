#!/usr/bin/env python
"""bench.py -- env-steps/s (and rays/s) of the batched drone2d hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--envs B]

A "step" is one pass of Drone2DEnv2.step over one batch of synthetic (seeded) environments.  Default workload is
BASELINE.json configs[1]: empty_map.npy, 4096 envs per GPU, 10 agents, perception + dynamics (NoMove planner,
Kalman trackers on as in the reference).  For N > 1 launch with torchrun (one rank per GPU, weak scaling: every
rank steps its own 4096 envs; the only collective is one NCCL all-reduce of the episode statistics).

Prints ONE JSON line (rank 0).  `value`: whole-job env-steps/s with inputs resident in HBM: exactly K steps back to back
(one CUDA graph of the K step launches) between barrier + synchronize, one CUDA-event pair; the L2 is kept cold by
stepping R independent replicas of the batch round robin so that >= 256 MiB of state is touched between two visits of a
replica (`per_step_events` repeats the measurement one launch at a time with an explicit L2 flush).  `e2e`: the same metric through d2d_step_host with pinned HOST buffers
(actions H2D + observation D2H inside the timed region).  `roofline`: algorithmic bytes (SURVEY.md §8d:
2406 + 104*N per env-step) / measured kernel time vs the measured HBM copy peak.  `cpu_baseline`: the oracle port
of the reference path (oracle/drone2d_oracle.c) on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CONFIGS = {
    # BASELINE.json configs[1..4] (SURVEY.md §8d parameters)
    2: dict(name="configs[1]: empty_map.npy, 4096 envs/GPU, 10 agents, perception+dynamics (NoMove)", envs=4096,
            params=dict(planner="NoMove", static_map="maps/empty_map.npy", agent_number=10, agent_radius=15,
                        agent_max_speed=20, map_id=1)),
    3: dict(name="configs[2]: random_map_0.npy, 65536 envs/GPU, 20 agents radius 15, Primitive planner checks on device",
            envs=65536,
            params=dict(planner="Primitive", static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15,
                        agent_max_speed=40, map_id=1)),
    4: dict(name="configs[3]: obstacle_map.npy, 65536 envs/GPU, 10 agents speed 20, Primitive planner + Oxford gaze scoring",
            envs=65536, gaze="Oxford",
            params=dict(planner="Primitive", static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10,
                        agent_max_speed=20, map_id=1)),
    5: dict(name="configs[4]: shaped_obstacle_map.npy, 131072 envs/GPU, 50 agents", envs=131072,
            params=dict(planner="NoMove", static_map="maps/shaped_obstacle_map.npy", agent_number=50, agent_radius=10,
                        agent_max_speed=40, map_id=1)),
}
METRIC = "env-steps/sec"
N_RAYS = 50
D2D_STATE_BYTES = 2560 + 400 + 128            # read per env and step: belief grid, ground-truth rows, env record


def _gen_chunk(args):
    from gym_drone2d_activeperception_b200 import Params, generate_worlds
    pk, seeds = args
    return generate_worlds(Params(debug=False, **pk), seeds)


def make_worlds(pk, seeds, unique=None):
    """Host-side world generation (seeded, the reference's own procedure).  `unique` caps the number of distinct
    worlds (tiled to the batch) to bound set-up time for the very large configs; the default keeps every env unique."""
    import multiprocessing as mp
    seeds = np.asarray(seeds)
    B = len(seeds)
    gen = seeds if unique is None or unique >= B else seeds[:unique]
    nproc = max(1, min(len(os.sched_getaffinity(0)), 32, len(gen) // 64))
    chunks = np.array_split(gen, nproc)
    if nproc > 1:
        with mp.get_context("fork").Pool(nproc) as pool:
            parts = pool.map(_gen_chunk, [(pk, c) for c in chunks])
    else:
        parts = [_gen_chunk((pk, gen))]
    w = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    if len(gen) < B:
        reps = (B + len(gen) - 1) // len(gen)
        w = {k: np.concatenate([v] * reps)[:B] for k, v in w.items()}
    return w


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc, self.mark = gpu_index, [], None, 0

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        rows = self.rows[self.mark:] if len(self.rows) - self.mark >= 3 else self.rows
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "back-to-back steps for 1.5 s right after the timed region (reasons: whole run)"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(n_agents):
    return 2406 + 104 * n_agents          # SURVEY.md §8(d); tracker traffic deliberately NOT counted


def cpu_port_run(pk, n_envs, steps, threads, seed0=1, oxford=False):
    """Times the oracle port of the reference path on the host: n_envs envs x steps steps over `threads` threads
    (ctypes releases the GIL).  Returns env-steps/s."""
    import ctypes as C
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from gym_drone2d_activeperception_b200 import Params
    p = Params(debug=False, **pk)
    worlds = make_worlds(pk, seed0 + np.arange(n_envs))
    op = oracle.make_params(dt=p.dt, map_scale=p.map_scale, map_size=p.map_size, agent_radius=p.agent_radius,
                            drone_max_acceleration=p.drone_max_acceleration, drone_radius=p.drone_radius,
                            drone_max_yaw_speed=p.drone_max_yaw_speed, drone_view_depth=p.drone_view_depth,
                            drone_view_range=p.drone_view_range, max_flight_time=p.max_flight_time, var_cam=p.var_cam,
                            drone_max_speed=p.drone_max_speed, planner=p.planner)
    envs = [oracle.OracleEnv(op, worlds["agent_pos"][i], worlds["agent_pref"][i], worlds["agent_radius"][i],
                             worlds["gt_grid"][i], worlds["tracker_radius"][i], drone=worlds["drone_pose"][i],
                             targets=p.target_list) for i in range(n_envs)]
    if p.motion_profile == "RVO":
        for i, e in enumerate(envs):
            e.set_rvo(worlds["agent_vel"][i], worlds["obstacles"][i])
    L = oracle.lib()
    table = np.arange(-80, 80, 80 / 3) / 80
    rng = np.random.RandomState(0)
    slices = np.array_split(np.arange(n_envs), threads)

    def run_slice(idx, nsteps):
        arr = (C.POINTER(oracle.Env) * len(idx))(*[envs[i]._ptr for i in idx])
        if oxford:      # Oxford.plan picks every action (config 4 workload)
            L.d2do_run_many_oxford(arr, len(idx), nsteps)
            return
        acts = np.ascontiguousarray(table[rng.randint(0, 6, (len(idx), nsteps))])
        L.d2do_run_many(arr, len(idx), nsteps, acts.ctypes.data_as(C.POINTER(C.c_double)))

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda s: run_slice(s, 2), slices))            # warm-up
        t0 = time.perf_counter()
        list(ex.map(lambda s: run_slice(s, steps), slices))
        dt = time.perf_counter() - t0
    for e in envs:
        e.close()
    return n_envs * steps / dt, dt


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python reference cannot
    travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = len(os.sched_getaffinity(0))
    n_envs = 4096 if cfg["envs"] >= 4096 else cfg["envs"]
    steps_per = 200
    vals = []
    t_all = time.perf_counter()
    for it in range(args.warmup + args.steps):
        v, dt = cpu_port_run({k: v for k, v in cfg["params"].items() if k != "gaze_method"}, n_envs, steps_per, cores,
                             seed0=1 + it, oxford=args.gaze == "Oxford")
        if it >= args.warmup:
            vals.append((n_envs * steps_per, dt))
        if time.perf_counter() - t_all > 150 and len(vals) >= 1:
            break
    tot_steps = sum(v[0] for v in vals)
    tot_t = sum(v[1] for v in vals)
    value = tot_steps / tot_t
    sample = "%d envs x %d steps per bench step, %d bench steps, %d threads" % (n_envs, steps_per, len(vals), cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(vals)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "rays_per_env_step": N_RAYS},
            "rays_per_sec": value * N_RAYS,
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else that libraries print there (NCCL's version banner
    under torchrun, for one) is sent to stderr by re-pointing fd 1; the JSON line goes to the saved descriptor."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the config's)")
    ap.add_argument("--envs-per-block", type=int, default=0)
    ap.add_argument("--unique-worlds", type=int, default=0, help="cap distinct generated worlds (0 = all unique up to 8192)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--burn-in", type=int, default=1000, help="untimed steps per replica before the timed region")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps (reported in config)")
    ap.add_argument("--strip-width", type=int, default=10, help="ray strip width: rays = ceil(500 / strip_width) (config 5 sweep: 10/5/2)")
    ap.add_argument("--view-range", type=int, default=0, help="drone_view_range in degrees (config 5 sweep: 90/180/360)")
    ap.add_argument("--planner", default=None, choices=["NoMove", "Primitive"], help="override the config's planner")
    ap.add_argument("--motion-profile", default=None, choices=["CVM", "RVO"], help="agent motion profile (default CVM)")
    ap.add_argument("--e2e-full-copy", action="store_true", help="e2e leg with plain D2H copies instead of the zero-copy mirror")
    ap.add_argument("--gaze", default=None, choices=["scripted", "Oxford"],
                    help="scripted: random actions from the Oxford action set; Oxford: d2d_plan_oxford every step")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.gaze is None:
        args.gaze = cfg.get("gaze", "scripted")
    if args.planner:
        cfg["params"] = dict(cfg["params"], planner=args.planner)
        if args.planner != CONFIGS[args.config]["params"]["planner"]:
            cfg["name"] += " [planner=%s]" % args.planner
    if args.motion_profile and args.motion_profile != "CVM":
        cfg["params"] = dict(cfg["params"], motion_profile=args.motion_profile)
        cfg["name"] += " [motion_profile=%s]" % args.motion_profile
    if args.view_range:
        cfg["params"] = dict(cfg["params"], drone_view_range=args.view_range)
        cfg["name"] += " [view_range=%d]" % args.view_range
    if args.strip_width != 10:
        cfg["name"] += " [strip_width=%d]" % args.strip_width
    if args.gaze == "Oxford":
        cfg["params"] = dict(cfg["params"], gaze_method="Oxford")
        if CONFIGS[args.config].get("gaze") != "Oxford":
            cfg["name"] += " [gaze=Oxford on device]"
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.envs or cfg["envs"]
    pk = cfg["params"]
    unique = args.unique_worlds or min(B, 16384)
    # worlds are generated before CUDA is touched (fork-based pool)
    seeds = pk["map_id"] + rank * B + np.arange(B)
    t0 = time.perf_counter()
    worlds = make_worlds(pk, seeds, unique=unique)
    t_world = time.perf_counter() - t0

    import torch
    import torch.distributed as dist
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    p = Params(debug=False, **pk)
    use_ox = args.gaze == "Oxford"
    def make_env():
        return Drone2DVecEnv(p, B, seeds=seeds, worlds=worlds, device=dev, auto_reset=True, trackers=True, oxford=use_ox,
                             envs_per_block=args.envs_per_block, strip_width=args.strip_width)
    env = make_env()
    n_rays = int(env.cfg.n_rays)
    N = env.num_agents
    K, W = args.steps, args.warmup
    # Cold-L2 rule: consecutive timed steps must not find their inputs in the 126 MB L2.  The timed region steps R
    # independent replicas of the batch round robin (same config, same worlds, own state and own action stream), R chosen
    # so that the state touched between two visits of a replica is >= 256 MiB; big batches (R = 1) already exceed L2.
    touched = B * (D2D_STATE_BYTES + 56 * N)       # + agents (40 B) and tracker flags / hit mask per agent; = ncu's DRAM read
    R = 1 if args.no_flush else max(1, min(64, -(-(256 << 20) // touched)))
    envs = [env] + [make_env() for _ in range(R - 1)]
    ox_out = torch.empty(B, dtype=torch.float64, device=dev)

    def do_step(a, e=None):
        e = env if e is None else e
        if use_ox:
            e.step(e.plan_oxford(ox_out))
        else:
            e.step(a)
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    actions = table[torch.randint(0, 6, (K + W, B), device=dev, generator=gen)].contiguous()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: EXACTLY K steps back to back (one CUDA graph of the K step launches on the launch
    #      stream), bracketed by barrier + synchronize and one CUDA-event pair
    for t in range(max(W, R)):                      # eager warm-up: every replica steps at least once
        do_step(actions[t % (K + W)], envs[t % R])
    # burn-in (untimed): episodes last up to 800 steps, so the mix of fresh / explored / tracking envs -- and with it the
    # cost of a step -- only becomes stationary after about a thousand steps (1000 steps with auto-reset)
    for t in range(args.burn_in):
        for e in envs:
            do_step(actions[t % (K + W)], e)
    barrier()
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    l0 = sum(e.launch_count() for e in envs)
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for t in range(K):
                do_step(actions[W + t], envs[t % R])
    launches = sum(e.launch_count() for e in envs) - l0          # kernel nodes of the graph == launches per K steps
    barrier()
    graph.replay()                                  # untimed: graph upload + K more warm-up steps
    barrier()
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    ev0.record()
    graph.replay()
    ev1.record()
    barrier()
    wall = time.perf_counter() - wall0
    total_ms = float(ev0.elapsed_time(ev1))
    repeats = []
    for _ in range(4):                              # run-to-run spread of the same K-step region (reported, not used)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        repeats.append(float(e0.elapsed_time(e1)) / K)
    # ---- secondary: the same step timed one launch at a time (CUDA events around every step, 256 MiB L2 flush in
    #      between, untimed): includes ~3-4 us of event / launch gap per step; kept for the per-step distribution
    Kp = min(K, 100)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pe0 = [torch.cuda.Event(enable_timing=True) for _ in range(Kp)]
    pe1 = [torch.cuda.Event(enable_timing=True) for _ in range(Kp)]
    for t in range(Kp):
        flush.fill_(t & 0xFF)
        pe0[t].record()
        do_step(actions[W + t])
        pe1[t].record()
    barrier()
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(pe0, pe1)])
    del flush
    # ---- end to end through the C ABI with pinned host buffers
    a_host = actions[W:].cpu().pin_memory()
    lm_host = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory()
    yaw_host = torch.empty((B,), dtype=torch.float32).pin_memory()
    done_host = torch.empty((B,), dtype=torch.uint8).pin_memory()
    Ke = min(K, 100)
    a_rows = [a_host[t] for t in range(Ke)]          # views of the pinned action buffer, one per step
    stage = env.buffer("actions_staging")

    def e2e_step(t):
        if use_ox:      # the policy runs on the device: its actions never leave the GPU, the observation still does
            env.plan_oxford(stage)
            env.step_host(None, lm_host, yaw_host, done_host)
        else:
            env.step_host(a_rows[t], lm_host, yaw_host, done_host)

    def e2e_loop():
        for t in range(3):
            e2e_step(t)
        barrier()
        st0 = env.stats()
        t0 = time.perf_counter()
        for t in range(Ke):
            e2e_step(t)
        barrier()
        return time.perf_counter() - t0, env.stats() - st0

    # (a) plain copies: every step moves the whole observation tensor device -> host
    e2e_copy_s, _ = e2e_loop()
    # (b) the library's zero-copy mirror (d2d_bind_host_mirror): the kernels store the observation bytes that change
    #     straight into the pinned host buffers; the step call copies nothing back.  This is the headline e2e.
    if args.e2e_full_copy:
        e2e_s, mirror_bytes = e2e_copy_s, None
    else:
        env.bind_host_mirror(lm_host, yaw_host, done_host)
        e2e_s, dst = e2e_loop()
        mirror_bytes = float(dst[14]) / Ke
        assert torch.equal(lm_host, env.buffer("local_map").cpu()) and torch.equal(done_host, env.buffer("done").cpu())
        env.bind_host_mirror(None, None, None)
    # clock / throttle sampling needs a loaded window much longer than the millisecond-scale timed region:
    # keep stepping back to back (untimed, same kernel) for ~1.5 s while nvidia-smi samples every 100 ms
    t_probe = time.perf_counter()
    sampler.mark = len(sampler.rows)
    while time.perf_counter() - t_probe < 1.5:
        for t in range(50):
            do_step(actions[W + (t % K)])
        torch.cuda.synchronize()
    clocks = sampler.stop()

    # max over ranks (device time), whole-job throughput
    tt = torch.tensor([total_ms, e2e_s * 1e3, e2e_copy_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max, e2e_copy_ms_max = float(tt[0]), float(tt[1]), float(tt[2])
    # the one collective of the path: all-reduce of the episode statistics
    stats = torch.as_tensor(sum(np.asarray(e.stats(), dtype=np.int64) for e in envs), device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    value = world * B * K / (total_ms_max * 1e-3)
    e2e = world * B * Ke / (e2e_ms_max * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak()
        kern_ms = total_ms_max / K            # average launch-to-launch duration inside the timed region
        if cfg["params"]["planner"] == "NoMove":
            kernel_name = "d2d_step_fused_warp_kernel (1 launch/step; K launches back to back, one event pair)"
        else:
            kernel_name = "d2d_step_prim_warp_kernel + d2d_plan_kernel + d2d_step_post_list_kernel" + \
                          (" + d2d_oxford_kernel" if use_ox else "") + " (whole step timed)"
        bytes_launch = algorithmic_bytes(N) * B
        achieved = bytes_launch / (kern_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_config%d.json" % args.config)
        if os.path.isfile(tpath) and cfg["params"]["planner"] == "NoMove":     # captured for the fused NoMove kernel only
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "envs_per_gpu": B, "agents_per_env": N, "rays_per_env_step": n_rays,
                       "planner": cfg["params"]["planner"], "gaze": args.gaze, "trackers": True, "auto_reset": True,
                       "unique_worlds_per_gpu": int(unique), "world_gen_s": round(t_world, 2),
                       "l2": "no rotation (--no-flush): state stays L2-resident" if args.no_flush else
                             "inputs larger than L2: %d replica(s) of the batch stepped round robin, %.0f MB of state "
                             "touched between two visits of a replica (L2 = 126 MB)" % (R, R * touched / 1e6),
                       "timing": "K steps back to back as one CUDA graph (K step launches), one CUDA-event pair, "
                                 "barrier + synchronize on both sides",
                       "replicas": R, "burn_in_steps_per_replica": args.burn_in,
                       "envs_per_block": env.cfg.envs_per_block, "parallelism": "env-sharded x%d" % world},
            "rays_per_sec": value * n_rays,
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": 0 if use_ox else B * 8,
                    "d2h_bytes_per_step": B * (1089 + 4 + 1) if mirror_bytes is None else int(round(B * 5 + mirror_bytes)),
                    "steps": Ke,
                    "transport": "cudaMemcpyAsync of the whole observation every step" if mirror_bytes is None else
                                 "d2d_bind_host_mirror: kernels store changed observation bytes + yaw + done straight into "
                                 "the pinned host buffers (bytes counted on the device, mean per step, this rank)",
                    "full_copy": {"value": world * B * Ke / (e2e_copy_ms_max * 1e-3), "d2h_bytes_per_step": B * (1089 + 4 + 1)}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_env_step": algorithmic_bytes(N),
                         "kernel": kernel_name, "kernel_ms": kern_ms},
            "repeat_ms_per_step": repeats,
            "per_step_events": {"what": "same step, one launch at a time: CUDA events around every step, 256 MiB L2 flush "
                                        "(untimed) in between; includes the event / launch gap of each step",
                                "steps": int(Kp), "ms_min_med_max": [float(step_ms.min()), float(np.median(step_ms)),
                                                                     float(step_ms.max())],
                                "value": world * B / (float(np.mean(step_ms)) * 1e-3)},
            "wall_s_timed_region": wall,
            "episode_stats": {n: int(v) for n, v in zip(
                ["env_steps", "episodes", "success", "static_collision", "dynamic_collision", "freezing", "dead_lock",
                 "flight_steps", "grid_discovered", "agents_tracked", "tracked_steps", "plans", "plan_failures", "replans"],
                stats.tolist()[:14])},
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle
            oracle.build()
            cores = len(os.sched_getaffinity(0))
            # bounded sample of the same workload, sized for ~10 s of CPU work from a short calibration run
            n_envs = min(B, 4096)
            v0, dt0 = cpu_port_run(pk, n_envs, 20, cores, oxford=use_ox)
            steps_c = int(min(20000, max(50, 25.0 * v0 / n_envs)))   # the 20-step calibration under-reads the rate ~2x
            v, dt = cpu_port_run(pk, n_envs, steps_c, cores, oxford=use_ox)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d envs x %d steps of the same workload (%.1f s)" % (n_envs, steps_c, dt)}
        emit(line)
    del graph
    for e in envs:
        e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
