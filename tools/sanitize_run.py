"""Small workload for compute-sanitizer (GPU box):  compute-sanitizer --tool memcheck python tools/sanitize_run.py
Steps a NoMove batch (fused warp kernel, both march variants: seed 8 has a broken border ring), a Primitive + Oxford batch and
an RVO batch, a Primitive + Owl batch, a Jerk_Primitive batch, a pipelined NoMove run (d2d_step_pipelined: resident gated kernel
with its courier block, deferred mirror stores; one-step runs, since sanitizers make launches blocking) and d2d_rollout for a few dozen steps with auto-reset, host buffers bound, and prints the episode statistics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from gym_drone2d_activeperception_b200 import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80)
    g = torch.Generator().manual_seed(0)
    only_resident = "--resident-only" in sys.argv       # just the round-2 resident kernels (short racecheck reports)
    for name, kw, B, steps, ox in [] if only_resident else [("NoMove", dict(planner="NoMove", map_id=1), 40, 40, False),
                                   ("Primitive+Oxford", dict(planner="Primitive", gaze_method="Oxford", map_id=1), 24, 40, True),
                                   ("RVO", dict(planner="NoMove", motion_profile="RVO", map_id=3), 8, 10, False)]:
        p = Params(debug=False, agent_number=10, agent_radius=15, agent_max_speed=20, **kw)
        env = Drone2DVecEnv(p, B, seeds=1 + np.arange(B), device="cuda:0", auto_reset=True, oxford=ox)
        lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory()
        yaw = torch.empty((B,), dtype=torch.float32).pin_memory()
        dn = torch.empty((B,), dtype=torch.uint8).pin_memory()
        env.bind_host_mirror(lm, yaw, dn)
        for t in range(steps):
            if ox:
                env.plan_oxford(env.buffer("actions_staging"))
                env.step_host(None, lm, yaw, dn)
            else:
                env.step_host(table[torch.randint(0, 6, (B,), generator=g)].contiguous().pin_memory(), lm, yaw, dn)
        if ox:      # d2d_step_plan_oxford: A* searches + completion + scoring of the planning envs on the side stream
            env.bind_host_mirror(None, None, None)
            a = env.plan_oxford()
            for t in range(20):
                a = env.step_plan_oxford(a, out=a)
            torch.cuda.synchronize()
        print(name, env.stats()[:8])
        env.close()
    # round 2: Owl policy, Jerk_Primitive planner, pipelined bound stepping
    if not only_resident:
        run_owl_jerk(Params, Drone2DVecEnv)
    B = 40
    p = Params(debug=False, agent_number=10, agent_radius=15, agent_max_speed=20, planner="NoMove", map_id=1)
    env = Drone2DVecEnv(p, B, seeds=1 + np.arange(B), device="cuda:0", auto_reset=True)
    lm = torch.empty((B, 1, 33, 33), dtype=torch.uint8).pin_memory()
    yaw = torch.empty((B,), dtype=torch.float32).pin_memory()
    dn = torch.empty((B,), dtype=torch.uint8).pin_memory()
    acts = torch.zeros(B, dtype=torch.float64).pin_memory()
    env.bind_host_io(acts, lm, yaw, dn)
    for t in range(40):
        acts.copy_(table[torch.randint(0, 6, (B,), generator=g)])
        env.step_pipelined(prelaunch_next=False)     # sanitizers make launches blocking: a pre-launched kernel would starve
    env.bind_host_io(None, None, None, None)
    print("NoMove pipelined", env.stats()[:8])
    # d2d_rollout: 12 steps in one launch (28-warp blocks, one block barrier per step), across auto-resets
    env.rollout(table[torch.randint(0, 6, (12, B), generator=g)].cuda())
    env.rollout(table[torch.randint(0, 6, (12, B), generator=g)].cuda())
    print("NoMove rollout", env.stats()[:8])
    env.close()


def run_owl_jerk(Params, Drone2DVecEnv):
    import torch
    p = Params(debug=False, agent_number=10, agent_radius=15, agent_max_speed=20, planner="Primitive", gaze_method="Owl", map_id=1)
    env = Drone2DVecEnv(p, 16, seeds=1 + np.arange(16), device="cuda:0", auto_reset=True)
    for t in range(40):
        env.step(env.plan_gaze("Owl"))
    print("Primitive+Owl", env.stats()[:8])
    env.close()
    p = Params(debug=False, agent_number=10, agent_radius=15, agent_max_speed=20, planner="Jerk_Primitive", gaze_method="LookAhead", map_id=1)
    env = Drone2DVecEnv(p, 16, seeds=1 + np.arange(16), device="cuda:0", auto_reset=True)
    for t in range(40):
        env.step(env.plan_gaze("LookAhead"))
    print("Jerk_Primitive+LookAhead", env.stats()[:8])
    env.close()


if __name__ == "__main__":
    main()
