"""What-if: cost of a step when EVERY env is re-initialised every step (max_flight_time = 0.1 -> freezing -> done -> auto-reset),
against the stationary mix: python tools/fresh_step_bench.py [max_flight_time]  (GPU box only)"""
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from gym_drone2d_activeperception_b200 import Params
from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
cfg = bench.make_cfg(2); B, pk = cfg["envs"], dict(cfg["params"])
mft = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
pk["max_flight_time"] = mft
worlds = bench.make_worlds(pk, pk["map_id"] + np.arange(B))
env = Drone2DVecEnv(Params(debug=False, **pk), B, worlds=worlds, device="cuda:0", auto_reset=True)
table = torch.as_tensor(np.arange(-80, 80, 80 / 3) / 80, device="cuda:0")
acts = table[torch.randint(0, 6, (64, B), device="cuda:0")].contiguous()
for r in range(4): env.rollout(acts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for r in range(10):
    e0.record(); env.rollout(acts); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 64 * 1e3)
st = env.stats()
print(json.dumps({"max_flight_time": mft, "us_per_step": float(np.median(ts)), "episodes_per_env_step": float(st[1]) / float(st[0])}))
