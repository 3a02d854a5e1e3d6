import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import util
from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
g = util.load_golden(util.golden_files("episode_cfg1")[0])
p = util.params_from_golden(g)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
env = Drone2DVecEnv(p, B, worlds=util.world_from_golden(g, B), device="cuda:0", auto_reset=False, oxford=True)
for t in range(len(g["done"])):
    a = env.plan_oxford()
    env.step(a)
    torch.cuda.synchronize()
    lm = env.buffer("local_map").cpu().numpy()
    bel = env.buffer("belief").cpu().numpy()
    x = env.buffer("drone_x").cpu().numpy(); y = env.buffer("drone_y").cpu().numpy()
    bad = False
    for i in range(B):
        if not np.array_equal(lm[i, 0], g["local_map"][t]):
            d = np.argwhere(lm[i, 0] != g["local_map"][t])
            print("t", t, "env", i, "drone", x[i], y[i], "golden drone", g["drone"][t], "n diff", len(d), d[:6].tolist(),
                  "belief equal", np.array_equal(bel[i], g["belief"][t]), "action", float(a[i]), g["action"][t])
            bad = True
    if bad:
        print("row0 env0", lm[0,0,0,:6], "env1", lm[1,0,0,:6], "env2", lm[2,0,0,:6], "golden", g["local_map"][t][0,:6])
        print("ix/iy", env.buffer("drone_x").cpu().numpy()//10, env.buffer("drone_y").cpu().numpy()//10)
        break
print("done at", t)
