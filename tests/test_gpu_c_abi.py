"""The C ABI driven from plain C: tests/helpers/c_caller.c is compiled with gcc against include/drone2d.h, linked to
libdrone2d.so, and must reproduce what the Python host gets for the same seeded batch (observation, yaw, done, statistics)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("planner", ["NoMove", "Primitive"])
def test_plain_c_caller_matches_python_host(planner, tmp_path):
    from gym_drone2d_activeperception_b200 import Params, generate_worlds, count_agents, _native
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv, make_config
    pkg = os.path.dirname(_native.LIB_PATH)
    exe = str(tmp_path / "c_caller")
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(util.ROOT, "include"),
                    os.path.join(util.ROOT, "tests", "helpers", "c_caller.c"), "-o", exe, "-L", pkg, "-ldrone2d",
                    "-Wl,-rpath," + pkg], check=True)
    B, T = 13, 90
    p = Params(debug=False, planner=planner, map_id=21, agent_number=10, agent_radius=15, agent_max_speed=40)
    N = count_agents(p)
    worlds = generate_worlds(p, 21 + np.arange(B))
    cfg = make_config(p, B, N, 0, auto_reset=True, trackers=True, oxford=False)
    acts = util.action_table()[np.random.RandomState(4).randint(0, 6, (T, B))]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([0x44324432, B, N, T], dtype=np.int32).tobytes())
        f.write(bytes(cfg))
        for k, dt in (("agent_pos", np.float64), ("agent_pref", np.float64), ("agent_radius", np.float64),
                      ("tracker_radius", np.float64), ("gt_grid", np.uint8), ("drone_pose", np.float64)):
            f.write(np.ascontiguousarray(worlds[k], dtype=dt).tobytes())
        f.write(np.ascontiguousarray(acts, dtype=np.float64).tobytes())
    r = subprocess.run([exe, fin, fout], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "c_caller ok" in r.stdout, r.stdout
    raw = np.fromfile(fout, dtype=np.uint8)
    per = B * 1089 + B * 4 + B
    assert raw.size == T * per + 8 * _native.NUM_STATS
    env = Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", auto_reset=True)
    for t in range(T):
        obs, _, done, _ = env.step(torch.as_tensor(acts[t], device="cuda:0"))
        blk = raw[t * per:(t + 1) * per]
        assert np.array_equal(blk[:B * 1089].reshape(B, 1, 33, 33), obs["local_map"].cpu().numpy()), t
        assert np.array_equal(blk[B * 1089:B * 1089 + 4 * B].view(np.float32), obs["yaw_angle"].cpu().numpy()[:, 0]), t
        assert np.array_equal(blk[B * 1093:], done.cpu().numpy()), t
    assert np.array_equal(raw[T * per:].view(np.int64), env.stats())
    env.close()
