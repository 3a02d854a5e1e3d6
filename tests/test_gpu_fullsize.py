"""Size-independent properties at BASELINE.json's full batch size (config 2: 4096 envs), where stepping the oracle
for every env is too slow: batch-composition invariance, observation == zero-padded crop of the belief grid, belief
monotonicity, flag consistency, statistics bookkeeping, and an oracle spot check on a few envs of the big batch."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _crop(belief, x, y):
    """torch restatement of Drone2D.get_local_map (utils.py:780-784) for a batch: [B,50,50] -> [B,33,33]."""
    B = belief.shape[0]
    pad = torch.zeros((B, 82, 82), dtype=torch.uint8, device=belief.device)
    pad[:, 16:66, 16:66] = belief
    ix = torch.div(x, 10, rounding_mode="floor").long()
    iy = torch.div(y, 10, rounding_mode="floor").long()
    ar = torch.arange(33, device=belief.device)
    rows = (ix[:, None] + ar[None, :])[:, :, None].expand(B, 33, 33)
    cols = (iy[:, None] + ar[None, :])[:, None, :].expand(B, 33, 33)
    return pad[torch.arange(B, device=belief.device)[:, None, None], rows, cols]


@pytest.mark.parametrize("planner,steps", [("NoMove", 120), ("Primitive", 80)])
def test_full_batch_properties(planner, steps):
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    from gym_drone2d_activeperception_b200.world import generate_worlds
    B, SUB0, SUBN = 4096, 1000, 96
    p = Params(debug=False, planner=planner, gaze_method="Oxford", map_id=1, agent_number=10, agent_radius=15,
               agent_max_speed=20)
    seeds = 1 + np.arange(B)
    uniq = generate_worlds(p, seeds[:512])
    worlds = {k: np.concatenate([v] * (B // 512)) for k, v in uniq.items()}     # 512 distinct worlds tiled 8x
    big = Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", auto_reset=True, oxford=(planner == "Primitive"))
    sub_w = {k: np.ascontiguousarray(v[SUB0:SUB0 + SUBN]) for k, v in worlds.items()}
    sub = Drone2DVecEnv(p, SUBN, worlds=sub_w, device="cuda:0", auto_reset=True, oxford=(planner == "Primitive"))
    oracles = [util.oracle_env_from_world(p, worlds, i) for i in (0, 777, 4095)] if planner == "NoMove" else []
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(5)
    table = torch.as_tensor(util.action_table(), device="cuda:0")
    prev_belief = big.buffer("belief").clone()
    prev_done = torch.zeros(B, dtype=torch.bool, device="cuda:0")
    for t in range(steps):
        if planner == "Primitive":
            a = big.plan_oxford().clone()
            a_sub = sub.plan_oxford()
            assert torch.equal(a[SUB0:SUB0 + SUBN], a_sub), ("oxford action depends on batch composition", t)
        else:
            a = table[torch.randint(0, 6, (B,), device="cuda:0", generator=gen)]
        obs, rew, done, info = big.step(a)
        sub.step(a[SUB0:SUB0 + SUBN].contiguous())
        bel = big.buffer("belief")
        # (1) tiled copies of the same world evolve identically (no cross-env interference, any warp/block placement)
        if planner == "NoMove":
            pass  # actions differ per env, copies diverge by design
        # (2) observation == crop of the belief grid at the drone's cell; swep_map aliases local_map
        x, y = big.buffer("drone_x"), big.buffer("drone_y")
        inside = (x >= 0) & (x < 500) & (y >= 0) & (y < 500)
        crop = _crop(bel, x.clamp(0, 499.9), y.clamp(0, 499.9))
        assert torch.equal(obs["local_map"][:, 0][inside], crop[inside]), ("local_map != crop(belief)", t)
        assert obs["swep_map"].data_ptr() == obs["local_map"].data_ptr() and float(rew.abs().sum()) == 0.0
        # (3) belief is monotone within an episode: an explored cell never changes value
        same_episode = ~prev_done
        changed = (prev_belief != 0) & (bel != prev_belief)
        assert not bool(changed[same_episode].any()), ("belief cell changed value", t)
        assert int(bel.max()) <= 2
        # (4) flag consistency (drone_v2.py:228-231)
        col, dead, frz = big.buffer("collision_flag"), big.buffer("dead_lock_flag"), big.buffer("freezing_flag")
        sm = big.buffer("state_machine")
        expect_done = (col != 0) | (dead != 0) | (frz != 0) | (sm == 1)
        assert torch.equal(done.bool(), expect_done), ("done flag inconsistent", t)
        assert not bool(((dead != 0) & (frz != 0)).any())
        # (5) batch-composition invariance: the sub-batch reproduces envs [SUB0, SUB0+SUBN) of the big batch exactly
        for name in ("belief", "local_map", "drone_x", "drone_y", "drone_yaw", "agent_pos", "done", "collision_flag",
                     "tracker_active", "tracker_mu", "steps", "traj_nseg", "traj_cursor"):
            assert torch.equal(big.buffer(name)[SUB0:SUB0 + SUBN], sub.buffer(name)), (name, t)
        # (6) oracle spot check inside the big batch
        for o, i in zip(oracles, (0, 777, 4095)):
            if o.c.done:
                continue
            o.step(float(a[i]))
            assert np.array_equal(bel[i].cpu().numpy(), o.belief) and int(done[i]) == o.c.done, ("oracle spot", t, i)
        prev_belief = bel.clone()
        prev_done = done.bool().clone()
    st = big.stats()
    assert st[0] == B * steps and st[1] == st[2] + st[3] + st[4] + st[5] + st[6]
    big.close()
    sub.close()
