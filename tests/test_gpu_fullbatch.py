"""Full-batch parity on the kernels bench.py times (default warp-per-env kernels, envs_per_block = 0, auto-reset on):
EVERY env of BASELINE config 2 (4096 envs) for 300 steps, and 4096-env slices of the config 3 / 4 / 5 batches
(Primitive planner, Primitive + Oxford on the device, 96 agents), against the oracle stepped beside the GPU batch
(oracle.OracleBatch: thread pool over d2do_step_batch, same auto-reset rule).

Belief grids, hit masks, flags, done, observation, integer bookkeeping: bit-exact, every env, every step.
Continuous state (drone / agents every step, tracker mu / Sigma every `trk_every` steps): <= 1e-9 relative (north_star)."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL = 1e-9

CASES = {
    # BASELINE.json configs[1]: the bench.py default workload, all 4096 envs
    "cfg2_all4096": dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20, planner="NoMove",
                         gaze="scripted", B=4096, steps=300, trk_every=5),
    # same batch with the drones scattered over the map (externally set poses, as the metric scripts do): collisions,
    # walls in view, auto-resets from step 1 on
    "cfg2_scattered": dict(static_map="maps/empty_map.npy", agent_number=10, agent_radius=15, agent_max_speed=20,
                           planner="NoMove", gaze="scripted", B=4096, steps=120, trk_every=5, scatter=True),
    # configs[2]: random_map_0, 142 agents, Primitive planner checks on the device
    "cfg3_slice4096": dict(static_map="maps/random_map_0.npy", agent_number=20, agent_radius=15, agent_max_speed=40,
                           planner="Primitive", gaze="scripted", B=4096, steps=100, trk_every=20),
    # configs[3]: obstacle_map, Primitive + Oxford gaze scoring on the device
    "cfg4_slice4096": dict(static_map="maps/obstacle_map.npy", agent_number=10, agent_radius=10, agent_max_speed=20,
                           planner="Primitive", gaze="Oxford", B=4096, steps=260, trk_every=10),
    # configs[4]: shaped_obstacle_map, 96 agents, perception + dynamics
    "cfg5_slice4096": dict(static_map="maps/shaped_obstacle_map.npy", agent_number=50, agent_radius=10, agent_max_speed=40,
                           planner="NoMove", gaze="scripted", B=4096, steps=100, trk_every=20),
}


def _first(mask):
    return np.nonzero(mask)[0][:8].tolist()


@pytest.mark.parametrize("name", list(CASES))
def test_full_batch_matches_oracle(name):
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    from gym_drone2d_activeperception_b200.world import generate_worlds
    cfg = CASES[name]
    B, steps = cfg["B"], cfg["steps"]
    use_ox = cfg["gaze"] == "Oxford"
    p = Params(debug=False, planner=cfg["planner"], gaze_method="Oxford" if use_ox else "NoControl", map_id=1,
               static_map=cfg["static_map"], agent_number=cfg["agent_number"], agent_radius=cfg["agent_radius"],
               agent_max_speed=cfg["agent_max_speed"])
    worlds = generate_worlds(p, 1 + np.arange(B))                 # seeds map_id + env index
    env = Drone2DVecEnv(p, B, worlds=worlds, device="cuda:0", auto_reset=True, oxford=use_ox)   # envs_per_block = 0
    assert env.cfg.envs_per_block == 0
    n = env.num_agents
    poses = None
    rng = np.random.RandomState(11)
    if cfg.get("scatter"):
        poses = worlds["drone_pose"].copy()
        poses[:, 0] = rng.uniform(12, 488, B); poses[:, 1] = rng.uniform(12, 488, B); poses[:, 2] = rng.uniform(0, 360, B)
        poses[::7, :2] = np.round(poses[::7, :2])                 # some on the integer lattice
        env.set_drone_pose(poses)
        # the reset pose of the batched env is its initial pose: keep the oracle's snapshot consistent
        env.buffer("drone_pose0").copy_(torch.as_tensor(poses.T.copy(), device="cuda:0"))
    ob = util.oracle_batch(p, worlds, poses)
    n_way = int(env.cfg.n_way) if cfg["planner"] == "Primitive" else 0
    fields = util.BATCH_FIELDS + (util.PLANNER_FIELDS if n_way else [])
    table = util.action_table()
    episodes = 0
    for t in range(steps):
        if use_ox:
            a_dev = env.plan_oxford()
            want, _ = ob.step(policy="Oxford", auto_reset=True)
            got = a_dev.cpu().numpy()
            assert np.array_equal(got, want), ("oxford action", t, _first(got != want))
        else:
            acts = table[rng.randint(0, 6, B)]
            a_dev = torch.as_tensor(acts, device="cuda:0")
            ob.step(acts, auto_reset=True)
        env.step(a_dev)
        trk = (t % cfg["trk_every"] == 0) or t == steps - 1
        h = util.gpu_fields(env, fields + (util.TRACKER_FIELDS if trk else []))
        o = ob.gather(trackers=trk)
        d, r = util.batch_mismatch(h, o, n, trackers=trk, planner=n_way)
        for k, m in d.items():
            assert not m.any(), (name, k, "step", t, "envs", _first(m), "of", int(m.sum()))
        for k, v in r.items():
            assert float(v.max()) <= RTOL, (name, k, "step", t, "env", int(v.argmax()), float(v.max()))
        episodes += int(o["done"].sum())
    st = env.stats()
    assert st[0] == B * steps and st[1] == episodes
    if cfg.get("scatter") or use_ox:
        assert episodes > B // 8, "the case must exercise auto-reset on the default kernels"
    ob.close()
    env.close()
