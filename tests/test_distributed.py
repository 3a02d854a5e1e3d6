"""N > 1 host logic on CPU: world_size-2 gloo process group, shard invariance, statistics all-reduce."""
import os
import socket

import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, util.ROOT)
    sys.path.insert(0, os.path.join(util.ROOT, "oracle"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from gym_drone2d_activeperception_b200 import Params, generate_worlds
    from gym_drone2d_activeperception_b200 import distributed as D
    r, lr, w = D.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    B = 6
    p = Params(debug=False, planner="NoMove", map_id=40, agent_number=5, agent_radius=15, agent_max_speed=20)
    seeds = D.shard_seeds(p.map_id, B, rank)
    worlds = generate_worlds(p, seeds)
    # each rank steps its shard with the CPU oracle (stands in for the per-GPU batch) and builds the stats vector
    envs = [util.oracle_env_from_world(p, worlds, i) for i in range(B)]
    stats = np.zeros(16, dtype=np.int64)
    for t in range(30):
        for e in envs:
            e.step(1.0)
            stats[0] += 1
            if e.c.done:
                stats[1] += 1
                stats[4] += int(e.c.collision == 2)
    total = D.allreduce_stats(stats)
    tmax = D.max_over_ranks(float(rank + 1))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), seeds=seeds, pos=worlds["agent_pos"], stats=stats, total=total,
             tmax=tmax)
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_stats_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # disjoint, contiguous seeds; the all-reduced vector is the host-side sum on both ranks
    assert r0["seeds"].tolist() == list(range(40, 46)) and r1["seeds"].tolist() == list(range(46, 52))
    assert np.array_equal(r0["total"], r0["stats"] + r1["stats"]) and np.array_equal(r1["total"], r0["total"])
    assert r0["total"][0] == 2 * 6 * 30
    assert float(r0["tmax"]) == 2.0 and float(r1["tmax"]) == 2.0
    # shard invariance: global env 46..51 generated on "rank 1" equals the same seeds generated in one piece
    from gym_drone2d_activeperception_b200 import Params, generate_worlds
    p = Params(debug=False, planner="NoMove", map_id=40, agent_number=5, agent_radius=15, agent_max_speed=20)
    whole = generate_worlds(p, 40 + np.arange(12))
    assert np.array_equal(whole["agent_pos"][:6], r0["pos"]) and np.array_equal(whole["agent_pos"][6:], r1["pos"])


def test_shard_range_covers_everything():
    from gym_drone2d_activeperception_b200.distributed import shard_range
    for total in (1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
