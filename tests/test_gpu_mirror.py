"""Zero-copy host mirror of the observation (d2d_bind_host_mirror): the pinned host buffers must hold exactly what the
device tensors hold after every d2d_step_host, on every kernel path (fused NoMove with in-place patching, Primitive with
moving windows, block-per-E-envs kernels), across auto-resets, eager resets, pose writes and re-binds."""
import numpy as np
import pytest

import util

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(p, B, worlds, **kw):
    from gym_drone2d_activeperception_b200.vec_env import Drone2DVecEnv
    return Drone2DVecEnv(p, B, worlds=worlds, seeds=None if worlds is not None else 7 + np.arange(B), device="cuda:0", **kw)


def _pinned(B):
    return (torch.full((B, 1, 33, 33), 77, dtype=torch.uint8).pin_memory(),
            torch.full((B,), -1.0, dtype=torch.float32).pin_memory(),
            torch.full((B,), 9, dtype=torch.uint8).pin_memory())


def _check(env, lm, yaw, dn, tag):
    assert torch.equal(lm, env.buffer("local_map").cpu()), ("local_map", tag)
    assert torch.equal(yaw, env.buffer("yaw_angle").cpu()[:, 0]), ("yaw", tag)
    assert torch.equal(dn, env.buffer("done").cpu()), ("done", tag)


@pytest.mark.parametrize("planner,epb,B,steps", [("NoMove", 0, 37, 260), ("NoMove", 8, 21, 60), ("Primitive", 0, 23, 160),
                                                  ("Primitive", 8, 9, 60)])
def test_mirror_tracks_device_observation(planner, epb, B, steps):
    from gym_drone2d_activeperception_b200.params import Params
    p = Params(debug=False, planner=planner, map_id=7, agent_number=10, agent_radius=15, agent_max_speed=40)
    env = _env(p, B, None, auto_reset=True, oxford=False, envs_per_block=epb)
    lm, yaw, dn = _pinned(B)
    env.bind_host_mirror(lm, yaw, dn)
    table = torch.as_tensor(util.action_table())
    g = torch.Generator().manual_seed(3)
    acts = table[torch.randint(0, 6, (steps, B), generator=g)].contiguous().pin_memory()
    resets = 0
    for t in range(steps):
        env.step_host(acts[t], lm, yaw, dn)
        _check(env, lm, yaw, dn, t)
        resets += int(dn.sum())
        if t == steps // 2:                       # eager reset of some envs: the mirror is refreshed by the next call
            mask = torch.zeros(B, dtype=torch.uint8, device="cuda:0")
            mask[::3] = 1
            env.reset(mask)
        if t == steps // 2 + 5:                   # pose write: the window is rebuilt, into the mirror too
            pose = np.stack([np.full(B, 250.0), np.full(B, 250.0), np.full(B, 45.0)], 1)
            env.set_drone_pose(pose)
    assert resets > 0, "the run must cross at least one auto-reset"
    mb = env.stats()[14]
    assert mb > 0
    if planner == "NoMove" and epb == 0:          # patched in place: far less than one window per env-step
        assert mb < 0.2 * B * steps * 1089
    env.close()


def test_mirror_partial_rebind_and_unbind():
    from gym_drone2d_activeperception_b200.params import Params
    B = 12
    p = Params(debug=False, planner="NoMove", map_id=11, agent_number=10, agent_radius=15, agent_max_speed=20)
    env = _env(p, B, None, auto_reset=True, oxford=False)
    lm, yaw, dn = _pinned(B)
    acts = torch.full((B,), 1.0 / 3, dtype=torch.float64).pin_memory()
    env.bind_host_mirror(lm, None, None)          # only the window is mirrored; yaw / done are copied as before
    for t in range(10):
        env.step_host(acts, lm, yaw, dn)
        _check(env, lm, yaw, dn, ("partial", t))
    lm2, yaw2, dn2 = _pinned(B)                   # other buffers than the bound ones: plain full copies
    env.step_host(acts, lm2, yaw2, dn2)
    _check(env, lm2, yaw2, dn2, "unbound buffers")
    _check(env, lm, yaw2, dn2, "bound mirror kept in sync by the kernels")
    env.bind_host_mirror(lm2, yaw2, dn2)          # re-bind
    for t in range(10):
        env.step_host(acts, lm2, yaw2, dn2)
        _check(env, lm2, yaw2, dn2, ("rebound", t))
    env.bind_host_mirror(None, None, None)
    snap = lm2.clone()
    env.step(torch.full((B,), 1.0, dtype=torch.float64, device="cuda:0"))
    torch.cuda.synchronize()
    assert torch.equal(lm2, snap), "an unbound buffer must not be written any more"
    env.close()


def test_mirror_rejects_pageable_memory():
    from gym_drone2d_activeperception_b200 import _native
    from gym_drone2d_activeperception_b200.params import Params
    import ctypes as C
    env = _env(Params(debug=False, planner="NoMove", agent_number=3), 4, None)
    with pytest.raises(ValueError):
        env.bind_host_mirror(torch.empty((4, 1, 33, 33), dtype=torch.uint8), None, None)
    pageable = np.empty(4 * 1089, dtype=np.uint8)
    rc = env._lib.d2d_bind_host_mirror(env._h, C.c_void_p(pageable.ctypes.data), None, None)
    assert rc == -1 and b"pinned" in env._lib.d2d_last_error(env._h)
    env.close()


@pytest.mark.parametrize("planner", ["NoMove", "Primitive"])
def test_step_host_action_sources_agree(planner):
    """d2d_step_host takes its actions from pinned host memory (read in place by the kernels), from pageable host memory
    (copied to the staging buffer first) or, with actions_host == NULL, from the device buffer "actions_staging": three
    envs fed the same actions through the three routes must stay identical."""
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200 import generate_worlds
    B, steps = 19, 120
    p = Params(debug=False, planner=planner, map_id=5, agent_number=10, agent_radius=15, agent_max_speed=40)
    worlds = generate_worlds(p, 5 + np.arange(B))
    envs = [_env(p, B, worlds, auto_reset=True, oxford=False) for _ in range(3)]
    table = util.action_table()
    rng = np.random.RandomState(2)
    outs = [_pinned(B) for _ in range(3)]
    for t in range(steps):
        a = np.ascontiguousarray(table[rng.randint(0, 6, B)])
        envs[0].step_host(torch.from_numpy(a).pin_memory(), *outs[0])          # pinned: zero-copy read
        envs[1].step_host(a, *outs[1])                                         # pageable numpy: staged copy
        envs[2].buffer("actions_staging").copy_(torch.from_numpy(a))           # already on the device
        envs[2].step_host(None, *outs[2])
        for k in (1, 2):
            for x, y in zip(outs[0], outs[k]):
                assert torch.equal(x, y), (planner, t, k)
    for name in ("belief", "drone_yaw", "agent_pos", "steps", "tracker_mu"):
        for k in (1, 2):
            assert torch.equal(envs[0].buffer(name), envs[k].buffer(name)), (name, k)
    for e in envs:
        e.close()


# knob: which transport d2d_step_pipelined uses for a NoMove batch.  Default: the resident kernel with a courier block (one SM is
# free: B <= 4116); D2D_NO_COURIER: resident kernel fed by the copy engine (also what a batch that fills all 148 SMs gets:
# B = 4144); D2D_NO_RESIDENT: one pre-launched kernel per step (also what a batch above one wave gets: B = 4200).
@pytest.mark.parametrize("planner,B,steps,pipelined,knob", [
    ("NoMove", 37, 300, False, None), ("NoMove", 37, 300, True, None), ("NoMove", 4096, 60, True, None),
    ("NoMove", 37, 120, True, "D2D_NO_COURIER"), ("NoMove", 4096, 48, True, "D2D_NO_COURIER"),
    ("NoMove", 37, 120, True, "D2D_NO_RESIDENT"), ("NoMove", 4096, 48, True, "D2D_NO_RESIDENT"),
    ("NoMove", 4144, 48, True, None), ("NoMove", 4200, 48, True, None),     # 48: no sync-refresh step lands on a t % 7 == 0 probe
    ("Primitive", 23, 120, False, None), ("Primitive", 23, 60, True, None)])
def test_bound_host_io_equals_step_host(planner, B, steps, pipelined, knob, monkeypatch):
    """d2d_bind_host_io + d2d_step_bound / d2d_step_pipelined against d2d_step_host on a twin env: host observation buffers
    and device state identical after every step, across auto-resets, an eager reset and a pose write (which end the pipelined
    run and force the synchronising refresh path once).  Pipelined: when the call returns, the next step's kernel is already
    running behind the current one (it owns the GPU until its actions arrive, so the twin runs afterwards, not interleaved),
    and the observation buffers must still hold THIS step's observation."""
    import time
    from gym_drone2d_activeperception_b200.params import Params
    from gym_drone2d_activeperception_b200 import generate_worlds
    if knob:
        monkeypatch.setenv(knob, "1")
    p = Params(debug=False, planner=planner, map_id=7, agent_number=10, agent_radius=15, agent_max_speed=40)
    worlds = generate_worlds(p, 7 + np.arange(min(B, 256)))
    worlds = {k: np.concatenate([v] * (-(-B // len(v))))[:B] for k, v in worlds.items()}
    table = torch.as_tensor(util.action_table())
    g = torch.Generator().manual_seed(5)
    acts = table[torch.randint(0, 6, (steps, B), generator=g)].contiguous()
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda:0")
    mask[::3] = 1
    pose = np.stack([np.full(B, 250.0), np.full(B, 250.0), np.full(B, 45.0)], 1)
    names = ("belief", "agent_pos", "steps", "tracker_mu", "drone_yaw", "local_map", "done")

    def run(bound):
        env = _env(p, B, worlds, auto_reset=True, oxford=False)
        lm, yaw, dn = _pinned(B)
        a_bound = torch.zeros(B, dtype=torch.float64).pin_memory()
        if bound:
            env.bind_host_io(a_bound, lm, yaw, dn)
        obs, state = [], []
        for t in range(steps):
            last = t in (steps // 2, steps // 2 + 5, steps - 1) or t % 20 == 0     # steps followed by another API call
            if not bound:
                env.step_host(acts[t].clone().pin_memory(), lm, yaw, dn)
            else:
                a_bound.copy_(acts[t])             # the caller rewrites the bound action buffer in place
                if pipelined:
                    env.step_pipelined(prelaunch_next=not last)
                else:
                    env.step_bound()
            obs.append((lm.clone(), yaw.clone(), dn.clone()))
            if bound and pipelined and not last and planner == "NoMove":
                if t % 7 == 0:                      # a pre-launched step is in flight: only the call that completes it is accepted
                    with pytest.raises(Exception):
                        env.stats()
                if t % 11 == 0:
                    time.sleep(0.002)               # let the pre-launched kernel reach its gate and wait there
                    assert torch.equal(lm, obs[-1][0]) and torch.equal(dn, obs[-1][2]), "observation changed before the next call"
            if last:
                _check(env, lm, yaw, dn, ("bound" if bound else "host", t))
                state.append({k: env.buffer(k).clone() for k in names})
            if t == steps // 2:
                env.reset(mask)
            if t == steps // 2 + 5:
                env.set_drone_pose(pose)
        if bound:
            env.bind_host_io(None, None, None, None)
            with pytest.raises(Exception):
                env.step_bound()
        env.close()
        return obs, state

    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):                   # the bound env runs on its own stream
        oa, sa = run(True)
    torch.cuda.synchronize()
    ob_, sb = run(False)
    assert sum(int(o[2].sum()) for o in ob_) > 0, "the run must cross auto-resets"
    for t, (x, y) in enumerate(zip(oa, ob_)):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) and torch.equal(x[2], y[2]), (planner, "host buffers", t)
    for k, (x, y) in enumerate(zip(sa, sb)):
        for name in names:
            assert torch.equal(x[name], y[name]), (name, k)
