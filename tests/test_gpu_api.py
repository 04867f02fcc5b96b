"""C-ABI behaviours beyond parity: state get/set round trip + replay, masked reset, non-finite state recovery,
argument checking, the host-buffer path on ragged sizes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(task, n, **kw):
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    return BatchedSo100Env(task, n, device=0, **kw)


@pytest.mark.parametrize("task", [1, 2, 5, 6])
def test_state_roundtrip_replays_bit_exactly(task):
    n = 200
    env = _env(task, n, seed=5)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    acts = [torch.rand((n, 6), device="cuda", generator=g) * 2 - 1 for _ in range(12)]
    for a in acts[:5]:
        env.step(a)
    snap, tick = env.get_state(), env.tick
    out1 = [(r.obs.clone(), r.reward.clone()) for r in (env.step(a) for a in acts[5:])]
    env.set_state(snap)
    env.tick = tick
    out2 = [(r.obs.clone(), r.reward.clone()) for r in (env.step(a) for a in acts[5:])]
    for (o1, r1), (o2, r2) in zip(out1, out2):
        assert torch.equal(o1, o2) and torch.equal(r1, r2)


def test_masked_reset_touches_only_selected_envs():
    n = 64
    env = _env(1, n, seed=1)
    env.reset()
    for _ in range(3):
        env.step(torch.zeros((n, 6), device="cuda"))
    before = env.get_state()
    obs_before = env.obs.clone()
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mask[::4] = 1
    obs = env.reset(mask)
    after = env.get_state()
    keep = mask == 0
    assert torch.equal(after["qpos"][:, keep], before["qpos"][:, keep]) and torch.equal(obs[keep], obs_before[keep])
    assert (after["counters"][0, mask == 1] == 0).all() and (after["qvel"][:, mask == 1] == 0).all()
    assert (obs[mask == 1, 6:] == 0).all()


def test_non_finite_state_is_reset_and_counted():
    n = 32
    env = _env(1, n, seed=2)
    env.reset()
    env.step(torch.zeros((n, 6), device="cuda"))
    st = env.get_state()
    st["qvel"][2, 7] = float("nan")
    st["qpos"][0, 9] = float("inf")
    env.set_state(st)
    r = env.step(torch.zeros((n, 6), device="cuda"))
    assert env.stats()["nan_resets"] == 2
    assert torch.isfinite(r.obs).all()
    # a blown-up env is TERMINATED, never truncated (learners bootstrap truncations with V(terminal_obs)), and what it
    # reports is finite: reward, terminal observation (the reset one), episode return
    assert r.terminated[7] == 1 and r.terminated[9] == 1 and int(r.terminated.sum()) == 2 and int(r.truncated.sum()) == 0
    assert torch.isfinite(r.terminal_obs[[7, 9]]).all() and torch.equal(r.terminal_obs[[7, 9]], r.obs[[7, 9]])
    assert torch.isfinite(r.reward).all() and torch.isfinite(r.ep_return[[7, 9]]).all()
    assert (env.get_state()["counters"][0, [7, 9]] == 0).all()


@pytest.mark.parametrize("task", [1, 5])
def test_non_finite_env_does_not_poison_the_fused_learner(task):
    from so100_mujoco_rl_b200.ppo import FusedPPO, PPOConfig
    env = _env(task, 256, seed=4)
    algo = FusedPPO(env, PPOConfig(n_steps=4, n_minibatches=2, n_epochs=1, seed=1))
    algo.collect()
    st = env.get_state()
    st["qvel"][1, 3] = float("nan")
    env.set_state(st)
    adv, ret = algo.collect()
    assert env.stats()["nan_resets"] >= 1
    assert torch.isfinite(adv).all() and torch.isfinite(ret).all() and torch.isfinite(algo.buf["obs"]).all()
    algo.update(adv, ret)
    assert torch.isfinite(algo.params).all()


def test_argument_errors_surface_as_exceptions():
    from so100_mujoco_rl_b200._native import So100Error
    env = _env(2, 8)
    with pytest.raises(ValueError):
        env.step(torch.zeros((7, 6), device="cuda"))
    with pytest.raises(ValueError):
        env.reset(torch.zeros(3, dtype=torch.uint8, device="cuda"))
    with pytest.raises(So100Error):
        env.tick = -1
    with pytest.raises(ValueError):
        _env("Env03", 8)
    with pytest.raises(So100Error):
        _env(1, 0)


@pytest.mark.parametrize("n", [1, 255, 257, 16385 + 77])
def test_host_path_on_ragged_sizes(n):
    """so100_step_host pipelines chunks of envs; every size must give what the device path gives."""
    e1, e2 = _env(5, n, seed=4), _env(5, n, seed=4)
    host = e2.alloc_host()
    assert torch.equal(e1.reset().cpu(), e2.reset_host(host))
    rng = np.random.default_rng(0)
    for _ in range(4):
        a = torch.from_numpy(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
        r = e1.step(a.cuda())
        host["actions"].copy_(a)
        e2.step_host(host)
        assert torch.equal(r.obs.cpu(), host["obs"]) and torch.equal(r.reward.cpu(), host["reward"])
        assert torch.equal(r.terminated.cpu(), host["terminated"]) and torch.equal(r.truncated.cpu(), host["truncated"])


def test_host_path_delivers_terminal_rows_when_an_episode_ends():
    n = 300
    env = _env(1, n, seed=6, max_episode_steps=3)
    host = env.alloc_host()
    env.reset_host(host)
    for t in range(1, 7):
        host["actions"].zero_()
        env.step_host(host)
        if t % 3 == 0:
            assert host["truncated"].all() and (host["ep_len"] == 3).all()
            assert not (host["terminal_obs"][:, 6:] == 0).all() and (host["obs"][:, 6:] == 0).all()


def test_vec_env_adapter_on_the_gpu_backend():
    from so100_mujoco_rl_b200 import So100VecEnv
    env = So100VecEnv("Env02", 128, device=0, seed=3, max_episode_steps=5)
    obs = env.reset()
    assert obs.shape == (128, 15) and obs.dtype == np.float32
    for t in range(1, 11):
        obs, rew, dones, infos = env.step(env.action_space.sample()[None].repeat(128, 0))
        assert dones.all() == (t % 5 == 0)
        if dones.all():
            assert all(i["TimeLimit.truncated"] and i["episode"]["l"] == 5 for i in infos)
    env.close()


@pytest.mark.parametrize("n", [300, 16384 + 300])
def test_pageable_host_buffers_take_the_copy_pipeline_and_agree_with_pinned(n):
    """Pinned buffers: one zero-copy launch.  Pageable numpy arrays: chunked H2D -> kernel -> D2H.  Same results."""
    import ctypes
    from so100_mujoco_rl_b200 import _native
    e1, e2 = _env(2, n, seed=9, max_episode_steps=5), _env(2, n, seed=9, max_episode_steps=5)
    pinned = e1.alloc_host()
    e1.reset_host(pinned)
    od = e2.obs_dim
    buf = {"actions": np.zeros((n, 6), np.float32), "obs": np.zeros((n, od), np.float32), "reward": np.zeros(n, np.float32),
           "terminated": np.zeros(n, np.uint8), "truncated": np.zeros(n, np.uint8), "terminal_obs": np.zeros((n, od), np.float32),
           "ep_return": np.zeros(n, np.float32), "ep_len": np.zeros(n, np.int32)}
    p = lambda k: buf[k].ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    _native.check(e2._L.so100_reset_host(e2._h, p("obs"), None))
    assert np.array_equal(buf["obs"], pinned["obs"].numpy())
    rng = np.random.default_rng(2)
    for t in range(7):  # crosses the 5-step TimeLimit: terminal rows travel on that step
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        pinned["actions"].copy_(torch.from_numpy(a)); buf["actions"][...] = a
        e1.step_host(pinned)
        _native.check(e2._L.so100_step_host(e2._h, p("actions"), p("obs"), p("reward"), p("terminated"), p("truncated"),
                                            p("terminal_obs"), p("ep_return"), p("ep_len"), None))
        for k in ("obs", "reward", "terminated", "truncated"):
            assert np.array_equal(buf[k], pinned[k].numpy()), (t, k)
        if t == 4:
            assert buf["truncated"].all()
            for k in ("terminal_obs", "ep_return", "ep_len"):
                assert np.array_equal(buf[k], pinned[k].numpy()), k


def test_static_block_flag_and_block_contact_state():
    from so100_mujoco_rl_b200.tasks import FLAG_STATIC_BLOCK
    moving, held = _env(1, 64, seed=3), _env(1, 64, seed=3, flags=FLAG_STATIC_BLOCK)
    moving.reset(); held.reset()
    a = torch.zeros((64, 6), device="cuda")
    for _ in range(12):
        om, oh = moving.step(a).obs.clone(), held.step(a).obs.clone()
    bm, bh = moving.get_state()["block"].cpu().numpy(), held.get_state()["block"].cpu().numpy()
    assert (bh[2] == 0).all() and (bh[3] == 0).all()                     # held at the spawn height
    assert (np.abs(bm[2] - 0.009892) < 2e-6).all() and (np.abs(bm[3]) < 1e-5).all()   # resting 0.108 mm inside the plane
    assert torch.equal(om[:, :6], oh[:, :6])                              # the arm does not feel the block
    assert (np.abs((om - oh).cpu().numpy()[:, [8, 11]] - bm[2][:, None]) < 1e-6).all()  # obs z columns carry the lift


def test_reseeding_rekeys_the_device_rng():
    """gymnasium reset(seed=...) / SB3 VecEnv.seed: the same seed reproduces the same episode starts, another one does not."""
    a, b = _env(1, 256, seed=1), _env(1, 256, seed=2)
    oa, ob = a.reset().clone(), b.reset().clone()
    assert not torch.equal(oa, ob)
    b.seed(1)
    assert torch.equal(b.reset(), oa)
    act = torch.zeros((256, 6), device="cuda")
    for _ in range(3):
        ra, rb = a.step(act), b.step(act)
    assert torch.equal(ra.obs, rb.obs)
    from so100_mujoco_rl_b200 import So100VecEnv
    v = So100VecEnv("Env01", 64, device=0, seed=5)
    first = v.reset().copy()
    v.seed(9); second = v.reset().copy()
    v.seed(5); third = v.reset().copy()
    assert not np.array_equal(first, second) and np.array_equal(first, third)
    v.close()


@pytest.mark.parametrize("task,n,groups", [(1, 1000, 3), (5, 4096, 4), (2, 300, 2)])
def test_async_env_groups_equal_the_full_batch_step(task, n, groups):
    """so100_step_host_async over every group once == one so100_step_host call, bit for bit (each group counts its own
    steps; the RNG tick of a group's k-th step is k), also with short episodes so that auto-reset and the terminal rows
    are exercised, and with the groups finishing in a different order than they were issued."""
    from so100_mujoco_rl_b200 import _native
    e1, e2 = _env(task, n, seed=6, max_episode_steps=7), _env(task, n, seed=6, max_episode_steps=7)
    h1, h2 = e1.alloc_host(), e2.alloc_host()
    assert torch.equal(e1.reset_host(h1), e2.reset_host(h2))
    ranges = e2.host_groups(groups)
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    rng = np.random.default_rng(0)
    for t in range(20):
        a = torch.from_numpy(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
        h1["actions"].copy_(a); h2["actions"].copy_(a)
        e1.step_host(h1)
        order = list(range(groups)) if t % 2 == 0 else list(reversed(range(groups)))
        for g in order:
            e2.step_host_async(h2, g)
        if t % 3 == 0:   # served in completion order
            got = sorted(e2.step_host_wait_any() for _ in range(groups))
            assert got == list(range(groups)) and e2.step_host_wait_any() == -1
        else:
            for g in range(groups):
                e2.step_host_wait(g)
        for k in ("obs", "reward", "terminated", "truncated"):
            assert torch.equal(h1[k], h2[k]), (k, t)
        done = (h1["terminated"] | h1["truncated"]).bool()
        if done.any():
            for k in ("terminal_obs", "ep_return", "ep_len"):
                assert torch.equal(h1[k][done], h2[k][done]), (k, t)
    assert e1.tick == e2.tick == 20
    # groups out of step: the full-batch calls refuse until the lagging groups have caught up
    e2.step_host_async(h2, 0)
    with pytest.raises(_native.So100Error, match="in flight"):
        e2.step_host(h2)
    e2.step_host_wait(0)
    with pytest.raises(_native.So100Error, match="different step counts"):
        e2.step(torch.zeros((n, 6), device="cuda"))
    with pytest.raises(_native.So100Error, match="already has a step in flight"):
        e2.step_host_async(h2, 1); e2.step_host_async(h2, 1)
    for g in range(1, groups):
        e2.step_host_wait(g)
        if g > 1:
            e2.step_host_async(h2, g); e2.step_host_wait(g)
    e2.step_host(h2)   # all groups at 21 now
    assert e2.tick == 22
    with pytest.raises(_native.So100Error, match="page-locked"):
        pageable = {k: torch.zeros_like(v) for k, v in h2.items()}
        e2.step_host_async(pageable, 0)
