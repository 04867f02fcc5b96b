"""SB3 VecEnv contract of So100VecEnv (shapes, dtypes, auto-reset, terminal_observation, TimeLimit.truncated,
episode infos), on CPU through a test-only oracle backend; the same adapter drives the CUDA backend on the GPU."""
import numpy as np
import pytest

from oracle_backend import OracleBackend
from so100_mujoco_rl_b200.vec_env import So100VecEnv


def make(task, n, limit):
    return So100VecEnv(f"Env0{task}", n, backend=OracleBackend(task, n, seed=1, max_episode_steps=limit),
                       max_episode_steps=limit)


@pytest.mark.parametrize("task,od", [(1, 15), (2, 15), (5, 8), (6, 15)])
def test_spaces_and_shapes(task, od, spec):
    env = make(task, 5, 10)
    assert env.num_envs == 5
    assert env.observation_space.shape == (od,) and env.action_space.shape == (6,)
    assert env.observation_space.dtype == np.float32 and env.action_space.dtype == np.float32
    assert np.allclose(env.action_space.low, -1) and np.allclose(env.action_space.high, 1)
    assert np.allclose(env.observation_space.low[:6], spec.jnt_range[:, 0]) and np.allclose(env.observation_space.high[:6], spec.jnt_range[:, 1])
    if task == 5:
        assert np.allclose(env.observation_space.low[6:], 0) and np.allclose(env.observation_space.high[6:], 5)   # env_base_02.py:64-68
    else:
        assert np.allclose(env.observation_space.low[6:9], -1) and np.allclose(env.observation_space.high[9:], 0.5)  # env_base_01.py:67-73
    obs = env.reset()
    assert obs.shape == (5, od) and obs.dtype == np.float32
    obs, rew, dones, infos = env.step(np.zeros((5, 6), np.float32))
    assert obs.shape == (5, od) and rew.shape == (5,) and rew.dtype == np.float32
    assert dones.shape == (5,) and dones.dtype == bool and len(infos) == 5 and all(isinstance(i, dict) for i in infos)
    env.close()


def test_autoreset_infos_follow_dummyvecenv():
    n, limit = 4, 6
    env = make(1, n, limit)
    env.reset()
    rets = np.zeros(n)
    for t in range(1, 2 * limit + 1):
        obs, rew, dones, infos = env.step(np.full((n, 6), 0.1, np.float32))
        rets += rew
        if t % limit == 0:
            assert dones.all()
            for i, info in enumerate(infos):
                assert info["TimeLimit.truncated"] is True
                assert info["terminal_observation"].shape == (15,) and info["terminal_observation"].dtype == np.float32
                assert info["episode"]["l"] == limit and abs(info["episode"]["r"] - rets[i]) < 1e-4
                # the returned obs is the first observation of the NEXT episode: zero kinematics (reference quirk)
                assert (obs[i, 6:] == 0).all() and not (info["terminal_observation"][6:] == 0).all()
            rets[:] = 0
        else:
            assert not dones.any() and all(info == {} for info in infos)
    env.close()


def test_terminated_is_not_reported_as_truncated():
    env = make(5, 2, 6000)
    env.reset()
    a = np.zeros((2, 6), np.float32); a[:, 0] = 1.0   # rotate the base: the cube leaves the image, Env05 terminates
    seen = False
    for _ in range(120):
        obs, rew, dones, infos = env.step(a)
        for i in np.flatnonzero(dones):
            assert infos[i]["TimeLimit.truncated"] is False
            assert (infos[i]["terminal_observation"][6:] == -5).all()   # scaled miss marker of the last obs
            assert (obs[i, 6:] == -1).all()                              # reset obs is un-scaled (-1, -1)
            seen = True
        if seen:
            break
    assert seen
    env.close()


def test_actions_are_clipped_and_validated():
    env = make(1, 3, 50)
    env.reset()
    big = np.full((3, 6), 7.0, np.float32)
    o1 = env.step(big)[0]
    env2 = make(1, 3, 50)
    env2.reset()
    o2 = env2.step(np.ones((3, 6), np.float32))[0]
    assert np.array_equal(o1, o2)
    with pytest.raises(ValueError):
        env.step(np.zeros((2, 6), np.float32))
    with pytest.raises(RuntimeError):
        env.step_wait()


def test_rest_of_the_contract():
    env = make(2, 3, 50)
    assert env.get_attr("render_mode") == [None, None, None]
    assert env.get_attr("num_envs", indices=[0]) == [3]
    assert env.env_is_wrapped(object) == [False] * 3
    assert env.seed(5) == [5, 6, 7]
    env.set_options({"x": 1})
    assert env.get_images() == [None] * 3 and env.render() is None
    env.set_attr("foo", 3)
    assert env.get_attr("foo") == [3, 3, 3]
    with pytest.raises(NotImplementedError):
        env.env_method("anything")
    assert env.metadata["render_fps"] == 31
    env.close()


def test_unknown_env_ids_are_rejected():
    with pytest.raises(ValueError):
        So100VecEnv("Env03", 2, backend=object())
    with pytest.raises(ValueError):
        So100VecEnv("Env04", 2, backend=object())
    from so100_mujoco_rl_b200.tasks import task_id
    assert task_id("Env01-v1") == 1 and task_id("Env05") == 5


def test_sb3_can_drive_it_if_installed():
    sb3 = pytest.importorskip("stable_baselines3")
    env = make(1, 4, 50)
    model = sb3.PPO("MlpPolicy", env, n_steps=16, batch_size=32, device="cpu")
    model.learn(total_timesteps=128)


def test_single_env_gymnasium_face_over_the_batched_backend():
    """So100Env: reset -> (obs, info); step -> 5-tuple; the step that ends an episode returns its LAST observation and
    the next reset() hands out the already-produced first observation of the next episode (gymnasium semantics)."""
    from oracle_backend import OracleBackend
    from so100_mujoco_rl_b200.gym_env import So100Env
    env = So100Env("Env01", backend=OracleBackend(1, 1, seed=4, max_episode_steps=3), max_episode_steps=3)
    with pytest.raises(RuntimeError):
        env.step(np.zeros(6, dtype=np.float32))
    obs, info = env.reset()
    assert obs.shape == (15,) and obs.dtype == np.float32 and info == {}
    assert env.observation_space.shape == (15,) and env.action_space.shape == (6,)
    seen = []
    for t in range(3):
        o, r, term, trunc, info = env.step(np.zeros(6, dtype=np.float32))
        seen.append(o)
        assert isinstance(r, float) and term is False
    assert trunc is True and "episode" in info and info["episode"]["l"] == 3
    assert np.allclose(env.get_block_pos() - env.get_end_effector_pos(), seen[-1][6:9], atol=1e-6)   # env_base_01.py:241-270
    assert abs(env.get_block_to_end_distance() - np.linalg.norm(seen[-1][6:9])) < 1e-7 and env.get_joint_angles().shape == (6,)
    assert np.any(seen[-1][6:] != 0)                      # last observation of the finished episode: real kinematics
    with pytest.raises(RuntimeError):
        env.step(np.zeros(6, dtype=np.float32))
    nxt, _ = env.reset()
    assert np.all(nxt[6:] == 0)                           # first observation of the next episode: zero kinematics (Q2)
    env.close()
