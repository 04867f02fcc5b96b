"""Reference behaviours ("quirks") exercised one by one on the oracle's task logic.
The same behaviours are pinned against the reference's own Python code by tests/test_golden_reference.py."""
import numpy as np
import pytest

from conftest import make_oracle
from so100_mujoco_rl_b200.tasks import REST_POSITION, START_POSITION_05, VALID_START_POSITIONS


def zeros(n):
    return np.zeros((n, 6), dtype=np.float32)


def test_q2_reset_obs_has_zero_kinematics_and_q5_jaw_untouched():
    o = make_oracle(1, 64, seed=3)
    obs = o.reset()
    assert (obs[:, 6:] == 0).all()
    assert (obs[:, 5] == 0).all()  # Jaw stays at qpos0 although the start rows carry a jaw value
    rows = np.array(VALID_START_POSITIONS, dtype=np.float32)
    for r in obs:
        assert (np.abs(rows[:, :5] - r[:5]).sum(axis=1) == 0).any()
    blk = o.gather("block")
    d = np.hypot(blk[:, 0], blk[:, 1])
    th = np.arctan2(blk[:, 1], blk[:, 0])
    assert (d >= 0.18).all() and (d <= 0.42).all() and (blk[:, 2] == 0).all()
    assert (th >= -0.75 * np.pi - 1e-12).all() and (th <= -0.25 * np.pi + 1e-12).all()


def test_q1_q4_first_reward_ignores_guarded_terms_then_sees_stale_zero_kinematics():
    o = make_oracle(1, 8, seed=1, max_episode_steps=3)
    o.reset()
    _, r1, *_ = o.step(zeros(8))
    # very first step of the env object: last_* are None -> only distance (0) and joint penalty terms
    q0 = o.gather("qpos")  # after the step, but the penalty was on the reset pose; recompute from the start rows
    assert (r1 <= 0).all()
    _, r2, *_ = o.step(zeros(8))
    _, r3, term, trunc, *_ = o.step(zeros(8))
    assert trunc.all() and not term.any()
    _, r4, *_ = o.step(zeros(8))
    # first reward after an auto-reset: kinematics are zero, guards are now open (Q4): end_z=0<0.02 -> -0.4,
    # wrist_z=0<0.08 -> clip(-0.8), distance 0 -> 0; plus the joint penalty of the new start pose (<= 0)
    assert (r4 <= -1.2 + 1e-12).all()
    assert (r1 > -1.2).any()  # ... which the very first step did not pay


def test_env02_first_step_after_reset_relocates_block_with_zero_bonus_then_bonus_later():
    o = make_oracle(2, 16, seed=5, max_episode_steps=4)
    obs0 = o.reset()
    assert np.allclose(obs0[:, :6], np.array(REST_POSITION, dtype=np.float32))
    b0 = o.gather("block").copy()
    _, r1, *_ = o.step(zeros(16))
    b1 = o.gather("block")
    assert (np.abs(b1 - b0).sum(axis=1) > 0).all()           # Q12: distance(0 kinematics) = 0 < 0.03 -> teleport
    d = np.hypot(b1[:, 0], b1[:, 1])
    assert (d >= 0.22).all() and (d <= 0.42).all()
    assert np.allclose(o.gather("last_block_pos"), b0)       # previous position remembered
    for _ in range(3):
        out = o.step(zeros(16))
    assert out[3].all()                                       # truncated at 4 steps, auto-reset
    b_prev = o.gather("last_block_pos").copy()
    _, r5, *_ = o.step(zeros(16))                             # first step of episode 2: relocation WITH bonus
    bonus = 20 * np.linalg.norm(o.gather("last_block_pos") - b_prev, axis=1)
    assert (bonus > 0).all()
    # reward = end_z term (-0.4) + wrist_z term (-0.8) + distance term (0) + joint penalty at REST_POSITION + bonus
    lo, hi = np.array([-2.2, -3.14158, 0, -2.0, -3.14158, -0.2]), np.array([2.2, 0.2, 3.14158, 1.8, 3.14158, 2.0])
    q = np.array(REST_POSITION)
    pen = -10 * (np.maximum(lo + 0.05 * (hi - lo) - q, 0) + np.maximum(q - (hi - 0.05 * (hi - lo)), 0)).sum()
    assert np.allclose(r5, -1.2 + pen + bonus, atol=1e-9)


def test_env05_reset_and_lost_cube_termination():
    o = make_oracle(5, 4, seed=2)
    obs0 = o.reset()
    assert np.allclose(obs0[:, :6], np.array(START_POSITION_05, dtype=np.float32)) and (obs0[:, 6:] == -1).all()
    assert np.allclose(o.gather("block"), [[0, -0.35, 0.01]] * 4)
    # rotate the base so that the cube leaves the image: misses accumulate, termination on the 32nd in a row
    a = zeros(4); a[:, 0] = 1.0
    miss_run = np.zeros(4, int); terminated_at = {}
    for t in range(200):
        obs, rew, term, trunc, tobs, epr, epl = o.step(a)
        for i in range(4):
            if term[i]:
                terminated_at.setdefault(i, (t, miss_run[i]))
                miss_run[i] = 0
            elif obs[i, 6] == -5 and obs[i, 7] == -5:
                miss_run[i] += 1
            else:
                miss_run[i] = 0
        if len(terminated_at) == 4:
            break
    assert len(terminated_at) == 4
    assert all(run == 31 for _, run in terminated_at.values())  # 31 earlier misses + the terminating 32nd


def test_env05_obs_lags_command_by_one_step_and_penalty_uses_commanded_angles():
    o = make_oracle(5, 2, seed=9)
    o.reset()
    a = zeros(2); a[:, 0] = 0.5
    obs1, *_ = o.step(a)
    assert np.allclose(obs1[:, 0], 0.0)                      # Q7: previous command (START_POSITION[0] = 0)
    obs2, *_ = o.step(a)
    assert np.allclose(obs2[:, 0], 0.5 * 0.075, atol=1e-7)
    assert np.allclose(o.gather("cmd")[:, 0], 2 * 0.5 * 0.075)


def test_env05_reward_decomposition():
    """reward = 0.5 - |centre - (0.5, 0.5)| + joint penalty(commanded, old) - 0.09375 * sum|a_t - a_{t-1}| * f  (Q7, Q8)."""
    o = make_oracle(5, 1, seed=4)
    o.reset()
    lo, hi = np.array([-2.2, -3.14158, 0, -2.0, -3.14158, -0.2]), np.array([2.2, 0.2, 3.14158, 1.8, 3.14158, 2.0])
    rng = np.random.default_rng(0)
    prev_a, checked = None, 0
    for t in range(60):
        a = rng.uniform(-0.3, 0.3, (1, 6)).astype(np.float32)
        cmd_old = np.array(o.state(0).cmd)
        obs, rew, term, *_ = o.step(a)
        s = o.state(0)
        f = min(t * 0.032 / 12.0, 1.0)
        r = 0.5
        if s.centre_valid:
            r -= np.hypot(0.5 - s.last_centre[0], 0.5 - s.last_centre[1])
        r += -10 * (np.maximum(lo + 0.05 * (hi - lo) - cmd_old, 0) + np.maximum(cmd_old - (hi - 0.05 * (hi - lo)), 0)).sum()
        if prev_a is not None:
            r -= 0.0025 * np.abs((a[0].astype(np.float64) - prev_a) * 0.075 / 0.002).sum() * f
            checked += 1
        assert abs(rew[0] - r) < 1e-9, t
        prev_a = a[0].astype(np.float64)
        if term[0]:
            break
    assert checked > 20


def test_timelimit_values_match_registration():
    from so100_mujoco_rl_b200.tasks import MAX_EPISODE_STEPS
    assert MAX_EPISODE_STEPS == {1: 4000, 2: 6000, 5: 6000, 6: 6000}  # __init__.py:8,15,36,43


def test_fresh_fk_flag_gives_real_kinematics_on_reset():
    from so100_mujoco_rl_b200.tasks import FLAG_FRESH_FK_ON_RESET
    o = make_oracle(1, 4, seed=1, flags=FLAG_FRESH_FK_ON_RESET)
    obs = o.reset()
    for i in range(4):
        k = o.fk(obs[i, :6].astype(np.float64))
        assert np.abs(obs[i, 12:15] - k["end_pos"]).max() < 1e-6
        assert np.abs(obs[i, 9:12] - o.gather("block")[i]).max() < 1e-7


def test_rng_is_keyed_by_global_env_id():
    full = make_oracle(1, 8, seed=7)
    hi = make_oracle(1, 4, seed=7, env_offset=4)
    assert np.array_equal(full.reset()[4:], hi.reset())


def test_env06_gripper_reward_and_no_relocation():
    """Env06: first step after a reset is "in reach" (zero kinematics): bonus + 100*sigmoid(10*(jaw_norm-0.3)); the block stays."""
    o = make_oracle(6, 4, seed=3)
    o.reset()
    b0 = o.gather("block").copy()
    _, r1, *_ = o.step(zeros(4))
    assert np.array_equal(o.gather("block")[:, :2], b0[:, :2])          # env06_v1.py:38: relocation is commented out
    assert (o.gather("block")[:, 2] > 0.004).all()                      # ... the floor contact is already lifting it
    jn = np.clip((0.0 + 0.2) / 2.2, 0, 1)                                # REST_POSITION jaw = 0
    grip = 100.0 / (1.0 + np.exp(-10 * (jn - 0.3)))
    lo, hi = np.array([-2.2, -3.14158, 0, -2.0, -3.14158, -0.2]), np.array([2.2, 0.2, 3.14158, 1.8, 3.14158, 2.0])
    q = np.array(REST_POSITION)
    pen = -10 * (np.maximum(lo + 0.05 * (hi - lo) - q, 0) + np.maximum(q - (hi - 0.05 * (hi - lo)), 0)).sum()
    assert np.allclose(r1, grip + pen, atol=1e-9)                       # very first step: guards closed, bonus 0
    _, r2, *_ = o.step(zeros(4))
    assert (r2 < 0).all()                                               # real kinematics now: not in reach any more


@pytest.mark.parametrize("task", [1, 2, 5])
def test_bulk_state_exchange_resumes_a_trajectory(task):
    """orc_set_state_soa (the product's so100_state_view layout) carries everything a run needs: an oracle that adopts
    another oracle's state (through float32, as it comes from the GPU) continues with the same flags and, to float32
    rounding of the hand-over, the same observations and rewards."""
    n = 24
    a, b = make_oracle(task, n, seed=3, max_episode_steps=40), make_oracle(task, n, seed=3, max_episode_steps=40)
    a.reset(); b.reset()
    rng = np.random.default_rng(1)
    for _ in range(30):
        a.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
    f32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64).T, dtype=np.float32)  # noqa: E731  [N, k] -> [k, N]
    snap, aux, cnt = np.zeros((12, n), np.float32), np.zeros((24, n), np.float32), np.zeros((4, n), np.int32)
    if task == 5:
        snap[:3], snap[3:12] = f32(a.gather("cam_xpos")), f32(a.gather("cam_xmat"))
        aux[:6], aux[6:12], aux[12:15] = f32(a.gather("cmd")), f32(a.gather("last_angvel")), f32(a.gather("target"))
        aux[15], aux[16:18] = a.gather("target_dt"), f32(a.gather("last_centre"))
        cnt[2], cnt[3] = a.gather("miss_count"), np.round(a.gather("target_time") / 0.032)
    else:
        snap[:3], snap[3], snap[4:7] = f32(a.gather("end_pos")), a.gather("wrist_pos")[:, 2], f32(a.gather("block_xpos"))
        aux[:3], aux[3:6] = f32(a.gather("task_block_pos")), f32(a.gather("last_block_pos"))
    cnt[0] = a.gather("elapsed_steps")
    cnt[1] = a.gather("ever_stepped") | (a.gather("has_last_block") << 1) | (a.gather("angvel_valid") << 2) | (a.gather("centre_valid") << 3)
    block = np.concatenate([f32(a.gather("block")), a.gather("block_vz")[None].astype(np.float32)])
    b.set_state_soa({"qpos": f32(a.gather("qpos")), "qvel": f32(a.gather("qvel")), "qacc_warm": f32(a.gather("qacc_warm")),
                     "block": block, "snap": snap, "aux": aux, "counters": cnt, "ep_return": a.gather("ep_return").astype(np.float32)})
    b.tick = a.tick
    qa, va, ba = a.get_state_soa(); qb, vb, bb = b.get_state_soa()
    assert np.abs(qa - qb).max() < 3e-7 and np.abs(va - vb).max() < 1e-5 and np.abs(ba - bb).max() < 1e-7
    for _ in range(15):   # crosses the 40-step TimeLimit
        act = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        oa, ra, ta, ca, *_ = a.step(act); ob, rb, tb, cb, *_ = b.step(act)
        assert (ta == tb).all() and (ca == cb).all()
        cols = slice(0, 6) if task == 5 else slice(None)
        assert np.abs(oa[:, cols] - ob[:, cols]).max() < 2e-5 and np.abs(ra - rb).max() < 5e-3
