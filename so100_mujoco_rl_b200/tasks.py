"""Task constants of Env01 / Env02 / Env05 -> `so100_task_cfg` (include/so100_b200.h).

Every literal below is a constant of the reference; the file:line it comes from is given next to it
(paths relative to /root/reference/src/so100_mujoco_rl/).
"""
from __future__ import annotations

import ctypes
import math

NJ = 6
MAX_START = 64

TASK_ENV01, TASK_ENV02, TASK_ENV05, TASK_ENV06 = 1, 2, 5, 6
FLAG_FRESH_FK_ON_RESET = 1
FLAG_CLIP_ACTIONS = 2
FLAG_GENERIC_KERNEL = 4
FLAG_STATIC_BLOCK = 8  # hold the Env01/02/06 block at its spawn pose (no gravity, no floor contact)
FLAG_ARM_CONTACT = 16  # jaw-pad <-> floor contact (exact, opt-in: ~18x the step time; see include/so100_b200.h)

JOINT_STEP_SCALE = 0.075  # envs/utils.py:9
REST_POSITION = [0.0, -3.141, 3.117, 1.0, 0.0, 0.0]  # envs/utils.py:11
START_POSITION_05 = [0.0, -2.04, 1.19, 1.5, -1.58, 0.5]  # envs/env03_v1.py:10

# envs/utils.py:13-50 (36 recorded poses of the real arm; duplicates are in the reference too)
VALID_START_POSITIONS = [
    [0.116, -2.848, 1.84, 1.198, -1.598, 0.191],
    [0.11504855751991272, -3.0602917671203613, 2.4727771282196045, -0.5859806537628174, -1.5968739986419678, 0.18762288987636566],
    [0.11504855751991272, -3.063359498977661, 2.474310874938965, -0.5844466686248779, -1.5968739986419678, 0.18762288987636566],
    [0.11658254265785217, -3.049553871154785, 2.420621633529663, 0.09817477315664291, -1.5846021175384521, 0.19053177535533905],
    [0.11658254265785217, -3.049553871154785, 2.420621633529663, 0.11198059469461441, -1.5846021175384521, 0.19053177535533905],
    [0.7209709882736206, -2.597029447555542, 1.8867963552474976, 0.21629129350185394, -1.5968739986419678, 0.19053177535533905],
    [0.731708824634552, -2.607767343521118, 1.9911071062088013, 1.1780972480773926, -1.5968739986419678, 0.18471401929855347],
    [0.731708824634552, -2.607767343521118, 1.9911071062088013, 1.1780972480773926, -1.5968739986419678, 0.18471401929855347],
    [0.7225049734115601, -2.437495470046997, 0.6519418358802795, -0.8682331442832947, -1.59073805809021, 0.18471401929855347],
    [0.6151263117790222, -2.7719032764434814, 0.029145635664463043, -0.8682331442832947, -1.59073805809021, 0.18471401929855347],
    [0.6151263117790222, -2.7719032764434814, 0.029145635664463043, -0.8682331442832947, -1.59073805809021, 0.18471401929855347],
    [0.03374757617712021, -2.932971239089966, 0.03067961521446705, 0.5905826091766357, -2.4190876483917236, 0.18907733261585236],
    [0.11044661700725555, -2.787243127822876, 1.718058466911316, -0.9295923709869385, -2.4221556186676025, 0.19053177535533905],
    [0.11044661700725555, -2.787243127822876, 1.718058466911316, -0.9295923709869385, -2.4221556186676025, 0.19053177535533905],
    [0.1702718734741211, -1.8116313219070435, 2.230407953262329, -0.22549517452716827, -2.161378860473633, 0.19053177535533905],
    [0.6902913451194763, -1.7978254556655884, 2.2319419384002686, -0.22549517452716827, -2.1629128456115723, 0.19053177535533905],
    [0.6902913451194763, -1.7978254556655884, 2.2319419384002686, -0.22549517452716827, -2.1629128456115723, 0.19053177535533905],
    [1.1903691291809082, -1.7057865858078003, 2.1629128456115723, 0.8605632185935974, -1.7241944074630737, 0.18616846203804016],
    [0.007669903803616762, -2.7488934993743896, 2.8808159828186035, 0.5445631742477417, -1.7257283926010132, 0.19198621809482574],
    [0.007669903803616762, -2.7488934993743896, 2.8808159828186035, 0.5445631742477417, -1.7257283926010132, 0.19198621809482574],
    [-0.04908738657832146, -3.0173401832580566, 2.702874183654785, -0.06442718952894211, -1.7257283926010132, 0.19198621809482574],
    [-0.07516505569219589, -2.7274177074432373, 0.5246214270591736, -1.3406991958618164, -1.7211264371871948, 0.19198621809482574],
    [-0.07516505569219589, -2.7274177074432373, 0.5246214270591736, -1.3406991958618164, -1.7211264371871948, 0.19198621809482574],
    [-0.06902913749217987, -2.730485677719116, 0.5077476501464844, -1.3284273147583008, -1.7226604223251343, 0.19198621809482574],
    [1.0154953002929688, -3.1293208599090576, 0.5046796798706055, -1.3406991958618164, -1.7195924520492554, 0.19198621809482574],
    [1.0154953002929688, -3.1293208599090576, 0.5046796798706055, -1.3406991958618164, -1.7195924520492554, 0.19198621809482574],
    [1.371378779411316, -2.471243143081665, 2.633845090866089, 0.5921165943145752, -1.7211264371871948, 0.19198621809482574],
    [2.0202527046203613, -1.023165225982666, 1.3176895380020142, 0.5905826091766357, -1.7211264371871948, 0.19198621809482574],
    [2.0202527046203613, -1.023165225982666, 1.3176895380020142, 0.5905826091766357, -1.7211264371871948, 0.19198621809482574],
    [0.5967185497283936, -2.178252696990967, 1.7165244817733765, 0.5905826091766357, -1.7211264371871948, 0.19198621809482574],
    [0.200951486825943, -2.5003886222839355, 0.9234564304351807, -1.339165210723877, -1.7195924520492554, 0.19198621809482574],
    [0.200951486825943, -2.5003886222839355, 0.9234564304351807, -1.339165210723877, -1.7195924520492554, 0.19198621809482574],
    [0.777728259563446, -2.842466354370117, 0.0, -1.3514370918273926, -1.718058466911316, 0.19198621809482574],
    [-0.5077476501464844, -2.7765052318573, 0.00920388475060463, -1.0860583782196045, -1.718058466911316, 0.19198621809482574],
    [-0.5077476501464844, -2.7765052318573, 0.00920388475060463, -1.0860583782196045, -1.718058466911316, 0.19198621809482574],
    [-0.5077476501464844, -2.764233350753784, 1.2394564151763916, 1.1520196199417114, -1.7211264371871948, 0.19198621809482574],
]

# Env05 block box: min xyz / max xyz, envs/env05_v1.py:13-20
BLOCK_SPACE_START_05 = [[-0.05, -0.4, 0.01], [0.05, -0.3, 0.01]]
BLOCK_SPACE_END_05 = [[-0.45, -0.45, 0.01], [0.45, -0.25, 0.5]]

# gymnasium registration, __init__.py:5-45
MAX_EPISODE_STEPS = {TASK_ENV01: 4000, TASK_ENV02: 6000, TASK_ENV05: 6000, TASK_ENV06: 6000}
REWARD_THRESHOLD = {TASK_ENV01: 6000, TASK_ENV02: 8000, TASK_ENV05: 8000, TASK_ENV06: 8000}
OBS_DIM = {TASK_ENV01: 15, TASK_ENV02: 15, TASK_ENV05: 8, TASK_ENV06: 15}
ENV_IDS = {"Env01": TASK_ENV01, "Env01-v1": TASK_ENV01, "Env02": TASK_ENV02, "Env02-v1": TASK_ENV02,
           "Env05": TASK_ENV05, "Env05-v1": TASK_ENV05, "Env06": TASK_ENV06, "Env06-v1": TASK_ENV06}


class So100TaskCfg(ctypes.Structure):
    """ctypes mirror of `so100_task_cfg` (include/so100_b200.h); the test oracle declares the same layout."""
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("task", ctypes.c_int32),
        ("num_envs", ctypes.c_int32),
        ("max_episode_steps", ctypes.c_int32),
        ("env_offset", ctypes.c_int64),
        ("seed", ctypes.c_uint64),
        ("flags", ctypes.c_uint32),
        ("n_start", ctypes.c_int32),
        ("joint_step_scale", ctypes.c_double),
        ("start_positions", (ctypes.c_double * NJ) * MAX_START),
        ("rest_position", ctypes.c_double * NJ),
        ("start_position05", ctypes.c_double * NJ),
        ("block_dist_range", ctypes.c_double * 2),
        ("block_theta_half", ctypes.c_double),
        ("reach_threshold", ctypes.c_double),
        ("block_space_start", (ctypes.c_double * 3) * 2),
        ("block_space_end", (ctypes.c_double * 3) * 2),
        ("block_speed_min", ctypes.c_double),
        ("block_speed_max", ctypes.c_double),
        ("ramp_seconds", ctypes.c_double),
        ("cam_res_w", ctypes.c_double),
        ("cam_res_h", ctypes.c_double),
        ("obs_noise", ctypes.c_double),
        ("lost_limit", ctypes.c_int32),
        ("_pad0", ctypes.c_int32),
    ]


def task_id(env: str | int) -> int:
    """Accepts 1/2/5, "Env01", "Env01-v1", ... (the ids of __init__.py:5-45)."""
    if isinstance(env, int):
        if env not in OBS_DIM:
            raise ValueError(f"unsupported task {env}; this build covers Env01, Env02, Env05, Env06")
        return env
    if env not in ENV_IDS:
        raise ValueError(f"unsupported environment id {env!r}; this build covers Env01, Env02, Env05, Env06")
    return ENV_IDS[env]


def make_task_cfg(task: str | int, num_envs: int, seed: int = 0, env_offset: int = 0, flags: int = 0,
                  max_episode_steps: int | None = None) -> So100TaskCfg:
    t = task_id(task)
    c = So100TaskCfg()
    c.struct_size = ctypes.sizeof(So100TaskCfg)
    c.task = t
    c.num_envs = int(num_envs)
    c.max_episode_steps = int(max_episode_steps if max_episode_steps is not None else MAX_EPISODE_STEPS[t])
    c.env_offset = int(env_offset)
    c.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c.flags = int(flags)
    c.n_start = len(VALID_START_POSITIONS)
    c.joint_step_scale = JOINT_STEP_SCALE
    for i, row in enumerate(VALID_START_POSITIONS):
        for j in range(NJ):
            c.start_positions[i][j] = row[j]
    for j in range(NJ):
        c.rest_position[j] = REST_POSITION[j]
        c.start_position05[j] = START_POSITION_05[j]
    # envs/env01_v1.py:45 (0.18, 0.42); envs/env02_v1.py:55 (0.22, 0.42)
    lo = 0.22 if t in (TASK_ENV02, TASK_ENV06) else 0.18  # env06_v1.py:56 as Env02
    c.block_dist_range[0], c.block_dist_range[1] = lo, 0.42
    c.block_theta_half = 0.25 * math.pi  # env01_v1.py:47
    c.reach_threshold = 0.03  # env02_v1.py:29
    for a in range(2):
        for k in range(3):
            c.block_space_start[a][k] = BLOCK_SPACE_START_05[a][k]
            c.block_space_end[a][k] = BLOCK_SPACE_END_05[a][k]
    c.block_speed_min, c.block_speed_max = 0.0, 2.0  # env03_v1.py:21-22
    c.ramp_seconds = 12.0  # env03_v1.py:126
    c.cam_res_w, c.cam_res_h = 1080.0, 1920.0  # env_base_02.py:22-23
    c.obs_noise = 0.05  # env05_v1.py:44-45
    c.lost_limit = 30  # env03_v1.py:155
    return c
