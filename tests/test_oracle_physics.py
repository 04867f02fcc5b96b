"""Pins the fp64 oracle (oracle/so100_oracle.c) with analytic invariants and the survey's hand-derived anchors
(SURVEY.md Appendix A.4).  The reference holds no golden vectors for mj_step (its arithmetic is the MuJoCo wheel),
so these are the strongest pins available offline: PARITY WITH REAL MUJOCO REMAINS UNPINNED."""
import copy

import numpy as np
import pytest

from conftest import make_oracle
from oracle.pyoracle import philox
from so100_mujoco_rl_b200.tasks import REST_POSITION, START_POSITION_05, VALID_START_POSITIONS


@pytest.fixture(scope="module")
def orc():
    return make_oracle(1, 1)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10: counter 0 / key 0
    assert [hex(x) for x in philox(0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


def test_survey_anchors(orc):
    m0, kv, iw = orc.derived()
    assert np.allclose(m0, [0.131490, 0.125088, 0.108872, 0.101160, 0.100043, 0.100028], atol=1e-6)
    assert np.allclose(kv, [5.12815, 5.00176, 4.66631, 4.49799, 4.47311, 4.47276], atol=1e-5)
    assert np.allclose(iw, [7.605165, 8.126401, 9.329947, 9.905365, 9.995670, 9.997296], atol=1e-6)
    k = orc.fk([0] * 6)
    assert np.allclose(k["end_pos"], [0, -0.4834, 0.0962], atol=1e-4) and np.allclose(k["wrist_pos"], [0, -0.3233, 0.0962], atol=1e-4)
    k = orc.fk(REST_POSITION)
    assert np.allclose(k["end_pos"], [0, -0.1877, 0.0229], atol=1e-4) and np.allclose(k["wrist_pos"], [0, -0.0980, 0.1555], atol=1e-4)
    k = orc.fk(START_POSITION_05)
    assert np.allclose(k["end_pos"], [0, -0.2625, 0.2400], atol=1e-4) and np.allclose(k["cam_xpos"], [-0.0015, -0.2368, 0.3320], atol=1e-4)
    assert np.allclose(orc.fk(VALID_START_POSITIONS[0])["end_pos"], [0.0180, -0.1999, 0.2650], atol=1e-4)
    M = orc.mass_matrix([0] * 6)
    assert np.isclose(M[1, 2], 0.014293, atol=1e-6) and np.isclose(M[1, 3], 0.004342, atol=1e-6) and np.isclose(M[2, 3], 0.002892, atol=1e-6)
    z = [0] * 6
    assert np.allclose(-orc.bias(z, z), [0, 0.948174, 0.474886, 0.126042, 0.000352, -0.006027], atol=1e-6)
    assert np.allclose(-orc.bias(REST_POSITION, z), [0, -0.055260, 0.417972, 0.069496, 0.000197, -0.003111], atol=1e-6)
    assert np.allclose(-orc.bias(START_POSITION_05, z), [0, 0.192522, 0.321925, 0.100069, -0.001261, -0.001538], atol=1e-6)


def test_env05_start_projection_anchor():
    o = make_oracle(5, 1)
    k = o.fk(START_POSITION_05)
    p = k["cam_xmat"].reshape(3, 3).T @ (np.array([0, -0.35, 0.01]) - k["cam_xpos"])
    assert np.allclose(p, [0.0002, -0.1077, -0.3239], atol=2e-4)
    f = 0.5 * 1920 / np.tan(np.deg2rad(120) / 2)
    u, v = f * p[0] / p[2] + 540, f * p[1] / p[2] + 960
    assert abs((1080 - int(u)) / 1080 - 0.5009) < 2e-3 and abs((1920 - int(v)) / 1920 - 0.4042) < 2e-3


def test_mass_matrix_symmetric_positive_definite(orc, spec):
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.uniform(spec.jnt_range[:, 0], spec.jnt_range[:, 1])
        M = orc.mass_matrix(q)
        assert np.abs(M - M.T).max() < 1e-16
        w = np.linalg.eigvalsh(M)
        assert w.min() > 0.09 and w.max() / w.min() < 1.5  # armature-dominated


def test_bias_equals_lagrangian_terms(orc, spec):
    """qfrc_bias = C(q,qd) qd + dV/dq, with C from the Christoffel symbols of M(q) (central differences)."""
    rng = np.random.default_rng(1)
    eps = 1e-6
    for _ in range(10):
        q = rng.uniform(spec.jnt_range[:, 0], spec.jnt_range[:, 1]); qd = rng.normal(0, 2, 6)
        dM = np.zeros((6, 6, 6)); dV = np.zeros(6)
        for k in range(6):
            e = np.zeros(6); e[k] = eps
            dM[:, :, k] = (orc.mass_matrix(q + e) - orc.mass_matrix(q - e)) / (2 * eps)
            dV[k] = (orc.energy(q + e, qd)[1] - orc.energy(q - e, qd)[1]) / (2 * eps)
        c = np.zeros(6)
        for i in range(6):
            for j in range(6):
                for k in range(6):
                    c[i] += 0.5 * (dM[i, j, k] + dM[i, k, j] - dM[j, k, i]) * qd[j] * qd[k]
        assert np.abs(orc.bias(q, qd) - (c + dV)).max() < 2e-8


def _variant(spec, **over):
    s = copy.deepcopy(spec)
    for k, v in over.items():
        setattr(s, k, np.asarray(v, dtype=float) if not np.isscalar(v) else v)
    return s


def test_energy_is_conserved_without_actuation_and_friction(spec):
    """kp = 0, kv = 0, no friction-loss: the arm is a frictionless compound pendulum; semi-implicit Euler keeps the
    total energy within O(h) of its initial value while it swings well away from the limits."""
    s = _variant(spec, act_kp=np.zeros(6), act_dampratio=np.zeros(6), act_kv=np.zeros(6), jnt_frictionloss=np.zeros(6),
                 jnt_armature=np.full(6, 0.1))
    o = make_oracle(1, 1, spec=s)
    q = np.array([0.3, -1.6, 1.5, 0.2, 0.1, 0.5]); v = np.zeros(6); w = np.zeros(6)
    T0, V0 = o.energy(q, v)
    drift = 0
    for _ in range(40):
        q, v, w = o.substeps(q, v, w, np.zeros(6), 5)
        T, V = o.energy(q, v)
        drift = max(drift, abs(T + V - T0 - V0))
        assert (q > s.jnt_range[:, 0]).all() and (q < s.jnt_range[:, 1]).all()
    assert T > 1e-4          # it actually moved
    assert drift < 2e-3 * abs(V0) + 2e-4


def test_friction_holds_small_torques(orc, spec):
    """|gravity + servo torque| < frictionloss on a joint at rest => (almost) no acceleration: the Huber row sticks."""
    q = np.array(REST_POSITION, dtype=float); v = np.zeros(6)
    g = -orc.bias(q, v)
    ctrl = q - g / 50.0            # servo cancels gravity exactly
    ctrl[0] += 0.05 / 50.0         # +0.05 N m on Rotation, below the 0.1 N m friction loss
    a = orc.forward(q, v, ctrl)[0]
    m00 = orc.mass_matrix(q)[0, 0]
    assert abs(a[0]) < 0.05 / m00 * 0.2    # far below the free response 0.05/M00
    ctrl[0] += 0.25 / 50.0         # 0.30 N m: slips, net 0.2 N m over M00
    a = orc.forward(q, v, ctrl)[0]
    assert 0.95 * 0.2 / m00 < a[0] < 1.05 * 0.2 / m00


def test_servo_settles_to_target_within_friction_deadband(orc):
    q = np.array(START_POSITION_05); v = np.zeros(6); w = np.zeros(6)
    ctrl = q + np.array([0.05, -0.05, 0.05, -0.05, 0.05, -0.05])
    q, v, w = orc.substeps(q, v, w, ctrl, 1500)
    g = -orc.bias(q, np.zeros(6))
    # steady state: |kp (ctrl - q) + gravity| <= frictionloss (+ a little slack for the soft constraint)
    assert np.abs(50 * (ctrl - q) + g).max() < 0.1 + 5e-3
    assert np.abs(v).max() < 1e-3


def test_joint_limit_pushes_back(orc, spec):
    q = np.array(REST_POSITION, dtype=float); q[2] = spec.jnt_range[2, 1] + 0.01  # Elbow 0.01 rad beyond its upper limit
    v = np.zeros(6)
    a_in = orc.forward(q, v, q.copy())[0]
    q2 = q.copy(); q2[2] = spec.jnt_range[2, 1] - 0.01
    a_free = orc.forward(q2, v, q2.copy())[0]
    assert a_in[2] < a_free[2] - 5.0  # K*imp*dist = 2770*0.95*0.01 ~ 26 rad/s^2 of reference acceleration
    q, v, w = orc.substeps(q, v, np.zeros(6), q.copy() + 0.0, 400)
    assert q[2] < spec.jnt_range[2, 1] + 0.004


def test_newton_reaches_the_unique_minimiser(orc, spec):
    """KKT check of the oracle's solve: M(a - a_smooth) = J^T f with f inside the Huber / unilateral laws."""
    rng = np.random.default_rng(4)
    m0, kv, iw = orc.derived()
    for _ in range(100):
        q = rng.uniform(spec.jnt_range[:, 0] - 0.03, spec.jnt_range[:, 1] + 0.03); v = rng.normal(0, 1, 6)
        u = q + rng.uniform(-1, 1, 6) * 0.075
        a, a_s, fc, it = orc.forward(q, v, u)
        assert it < 30
        M = orc.mass_matrix(q)
        assert np.abs(M @ (a - a_s) - fc).max() < 1e-10 * (1 + np.abs(M @ a_s).max())
        inside = (q >= spec.jnt_range[:, 0]) & (q <= spec.jnt_range[:, 1])
        if len(orc.contacts(q)[1]) == 0:                     # (pad-floor contact rows add their own joint torques)
            assert (np.abs(fc[inside]) <= 0.1 + 1e-12).all()   # friction-loss bound where no limit row is active


def test_against_real_mujoco_if_available(orc):
    """Deferred physics pin: consumes tests/golden/mujoco_arm.npz (tools/dump_mujoco_golden.py) when someone has
    produced it on a machine with MuJoCo; skipped otherwise (the build image cannot run MuJoCo)."""
    import os
    from conftest import ROOT
    path = os.path.join(ROOT, "tests", "golden", "mujoco_arm.npz")
    if not os.path.exists(path):
        pytest.skip("no real-MuJoCo dump available (parity with MuJoCo itself stays unpinned)")
    g = np.load(path)
    m0, kv, iw = orc.derived()
    assert np.allclose(m0, g["dof_M0"], rtol=1e-9) and np.allclose(kv, g["kv"], rtol=1e-9) and np.allclose(iw, g["dof_invweight0"], rtol=1e-9)
    per = int(g["steps"]) * 16 + 1
    for e in range(int(g["episodes"])):
        q, v, w = g["qpos"][e * per].copy(), g["qvel"][e * per].copy(), np.zeros(6)
        for k in range(per - 1):
            q, v, w = orc.substeps(q, v, w, g["ctrl"][e * per + k], 1)
            assert np.abs(q - g["qpos"][e * per + k + 1]).max() < 1e-7
        if "block_z" in g:  # the block-floor contact restatement against MuJoCo's own 6-dof box
            z, vz = float(g["block_z"][e * per]), float(g["block_vz"][e * per])
            for k in range(per - 1):
                z, vz = orc.block_substeps(z, vz, 1)
                assert abs(z - g["block_z"][e * per + k + 1]) < 1e-7


# ------------------------------------------------------------------------------------------ block <-> floor contact
def _block_constants(spec):
    """(K, B, lam(imp)) of the 16 identical pyramid rows; see oracle/so100_oracle.c:block_accel."""
    d0, d1, w, mid, p = spec.contact_solimp
    tc, dr = spec.contact_solref
    K, B = 1 / (d1 * d1 * tc * tc * dr * dr), 2 / (d1 * tc)

    def imp(dist):
        x = min(abs(dist) / w, 1.0)
        y = x ** p / mid ** (p - 1) if x <= mid else 1 - (1 - x) ** p / (1 - mid) ** (p - 1)
        return d0 + y * (d1 - d0)

    mu = spec.block_friction
    lam = lambda i: 4 * spec.block_ncon / (2 * mu * mu * (1 + mu * mu)) * i / (1 - i)  # noqa: E731
    return K, B, imp, lam


def test_block_pops_out_of_the_floor_and_rests_at_the_force_balance(spec):
    """Spawned with its centre on the plane (env01_v1.py:51-52) the box is pushed up without overshoot and comes to
    rest where gravity equals the 16 soft rows: g + lam(imp) * (-K imp dist) = 0."""
    o = make_oracle(1, 1)
    z, v, zs = 0.0, 0.0, []
    for _ in range(2000):
        z, v = o.block_substeps(z, v, 1)
        zs.append(z)
    zs = np.array(zs)
    assert (np.diff(zs) > -1e-15).all() and zs.max() < spec.block_half_z        # critically damped: monotone, never leaves
    assert 0.009 < zs[3 * 16 - 1] < 0.0099                                       # ~3 env steps to get within 1 mm
    K, B, imp, lam = _block_constants(spec)
    dist = z - spec.block_half_z
    i = imp(dist)
    assert abs(spec.gravity[2] + lam(i) * (-K * i * dist)) < 1e-6 and abs(v) < 1e-9
    assert abs(dist + 1.078e-4) < 2e-6                                           # rests 0.108 mm inside the plane


def test_block_above_the_floor_is_in_free_fall_and_contact_rows_only_push(spec):
    o = make_oracle(1, 1)
    h, g = spec.timestep, spec.gravity[2]
    z, v = o.block_substeps(0.05, 0.0, 10)
    assert np.isclose(v, 10 * h * g) and np.isclose(z, 0.05 + h * h * g * 55)    # semi-implicit Euler: sum_{k<=10} k = 55
    # moving up fast enough the reference acceleration drops below gravity: the rows carry no force (unilateral)
    K, B, imp, lam = _block_constants(spec)
    z0, v0 = spec.block_half_z - 1e-5, 1.0
    assert -B * v0 - K * imp(-1e-5) * (-1e-5) < g
    z1, v1 = o.block_substeps(z0, v0, 1)
    assert np.isclose(v1, v0 + h * g)
    # an applied force that cancels gravity (env03_v1.py:118-122) leaves a block at rest on the surface untouched
    z2, v2 = o.block_substeps(spec.block_half_z, 0.0, 16, fz_applied=-spec.block_mass * g)
    assert z2 == spec.block_half_z and v2 == 0.0


def test_block_one_substep_matches_the_closed_form(spec):
    o = make_oracle(1, 1)
    K, B, imp, lam = _block_constants(spec)
    rng = np.random.default_rng(0)
    for _ in range(200):
        z0, v0 = rng.uniform(0.0, 0.0102), rng.uniform(-0.3, 0.3)
        dist, g = z0 - spec.block_half_z, spec.gravity[2]
        a = g
        if dist <= 0:
            i = imp(dist)
            aref = -B * v0 - K * i * dist
            if g < aref:
                a = (g + lam(i) * aref) / (1 + lam(i))
        z1, v1 = o.block_substeps(z0, v0, 1)
        assert np.isclose(v1, v0 + spec.timestep * a, rtol=1e-12, atol=1e-15)
        assert np.isclose(z1, z0 + spec.timestep * v1, rtol=1e-12, atol=1e-15)


def test_static_block_flag_restores_the_held_block(spec):
    from so100_mujoco_rl_b200.tasks import make_task_cfg  # noqa: F401
    o = make_oracle(2, 3, seed=4, flags=8)
    o.reset()
    o.step(np.zeros((3, 6), dtype=np.float32))
    assert (o.gather("block")[:, 2] == 0.0).all() and (o.gather("block_vz") == 0.0).all()


# ---------------------------------------------------------------------------------------------------------------
# arm <-> floor contact: the jaws' primitive pad colliders (reference so_arm100_camera.xml:60-61, :108-111, :120-123)
def _pad_corner_heights(o, spec, q):
    """World z of all 64 pad-box corners at configuration q (from the oracle's own kinematics)."""
    k = o.fk(q)
    zs = []
    for b, pos, size in zip(spec.pad_body, spec.pad_pos, spec.pad_size):
        R, p = k["xmat"][b].reshape(3, 3), k["xpos"][b]
        for i in range(8):
            v = np.array([size[0] if i & 1 else -size[0], size[1] if i & 2 else -size[1], size[2] if i & 4 else -size[2]])
            zs.append((p + R @ (pos + v))[2])
    return np.array(zs)


def test_pad_model_and_body_invweight(spec):
    o = make_oracle(1, 1, flags=16)
    assert len(spec.pad_body) == 8 and spec.pad_body == [4, 4, 4, 4, 5, 5, 5, 5]
    assert np.allclose(spec.pad_solimp, [2, 1, 0.01, 0.5, 2]) and np.allclose(spec.pad_solref, [0.01, 1]) and spec.pad_friction == 1.0
    w = o.body_invweight0()
    # translational inverse inertia grows along the chain (longer lever arms over 0.1 of armature) and is ~1/kg at the jaws
    assert (np.diff(w) > 0).all() and 0.5 < w[4] < 1.0 and 0.5 < w[5] < 1.2


def test_contacts_are_the_pad_corners_below_the_floor(spec):
    o = make_oracle(1, 1, flags=16)
    rng = np.random.default_rng(5)
    seen = 0
    for _ in range(300):
        q = rng.uniform(spec.jnt_range[:, 0], spec.jnt_range[:, 1])
        pos, dist, body = o.contacts(q)
        z = _pad_corner_heights(o, spec, q)
        below = np.sort(z[z < 0])
        # every contact is a penetrating corner (dist = its height, pos half way up to the plane), at most 4 per pad
        assert len(dist) <= len(below) and len(dist) <= 32
        assert all(np.abs(below - d).min() < 1e-12 for d in dist)
        assert np.allclose(pos[:, 2], dist / 2) if len(dist) else True
        assert set(body) <= {4, 5}
        seen += len(dist) > 0
    assert seen > 20
    assert len(make_oracle(1, 1).contacts(q)[1]) == 0      # without SO100_FLAG_ARM_CONTACT the pads are ignored


def test_arm_pushed_into_the_floor_rests_on_its_pads(spec):
    """Servo targets 0.3 rad "below the floor": the arm comes to rest on the pads, micrometres deep (impedance 0.9999),
    the contact forces only push (unilateral rows) and balance servo + gravity, and the end-effector point stays above
    the plane - where without the pads it ends up centimetres below."""
    o = make_oracle(1, 1, flags=16)
    q = np.array([0.0, -1.0, 1.1, 1.3, 0.0, 0.3])            # jaw tip 2.5 cm above the floor
    assert _pad_corner_heights(o, spec, q).min() > 0.005
    ctrl = q.copy(); ctrl[1] += 0.3                        # pitch the arm down through the floor
    qq, v, w = o.substeps(q, np.zeros(6), np.zeros(6), ctrl, 3000)
    assert np.abs(v).max() < 2e-3                          # at rest up to a slow creep along the floor
    zmin = _pad_corner_heights(o, spec, qq).min()
    assert -2e-6 < zmin < 0                                # resting penetration ~1e-7 m
    a, a_s, fc, it = o.forward(qq, v, ctrl, warm=w)
    assert np.abs(a).max() < 0.1
    M = o.mass_matrix(qq)
    assert np.abs(M @ (a - a_s) - fc).max() < 1e-9        # constraint forces = what the solve added to the smooth dynamics
    assert np.abs(fc).max() > 0.5                          # and they are substantial (N m): the floor carries the servo's push
    free = make_oracle(1, 1)
    qf, vf, _ = free.substeps(q, np.zeros(6), np.zeros(6), ctrl, 1500)
    assert free.fk(qf)["end_pos"][2] < -0.01 < 0 < o.fk(qq)["end_pos"][2]


def test_random_actions_no_longer_go_through_the_floor():
    """without arm-floor contact the end-effector point is below z = 0 in 20-36 % of random-action steps;
    with the pads it never is."""
    n, steps = 128, 250
    frac = {}
    for flags in (0, 16):
        o = make_oracle(1, n, seed=1, flags=flags)
        o.reset()
        rng = np.random.default_rng(0)
        below = tot = 0
        for t in range(steps):
            obs, *_ = o.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32), nthreads=0)
            if t >= 5:
                below += int((obs[:, 14] < 0).sum()); tot += n
        frac[flags] = below / tot
    assert frac[0] > 0.1 and frac[16] == 0.0
