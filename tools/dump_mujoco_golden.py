#!/usr/bin/env python
"""Dumps REAL MuJoCo trajectories of the mesh-free so100 scene for the deferred physics pin.

Not runnable in the build image (no `mujoco` wheel, no network).  On any machine with `mujoco>=3.3.1`:

    python tools/dump_mujoco_golden.py --out tests/golden/mujoco_arm.npz

It loads so100_mujoco_rl_b200/assets/so100_scene.xml (mesh-free: the only contact pair left is block <-> floor),
replays seeded ctrl sequences with mj_step (the block spawned with its centre on the floor plane as env01_v1.py:51-52
does) and stores arm qpos/qvel/ctrl and the block's z / vz per substep plus the model constants MuJoCo
derived (dof_M0, dof_invweight0, actuator kv).  tests/test_oracle_physics.py::test_against_real_mujoco_if_available
compares the oracle with the file when it exists and is skipped otherwise.
"""
import argparse
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    import mujoco  # noqa: F401  (intentionally a hard requirement of this tool only)
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "mujoco_arm.npz"))
    ap.add_argument("--episodes", type=int, default=8)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    m = mujoco.MjModel.from_xml_path(os.path.join(ROOT, "so100_mujoco_rl_b200", "assets", "so100_scene.xml"))
    d = mujoco.MjData(m)
    rng = np.random.default_rng(0)
    lo, hi = m.jnt_range[:6, 0], m.jnt_range[:6, 1]
    Q, V, U, BZ, BV = [], [], [], [], []
    for _ in range(args.episodes):
        mujoco.mj_resetData(m, d)
        d.qpos[:6] = rng.uniform(lo + 0.1, hi - 0.1)
        d.qpos[6:9] = [0.0, -0.3, 0.0]
        for _ in range(args.steps):
            d.ctrl[:] = d.qpos[:6] + rng.uniform(-1, 1, 6) * 0.075
            for _ in range(16):
                Q.append(d.qpos[:6].copy()); V.append(d.qvel[:6].copy()); U.append(d.ctrl.copy())
                BZ.append(d.qpos[8]); BV.append(d.qvel[8])
                mujoco.mj_step(m, d)
        Q.append(d.qpos[:6].copy()); V.append(d.qvel[:6].copy()); U.append(d.ctrl.copy())
        BZ.append(d.qpos[8]); BV.append(d.qvel[8])
    np.savez_compressed(args.out, qpos=np.array(Q), qvel=np.array(V), ctrl=np.array(U), block_z=np.array(BZ), block_vz=np.array(BV), episodes=args.episodes,
                        steps=args.steps, dof_M0=m.dof_M0[:6], dof_invweight0=m.dof_invweight0[:6],
                        kv=-m.actuator_biasprm[:6, 2], mujoco_version=mujoco.__version__)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
