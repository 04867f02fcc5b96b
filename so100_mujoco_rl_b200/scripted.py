"""Scripted controllers for workloads that need task events to fire (BASELINE config 3: Env02 relocations).

`reach_actions` servoes the end-effector of the first k envs onto their block with damped Jacobian-transpose steps.
The kinematics come from the library's own debug entry point (`so100_forward_dynamics`), so no second model lives
here; the controller is tooling around the hot path, never inside a timed region.
"""
from __future__ import annotations

import torch


def end_effector_jacobian(env, q: torch.Tensor, eps: float = 1e-3):
    """q [6, k] -> (end_pos [3, k], J [3, 6, k]) by forward differences of the kernel's kinematics."""
    z = torch.zeros_like(q)
    base = env.forward_dynamics(q, z, q)[3][:3]
    J = torch.zeros((3, 6, q.shape[1]), device=q.device)
    for j in range(6):
        qp = q.clone()
        qp[j] += eps
        J[:, j] = (env.forward_dynamics(qp, z, qp)[3][:3] - base) / eps
    return base, J


def reach_actions(env, obs: torch.Tensor, k: int, gain: float = 400.0) -> torch.Tensor:
    """Actions [k, 6] in [-1, 1] that move the end-effector of envs 0..k-1 towards their block (Env01/02/06 obs layout:
    joint angles in obs[:, :6]; the block position is read from the simulator state)."""
    q = obs[:k, :6].T.contiguous()
    base, J = end_effector_jacobian(env, q)
    err = env.get_state()["block"][:3, :k] - base
    return torch.clamp(torch.einsum("cjk,ck->kj", J, err) * gain, -1.0, 1.0)
