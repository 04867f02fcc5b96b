"""Head-less software rasteriser of the so100 scene (no OpenGL, no MuJoCo): what `record` writes to video.

The reference records with MuJoCo's offscreen GL renderer through `VecVideoRecorder` (src/so100_mujoco_rl/main.py:127-171,
3000 steps, file name `rec-<env>-step-0-to-step-3000.mp4`) and shows the same scene in `test` (main.py:78-124).  Rendering
is off the hot path, so this is deliberately small: host-side forward kinematics from the `ModelSpec` (the same MJCF
numbers the simulator uses), a pin-hole camera placed like the scene's free camera (`<global azimuth="120"
elevation="-20"/>`, `<statistic center="0 0 0.1" extent="0.8"/>`, env01.xml:9-15), and painter's-algorithm drawing of
the floor grid, the links (thick segments between joint origins), the jaw pads and the block (boxes) with OpenCV's 2-D
primitives.  It is a schematic of the state, not a photograph of the meshes (the STL files are not in the checkout).
"""
from __future__ import annotations

import math

import numpy as np

from .model import ModelSpec, load_model, quat_to_mat


def forward_kinematics(spec: ModelSpec, qpos) -> tuple[np.ndarray, np.ndarray]:
    """World positions [6, 3] and rotations [6, 3, 3] of the six moving bodies at joint angles qpos (MuJoCo semantics:
    body frame = parent frame * body_pos/quat, then the hinge rotation about jnt_axis at the body origin)."""
    R = quat_to_mat(spec.base_quat)
    p = np.asarray(spec.base_pos, dtype=np.float64).copy()
    pos, rot = np.zeros((6, 3)), np.zeros((6, 3, 3))
    for i in range(6):
        p = p + R @ spec.body_pos[i]
        R = R @ quat_to_mat(spec.body_quat[i])
        ax, a = spec.jnt_axis[i] / np.linalg.norm(spec.jnt_axis[i]), float(qpos[i])
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = R @ (np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * (K @ K))   # Rodrigues
        pos[i], rot[i] = p, R
    return pos, rot


_BOX_EDGES = [(0, 1), (1, 3), (3, 2), (2, 0), (4, 5), (5, 7), (7, 6), (6, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
_BOX_FACES = [(0, 1, 3, 2), (4, 5, 7, 6), (0, 1, 5, 4), (2, 3, 7, 6), (0, 2, 6, 4), (1, 3, 7, 5)]


def _box_corners(centre, R, half):
    s = np.array([[(1 if i & 1 else -1), (1 if i & 2 else -1), (1 if i & 4 else -1)] for i in range(8)], dtype=np.float64)
    return centre + (s * half) @ R.T


class SceneRenderer:
    """`render(qpos, block_pos)` -> uint8 [height, width, 3] (RGB)."""

    LINK_COLOURS = [(230, 120, 20), (240, 140, 30), (250, 160, 40), (235, 130, 25), (60, 60, 60), (90, 90, 90)]

    def __init__(self, spec: ModelSpec | None = None, width: int = 480, height: int = 480, azimuth: float = 120.0,
                 elevation: float = -20.0, lookat=(0.0, -0.15, 0.1), distance: float = 1.0, fovy: float = 45.0):
        self.spec = spec or load_model()
        self.w, self.h = int(width), int(height)
        az, el = math.radians(azimuth), math.radians(elevation)
        fwd = np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])   # MuJoCo free camera: looks along this
        self.eye = np.asarray(lookat, dtype=np.float64) - distance * fwd
        right = np.cross(fwd, [0.0, 0.0, 1.0]); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        self.R = np.stack([right, up, fwd])        # world -> camera (x right, y up, z forward)
        self.f = 0.5 * self.h / math.tan(math.radians(fovy) / 2)

    def project(self, pts):
        pc = (np.atleast_2d(pts) - self.eye) @ self.R.T
        z = np.maximum(pc[:, 2], 1e-3)
        uv = np.stack([self.w / 2 + self.f * pc[:, 0] / z, self.h / 2 - self.f * pc[:, 1] / z], axis=1)
        return uv, pc[:, 2]

    def render(self, qpos, block_pos=None, text: str | None = None) -> np.ndarray:
        import cv2
        img = np.full((self.h, self.w, 3), (200, 205, 210), dtype=np.uint8)
        # floor: checker lines every 10 cm on z = 0
        for k in np.arange(-0.6, 0.61, 0.1):
            for a, b in (((k, -0.6, 0), (k, 0.6, 0)), ((-0.6, k, 0), (0.6, k, 0))):
                uv, z = self.project(np.array([a, b]))
                if (z > 0.01).all():
                    cv2.line(img, tuple(map(int, uv[0])), tuple(map(int, uv[1])), (150, 160, 170), 1, cv2.LINE_AA)
        pos, rot = forward_kinematics(self.spec, qpos)
        items = []   # (depth, draw function): far to near
        chain = np.vstack([np.asarray(self.spec.base_pos, dtype=np.float64), pos])
        for i in range(6):
            seg = chain[i:i + 2]
            uv, z = self.project(seg)
            thick = max(2, int(self.f * 0.022 / max(z.mean(), 0.05)))
            items.append((z.mean(), lambda uv=uv, c=self.LINK_COLOURS[i], t=thick: cv2.line(
                img, tuple(map(int, uv[0])), tuple(map(int, uv[1])), c, t, cv2.LINE_AA)))
        # jaw bodies: segment to the end-effector point / along the moving jaw, and the pad boxes
        ee = pos[self.spec.ee_body] + rot[self.spec.ee_body] @ self.spec.ee_offset
        for a, b, col in ((pos[4], ee, (40, 40, 40)), (pos[5], pos[5] + rot[5] @ np.array([0.0, -0.08, 0.0]), (70, 70, 70))):
            uv, z = self.project(np.array([a, b]))
            items.append((z.mean(), lambda uv=uv, c=col: cv2.line(img, tuple(map(int, uv[0])), tuple(map(int, uv[1])), c, 3, cv2.LINE_AA)))
        boxes = [(pos[b] + rot[b] @ p, rot[b], s, (30, 30, 160)) for b, p, s in zip(self.spec.pad_body, self.spec.pad_pos, self.spec.pad_size)]
        if block_pos is not None:
            boxes.append((np.asarray(block_pos, dtype=np.float64), np.eye(3), np.full(3, self.spec.block_half_z), (40, 200, 40)))
        for centre, R, half, col in boxes:
            corners = _box_corners(centre, R, np.asarray(half))
            uv, z = self.project(corners)

            def draw(uv=uv, col=col):
                for f in _BOX_FACES:
                    cv2.fillConvexPoly(img, np.round(uv[list(f)]).astype(np.int32), col, cv2.LINE_AA)
                for a_, b_ in _BOX_EDGES:
                    cv2.line(img, tuple(map(int, uv[a_])), tuple(map(int, uv[b_])), (20, 20, 20), 1, cv2.LINE_AA)
            items.append((z.mean(), draw))
        uv, z = self.project(ee[None])
        items.append((z[0] - 1e-3, lambda uv=uv: cv2.circle(img, tuple(map(int, uv[0])), 3, (220, 30, 30), -1, cv2.LINE_AA)))
        for _, fn in sorted(items, key=lambda it: -it[0]):
            fn()
        if text:
            cv2.putText(img, text, (8, 18), cv2.FONT_HERSHEY_SIMPLEX, 0.45, (20, 20, 20), 1, cv2.LINE_AA)
        return img


class VideoSink:
    """mp4 writer with VecVideoRecorder's naming (`<prefix>-step-<a>-to-step-<b>.mp4`), 31 fps like the env's render_fps."""

    def __init__(self, directory: str, prefix: str, first_step: int, length: int, width: int, height: int, fps: int = 31):
        import os

        import cv2
        os.makedirs(directory, exist_ok=True)
        self.path = os.path.join(directory, f"{prefix}-step-{first_step}-to-step-{first_step + length}.mp4")
        self._w = cv2.VideoWriter(self.path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (width, height))
        if not self._w.isOpened():
            raise RuntimeError(f"cannot open {self.path} for writing")
        self.frames = 0

    def write(self, rgb: np.ndarray):
        self._w.write(np.ascontiguousarray(rgb[:, :, ::-1]))   # OpenCV wants BGR
        self.frames += 1

    def close(self):
        self._w.release()
