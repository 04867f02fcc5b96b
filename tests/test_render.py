"""Head-less rasteriser behind `record` (reference main.py:127-171): its host-side kinematics against the oracle, and that
frames / video files come out."""
import numpy as np
import pytest

from conftest import make_oracle

cv2 = pytest.importorskip("cv2")


def test_renderer_kinematics_match_the_oracle(spec):
    from so100_mujoco_rl_b200.render import forward_kinematics
    o = make_oracle(1, 1)
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.uniform(spec.jnt_range[:, 0], spec.jnt_range[:, 1])
        pos, rot = forward_kinematics(spec, q)
        k = o.fk(q)
        assert np.abs(pos - k["xpos"]).max() < 1e-12
        assert np.abs(rot.reshape(6, 9) - k["xmat"]).max() < 1e-12


def test_frames_show_the_arm_and_the_block_and_video_is_written(spec, tmp_path):
    from so100_mujoco_rl_b200.render import SceneRenderer, VideoSink
    from so100_mujoco_rl_b200.tasks import REST_POSITION
    r = SceneRenderer(spec, 320, 240)
    a = r.render(np.array(REST_POSITION), np.array([0.0, -0.3, 0.01]), text="Env02 step 1")
    b = r.render(np.array([0.5, -1.5, 1.5, 0.5, 0.0, 1.0]), np.array([0.1, -0.25, 0.01]))
    assert a.shape == (240, 320, 3) and a.dtype == np.uint8
    assert (a != b).any() and len(np.unique(a.reshape(-1, 3), axis=0)) > 10      # not a blank canvas; pose changes the image
    green = (a[:, :, 1] > 150) & (a[:, :, 0] < 100) & (a[:, :, 2] < 100)
    assert green.sum() > 10                                                       # the block is in view
    sink = VideoSink(str(tmp_path), "rec-Env02", 0, 5, r.w, r.h)
    for _ in range(5):
        sink.write(a)
    sink.close()
    assert sink.path.endswith("rec-Env02-step-0-to-step-5.mp4")
    import os
    assert os.path.getsize(sink.path) > 1000 and sink.frames == 5
