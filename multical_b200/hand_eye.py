"""`HandEyeCalibration` with the reference's surface (multical/optimization/hand_eye.py:13-97): a `Calibration` whose motion model is
`HandEye`, initialised from the arm's gripper poses by OpenCV's robot-world hand-eye solver (multical/transform/hand_eye.py:20-50, a one-off
on a few hundred poses: stays on the host like in the reference) and then refined by `bundle_adjust` / `adjust_outliers` -- which is where
the GPU path comes in (the reference's other route into the hot path)."""
from functools import cached_property

import cv2
import numpy as np

from .motion import HandEye
from .pose_set import pose_table


def hand_eye_robot_world(world_wrt_camera, base_wrt_gripper):
  """transform/hand_eye.py:20-50: AX = ZB through cv2.calibrateRobotWorldHandEye -> (base_wrt_world, gripper_wrt_camera, per-pose residual)."""
  world_wrt_camera, base_wrt_gripper = np.asarray(world_wrt_camera, np.float64), np.asarray(base_wrt_gripper, np.float64)
  assert world_wrt_camera.shape[0] == base_wrt_gripper.shape[0]
  Rbw, tbw, Rgc, tgc = cv2.calibrateRobotWorldHandEye(world_wrt_camera[:, :3, :3], world_wrt_camera[:, :3, 3],
                                                      base_wrt_gripper[:, :3, :3], base_wrt_gripper[:, :3, 3])
  base_wrt_world, gripper_wrt_camera = np.eye(4), np.eye(4)
  base_wrt_world[:3, :3], base_wrt_world[:3, 3] = Rbw, np.ravel(tbw)
  gripper_wrt_camera[:3, :3], gripper_wrt_camera[:3, 3] = Rgc, np.ravel(tgc)
  # matrix.transform(a, b) = b @ a (transform/matrix.py): residual of world_wrt_camera @ base_wrt_world = gripper_wrt_camera @ base_wrt_gripper
  err = world_wrt_camera @ base_wrt_world - gripper_wrt_camera @ base_wrt_gripper
  return base_wrt_world, gripper_wrt_camera, np.linalg.norm(err, axis=(1, 2))


class HandEyeCalibration:
  def __init__(self, calib, gripper_wrt_base, world_wrt_camera):
    assert isinstance(calib.motion, HandEye) or hasattr(calib.motion, "base_wrt_gripper")
    self.gripper_wrt_base = gripper_wrt_base
    self.world_wrt_camera = world_wrt_camera
    self.calib = calib

  @staticmethod
  def initialise(calib, gripper_wrt_base):
    """optimization/hand_eye.py:21-38: frame poses of a static calibration + the arm's poses -> HandEye motion model; cameras and camera
    poses are fixed from here on (line 37)."""
    world_wrt_camera = np.asarray(calib.motion.frame_poses.poses)
    valid = np.asarray(calib.motion.frame_poses.valid)
    base_wrt_gripper = np.linalg.inv(np.asarray(gripper_wrt_base, np.float64))
    base_wrt_world, gripper_wrt_camera, _ = hand_eye_robot_world(world_wrt_camera[valid], base_wrt_gripper[valid])
    model = HandEye(pose_table(base_wrt_gripper, valid), np.linalg.inv(base_wrt_world), gripper_wrt_camera)
    calib = calib.copy(motion=model).enable(camera_poses=False, cameras=False)
    return HandEyeCalibration(calib, gripper_wrt_base, world_wrt_camera)

  valid = property(lambda self: self.calib.motion.valid)
  model = property(lambda self: self.calib.motion)
  gripper_wrt_camera = property(lambda self: self.model.gripper_wrt_camera)
  base_wrt_world = property(lambda self: np.linalg.inv(self.model.world_wrt_base))

  def bundle_adjust(self, **kwargs):
    return self.copy(calib=self.calib.bundle_adjust(**kwargs))

  def adjust_outliers(self, **kwargs):
    return self.copy(calib=self.calib.adjust_outliers(**kwargs))

  @cached_property
  def cameras_wrt_gripper(self):
    """optimization/hand_eye.py:82-87: per camera, the inverse of gripper_wrt_camera with that camera as master."""
    return {k: np.linalg.inv(self.calib.with_master(k).motion.gripper_wrt_camera) for k in self.calib.cameras.names}

  def __getstate__(self):
    return {k: self.__dict__[k] for k in ("gripper_wrt_base", "world_wrt_camera", "calib")}

  def __setstate__(self, d): self.__dict__.update(d)

  def copy(self, **k):
    d = self.__getstate__(); d.update(k)
    return self.__class__(**d)
