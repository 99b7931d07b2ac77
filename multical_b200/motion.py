"""The two non-static motion models of the reference, host side.

`RollingFrames` (multical/motion/rolling_frames.py:66-166): a rolling-shutter rig -- every frame has a start and an end pose and a
corner's camera-frame point is the blend of the point under either pose, weighted by the image row it was observed in.
`HandEye` (multical/motion/hand_eye.py:14-107): a rig on a robot arm -- the frame pose is
gripper_wrt_camera @ base_wrt_gripper[f] @ world_wrt_base with the arm poses given and the two outer transforms optimised.

Same names, constructor arguments, parameter order and copy-on-write behaviour as the reference classes; the projection
arithmetic itself lives in the kernels (csrc/kernels.cuh: `corner_point<ROLL>`, `hand_eye_frame`) and `Calibration` hands the
state over through `mcba_set_rolling` / `mcba_set_hand_eye` (include/mcba.h).  `StaticFrames` stays in pose_set.py.
"""
from functools import cached_property

import numpy as np

from . import rtvec
from ._native import MOTION_HAND_EYE, MOTION_ROLLING
from .parameters import Parameters
from .pose_set import pose_table as make_pose_table
from .structs import Table, struct


class RollingFrames(Parameters):
  """rolling_frames.py:66-166.  params = [start rtvecs (F x 6), end rtvecs (F x 6)] (135-140)."""
  engine_motion = MOTION_ROLLING

  def __init__(self, pose_start, pose_end, valid, names, max_iterations=4):
    self.pose_start = np.asarray(pose_start, np.float64)
    self.pose_end = np.asarray(pose_end, np.float64)
    self.valid = np.asarray(valid, bool)
    self.names = list(names)
    self.max_iterations = max_iterations

  @staticmethod
  def init(pose_table, names=None, max_iterations=4):
    """rolling_frames.py:105-111: both poses start at the static estimate."""
    size = np.size(pose_table.valid)
    names = names or [str(i) for i in range(size)]
    return RollingFrames(pose_table.poses, pose_table.poses, pose_table.valid, names, max_iterations=max_iterations)

  size = property(lambda self: self.pose_start.shape[0])
  poses = property(lambda self: self.pose_start)            # what the solver takes as "the frames": the start poses

  @cached_property
  def start_table(self): return Table.create(poses=self.pose_start, valid=self.valid)

  @cached_property
  def end_table(self): return Table.create(poses=self.pose_end, valid=self.valid)

  @property
  def frame_poses(self): return self.start_table           # rolling_frames.py:91-93

  pose_table = frame_poses

  def pre_transform(self, t):
    t = np.asarray(t)
    return self.copy(pose_start=t @ self.pose_start, pose_end=t @ self.pose_end)

  def post_transform(self, t):
    t = np.asarray(t)
    return self.copy(pose_start=self.pose_start @ t, pose_end=self.pose_end @ t)

  @cached_property
  def params(self):
    return [rtvec.from_matrix(self.pose_start).ravel(), rtvec.from_matrix(self.pose_end).ravel()]

  def with_params(self, params):
    start, end = [rtvec.to_matrix(np.asarray(m, np.float64).reshape(-1, 6)) for m in params]
    return self.copy(pose_start=start, pose_end=end)

  @property
  def num_params(self): return 2 * rtvec.size * self.size

  def with_param_vec(self, param_vec):
    param_vec = np.asarray(param_vec)
    assert param_vec.size == self.num_params, f"inconsistent parameter sizes, got {param_vec.size}, expected {self.num_params}"
    return self.with_params([param_vec[:param_vec.size // 2], param_vec[param_vec.size // 2:]])

  def export(self):
    return {i: struct(start=start.tolist(), end=end.tolist())
            for i, start, end, valid in zip(self.names, self.pose_start, self.pose_end, self.valid) if valid}

  def __getstate__(self):
    return {k: getattr(self, k) for k in ("pose_start", "pose_end", "valid", "names", "max_iterations")}

  def __setstate__(self, d): self.__dict__.update(d)

  def copy(self, **k):
    d = self.__getstate__(); d.update(k)
    return self.__class__(**d)


class HandEye(Parameters):
  """hand_eye.py:14-107.  params = struct(world_wrt_base rtvec, gripper_wrt_camera rtvec) (76-81); `base_wrt_gripper` is a
  Table(poses [F,4,4], valid [F]) of constants."""
  engine_motion = MOTION_HAND_EYE

  def __init__(self, base_wrt_gripper, world_wrt_base, gripper_wrt_camera, names=None):
    self.base_wrt_gripper = base_wrt_gripper
    n = np.shape(base_wrt_gripper.poses)[0]
    self.names = names or [str(i) for i in range(n)]
    self.world_wrt_base = np.asarray(world_wrt_base, np.float64)
    self.gripper_wrt_camera = np.asarray(gripper_wrt_camera, np.float64)

  size = property(lambda self: np.shape(self.base_wrt_gripper.poses)[0])
  valid = property(lambda self: np.asarray(self.base_wrt_gripper.valid))

  @cached_property
  def pose_table(self):
    """hand_eye.py:43-46."""
    poses = self.gripper_wrt_camera[None] @ np.asarray(self.base_wrt_gripper.poses, np.float64) @ self.world_wrt_base[None]
    return make_pose_table(poses, self.valid)

  frame_poses = property(lambda self: self.pose_table)
  poses = property(lambda self: self.pose_table.poses)

  def index_of(self, k):
    if isinstance(k, str):
      if k not in self.names: raise KeyError(f"pose {k} not found in {self.names}")
      return self.names.index(k)
    return k

  def __getitem__(self, k): return self.poses[self.index_of(k)]

  def relative(self, src, dest): return self[dest] @ np.linalg.inv(self[src])

  def pre_transform(self, t): return self.copy(gripper_wrt_camera=np.asarray(t) @ self.gripper_wrt_camera)

  def post_transform(self, t): return self.copy(world_wrt_base=self.world_wrt_base @ np.asarray(t))

  @cached_property
  def params(self):
    return struct(world_wrt_base=rtvec.from_matrix(self.world_wrt_base), gripper_wrt_camera=rtvec.from_matrix(self.gripper_wrt_camera))

  def with_params(self, params):
    return self.copy(world_wrt_base=rtvec.to_matrix(np.asarray(params.world_wrt_base)),
                     gripper_wrt_camera=rtvec.to_matrix(np.asarray(params.gripper_wrt_camera)))

  num_params = 2 * rtvec.size

  def with_param_vec(self, param_vec):
    param_vec = np.asarray(param_vec)
    assert param_vec.size == self.num_params, f"inconsistent parameter sizes, got {param_vec.size}, expected {self.num_params}"
    return self.with_params(struct(world_wrt_base=param_vec[:6], gripper_wrt_camera=param_vec[6:]))

  def export(self):
    arm = {n: p.tolist() for n, p, v in zip(self.names, np.asarray(self.base_wrt_gripper.poses), self.valid) if v}
    return struct(base_wrt_gripper=arm, world_wrt_base=self.world_wrt_base.tolist(), gripper_wrt_camera=self.gripper_wrt_camera.tolist())

  def __getstate__(self):
    return {k: getattr(self, k) for k in ("base_wrt_gripper", "gripper_wrt_camera", "world_wrt_base", "names")}

  def __setstate__(self, d): self.__dict__.update(d)

  def copy(self, **k):
    d = self.__getstate__(); d.update(k)
    return self.__class__(**d)


def motion_kind(motion):
  """0 (static), MOTION_ROLLING or MOTION_HAND_EYE -- for this package's classes and, by duck typing, the reference's own
  (RollingFrames has pose_start / pose_end, HandEye has base_wrt_gripper)."""
  kind = getattr(motion, "engine_motion", None)
  if kind is not None: return kind
  if hasattr(motion, "pose_start") and hasattr(motion, "pose_end"): return MOTION_ROLLING
  if hasattr(motion, "base_wrt_gripper"): return MOTION_HAND_EYE
  return 0
